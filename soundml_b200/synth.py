"""Deterministic synthetic audio: per clip b, sample i

    x[b, i] = 0.6 * sin(2 pi f_b i / sr) + 0.1 * u(b, i),   f_b = 110 * 2^((b mod 60)/12) Hz

with u uniform in [-1, 1) from a counter-based hash (splitmix64 finaliser of
(seed, b, i), top 24 bits), cast to float32.  The numpy and torch versions
produce identical bits, so host oracles and device benchmarks see one signal.
"""
import numpy as np

MASK = (1 << 64) - 1
C1, C2 = 0xBF58476D1CE4E5B9, 0x94D049BB133111EB
GOLDEN = 0x9E3779B97F4A7C15


def _mix_np(z):
    z = (z ^ (z >> np.uint64(30))) * np.uint64(C1)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(C2)
    return z ^ (z >> np.uint64(31))


def clips_numpy(batch, n, sample_rate=22050, seed=42, first_clip=0):
    b = np.arange(first_clip, first_clip + batch, dtype=np.uint64)[:, None]
    i = np.arange(n, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) * np.uint64(GOLDEN)) ^ (b << np.uint64(32)) ^ i
        h = _mix_np(key + np.uint64(GOLDEN))
    u = (h >> np.uint64(40)).astype(np.float64) / float(1 << 23) - 1.0
    f = 110.0 * np.exp2(((b % np.uint64(60)).astype(np.float64)) / 12.0)
    phase = (2.0 * np.pi / float(sample_rate)) * f * i.astype(np.float64)
    return (0.6 * np.sin(phase) + 0.1 * u).astype(np.float32)


def _lsr(z, k):
    # logical shift right of a two's-complement int64 tensor
    return (z >> k) & ((1 << (64 - k)) - 1)


def _as_i64(v):
    v &= MASK
    return v - (1 << 64) if v >= (1 << 63) else v


def clips_torch(batch, n, device, sample_rate=22050, seed=42, first_clip=0, chunk=64):
    """Same signal generated on `device` (int64 arithmetic wraps like uint64)."""
    import torch
    out = torch.empty((batch, n), dtype=torch.float32, device=device)
    i = torch.arange(n, dtype=torch.int64, device=device)[None, :]
    base = _as_i64(seed * GOLDEN)
    for s in range(0, batch, chunk):
        e = min(batch, s + chunk)
        b = torch.arange(first_clip + s, first_clip + e, dtype=torch.int64, device=device)[:, None]
        z = (base ^ (b << 32) ^ i) + _as_i64(GOLDEN)
        z = (z ^ _lsr(z, 30)) * _as_i64(C1)
        z = (z ^ _lsr(z, 27)) * _as_i64(C2)
        z = z ^ _lsr(z, 31)
        u = _lsr(z, 40).to(torch.float64) / float(1 << 23) - 1.0
        f = 110.0 * torch.exp2((b % 60).to(torch.float64) / 12.0)
        phase = (2.0 * np.pi / float(sample_rate)) * f * i.to(torch.float64)
        out[s:e] = (0.6 * torch.sin(phase) + 0.1 * u).to(torch.float32)
    return out
