"""Multi-GPU sharding of the batch axis.

Clips are independent units — "leading axes broadcast, so a batch of clips is
one call" (stft.mli:216-218), per-slice equality pinned by
soundml/test/stft/stft_grid.ml:180-205 — so the path shards with **no
data-path collective**: one process per GPU, rank r owns a contiguous range of
the flattened leading axis.  ``torch.distributed`` (NCCL on GPUs, gloo in the
CPU tests) is used only for the optional final gather and for timing.
"""


def shard_bounds(total, world):
    """Contiguous split, B_g = ceil(B / G): [(start, stop)] per rank."""
    per = -(-total // world) if world > 0 else 0
    return [(min(r * per, total), min((r + 1) * per, total)) for r in range(world)]


def my_shard(total, rank=None, world=None):
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    return shard_bounds(total, world)[rank]


def gather_shards(local, total, group=None):
    """Optional final gather: every rank receives the [total, ...] result.
    ``local`` is this rank's [stop - start, ...] tensor.  Uneven shards are
    padded to the common length for the all-gather and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    bounds = shard_bounds(total, world)
    per = max(b - a for a, b in bounds)
    padded = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, bounds)], dim=0)


def max_over_ranks(value, device="cpu", group=None):
    """Timing reduction used by bench.py: the slowest rank defines the step."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
