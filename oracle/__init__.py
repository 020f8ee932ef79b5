"""CPU oracle for the SoundML spectral hot path.

TEST INFRASTRUCTURE ONLY.  This package restates the reference's arithmetic
(gabyfle/SoundML, OCaml + nx) in numpy float64 so the CUDA path can be checked
against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product
(``soundml_b200``) must never route through it.

Parity pinning: every function here is validated against the reference's own
committed librosa-0.11 golden vectors (``tests/golden/reference_vectors.npz``,
produced from ``/root/reference/soundml/test/*/vectors/*.json`` by
``tests/golden/make_golden.py``) and, for the resampler, against the
reference's C executor compiled from its own source (``oracle/_ref``).
"""
