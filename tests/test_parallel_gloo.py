"""World-size-2 gloo test (CPU) of the batch sharding used at N > 1: the shard
bounds partition the clips exactly, ranks work without exchanging data, and the
optional final gather reassembles the batch in order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from soundml_b200 import parallel


def test_shard_bounds_partition_exactly():
    for total in (0, 1, 2, 7, 1024, 1025, 4096, 65536):
        for world in (1, 2, 3, 4, 8):
            b = parallel.shard_bounds(total, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in b) == -(-total // world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, frames):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop = parallel.my_shard(total)
        # stand-in for the per-clip work: a value only the owner of clip b can produce
        clips = torch.arange(start, stop, dtype=torch.float32)
        local = clips[:, None, None] * 10.0 + torch.arange(frames, dtype=torch.float32)[None, None, :]
        full = parallel.gather_shards(local, total)
        want = (torch.arange(total, dtype=torch.float32)[:, None, None] * 10.0 +
                torch.arange(frames, dtype=torch.float32)[None, None, :])
        assert full.shape == want.shape and torch.equal(full, want)
        slowest = parallel.max_over_ranks(1.0 + rank)
        assert slowest == float(world)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [7, 1024])
def test_two_ranks_shard_and_gather(total):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, total, 5), nprocs=2, join=True)
