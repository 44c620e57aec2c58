"""World-size-2 gloo test (CPU) of the multi-GPU host logic: contiguous image sharding + ragged result gather + the
max-over-ranks timing reduction used by bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import amodal_depth_anything_b200  # noqa: F401
from amodal_depth_anything_b200.shard import gather_shards, max_over_ranks, shard_range


def test_shard_range_partitions():
    for n in (1, 4, 5, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(n_total * 6, dtype=torch.float32).view(n_total, 1, 2, 3)
    lo, hi = shard_range(n_total, rank, world)
    got = gather_shards(full[lo:hi].clone(), n_total)
    assert torch.equal(got, full)
    assert max_over_ranks(float(rank + 1), "cpu") == float(world)
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 5), nprocs=2, join=True)
