"""world_size-2 gloo run (CPU) of the multi-GPU host logic: contiguous frame sharding, one-off calibration
broadcast, max-over-ranks timing reduction."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imgprocessor_b200 import sharding


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 255, 256, 4096):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(1)
        dark = rng.random((6, 8)).astype(np.float32)
        K = np.arange(9, dtype=np.float64).reshape(3, 3)
        maps = {'dark': dark, 'flat': None, 'K': K} if rank == 0 else {'dark': None, 'flat': None, 'K': None}
        got = sharding.broadcast_calibration(maps, src=0)
        ok = got['flat'] is None and np.array_equal(got['dark'].numpy(), dark) and np.array_equal(got['K'].numpy(), K)
        ok = ok and got['dark'].dtype == torch.float32 and got['K'].dtype == torch.float64
        lo, hi = sharding.shard_range(11, world, rank)
        total = sharding.reduce_sum(hi - lo)
        slowest = sharding.reduce_max(1.0 + rank)
        ret[rank] = (ok, lo, hi, total, slowest)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_broadcast_and_reduce():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0][0] and ret[1][0]
    assert (ret[0][1], ret[0][2]) == (0, 6) and (ret[1][1], ret[1][2]) == (6, 11)
    assert ret[0][3] == ret[1][3] == 11.0
    assert ret[0][4] == ret[1][4] == 2.0
