"""The N>1 path on CPU: the spp split and the sum-reduce + 1/spp scale that bench.py runs over NCCL are exercised
with world_size 2 on the gloo backend (no GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xraytracer_b200 import dist as xdist


@pytest.mark.parametrize("spp,world", [(1024, 1), (1024, 2), (1024, 8), (7, 2), (5, 8), (1, 4)])
def test_sample_ranges_partition_exactly(spp, world):
    ranges = [xdist.sample_range(spp, r, world) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == spp
    for a, b in zip(ranges, ranges[1:]):
        assert a[1] == b[0]
    assert sum(hi - lo for lo, hi in ranges) == spp


def _worker(rank, world, port, spp, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = xdist.sample_range(spp, rank, world)
    # each "sample" k contributes the deterministic image f(k); the rank sums its own range like the SUM_ONLY render
    H, W = 6, 5
    part = torch.zeros(H, W, 3, dtype=torch.float32)
    for k in range(lo, hi):
        part += torch.full((H, W, 3), float(k + 1)) * torch.arange(1, 4, dtype=torch.float32)
    img = xdist.reduce_image(part, spp_total=spp, dst=0)
    if rank == 0:
        out_q.put(img.numpy())
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("spp", [8, 5])
def test_two_rank_gloo_reduce_equals_single_rank_mean(spp):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, spp, q)) for r in range(2)]
    for p in procs:
        p.start()
    img = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = sum(range(1, spp + 1)) / spp * np.arange(1, 4, dtype=np.float32)
    assert np.allclose(img, np.broadcast_to(expect, img.shape), rtol=1e-6)
