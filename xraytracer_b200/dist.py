"""Host-side multi-GPU plumbing (one process per GPU, torch.distributed): the path shards by SAMPLES — rank r of R
renders sample indices [r*spp/R, (r+1)*spp/R) of every pixel with the scene replicated (counter RNG keyed by sample
index, so the union over ranks is the 1-GPU sample set) — and the per-rank SUM buffers meet in ONE reduce, after which
rank 0 applies the reference's `image /= n_samples` (renderer.cpp:98). No other collective exists on the path."""
from __future__ import annotations

import torch
import torch.distributed as dist


def sample_range(spp_total: int, rank: int, world: int):
    """[lo, hi) sample indices of `rank`; ranges tile [0, spp_total) exactly."""
    return rank * spp_total // world, (rank + 1) * spp_total // world


def reduce_image(partial_sum: torch.Tensor, spp_total: int, dst: int = 0) -> torch.Tensor:
    """Sum-reduce the per-rank per-pixel sums to `dst` (NCCL on GPUs, gloo on CPU) and divide by the TOTAL spp there."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(partial_sum, dst=dst, op=dist.ReduceOp.SUM)
        if dist.get_rank() != dst:
            return partial_sum
    return partial_sum.div_(float(spp_total))
