"""Data-parallel plumbing of the training step (one process per GPU, torch.distributed).

The path has exactly one exchange step: a sum all-reduce of the single flat fp32 gradient bucket after the
backward pass (NCCL over NVLink on GPUs; gloo in the CPU tests).  The 1/world scaling is folded into the
optimiser kernel's clip coefficient (cg_optim_advance: grad_scale), so the bucket is touched once.
Every sample's loss term is independent and the loss is a batch mean (reference src/vae.py:454-457), hence
mean-of-rank-means over equal shards == the single-process full-batch gradient.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_batch(n_global: int, world: int, rank: int):
    """contiguous equal shards; the reference's DataLoader uses drop_last=True so batches divide evenly"""
    if n_global % world != 0:
        raise ValueError(f"global batch {n_global} is not divisible by world size {world}")
    per = n_global // world
    return rank * per, (rank + 1) * per


def broadcast_params_(flat_params: torch.Tensor, src: int = 0):
    world, _ = world_info()
    if world > 1:
        dist.broadcast(flat_params, src=src)
    return flat_params


def reduce_gradients_(flat_grad: torch.Tensor) -> float:
    """in-place SUM all-reduce of the flat bucket; returns the scale (1/world) the optimiser applies"""
    world, _ = world_info()
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / world


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    """in-place SUM all-reduce of a small statistics tensor (kl_free_bits per-channel KL sums)"""
    world, _ = world_info()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def rank_noise_seed(base_seed: int, rank: int) -> int:
    """rank-disjoint Philox key so shards draw independent eps (SURVEY 8e caveat 2)"""
    return (base_seed * 0x9E3779B97F4A7C15 + (rank << 48)) % (1 << 64)
