"""Data-parallel plumbing for the fused step: one process per GPU, torch.distributed (NCCL) for the one
exchange the path has -- the gradient all-reduce -- with the semantics Lightning's DDP strategy would
give the reference: gradients are AVERAGED over ranks, the clip norm is taken on the
averaged gradient, BatchNorm statistics and logged scalars stay per rank, every rank must step the same
species.  The functions are device agnostic (gloo on CPU tensors in the tests, NCCL on the flat CUDA
gradient buffers in the engine)."""
from __future__ import annotations

import random
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def rank() -> int:
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def allreduce_sum_(flat: torch.Tensor, async_op: bool = False):
    """In-place SUM all-reduce of a flat gradient buffer (or a contiguous slice of it).  The 1/world
    factor is NOT applied here: the fused clip+Adam kernel folds it into its gradient scale, and
    ``reduced_norm`` applies it to the logged norm."""
    if world_size() == 1:
        return None
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op)


def reduced_norm(sum_sq_of_summed_grads: float, world: Optional[int] = None) -> float:
    """||mean_r g_r||_2 from sum(( sum_r g_r )^2)."""
    w = world or world_size()
    return (sum_sq_of_summed_grads ** 0.5) / w


def chunk_bounds(n: int, boundaries: Sequence[int]) -> List[tuple]:
    """Split [0,n) at the given sorted offsets into contiguous (lo, hi) chunks, dropping empty ones.
    Used to all-reduce the part of a flat gradient buffer that is already final while the rest of the
    backward pass still runs."""
    cuts = [0] + [b for b in boundaries if 0 < b < n] + [n]
    return [(lo, hi) for lo, hi in zip(cuts[:-1], cuts[1:]) if hi > lo]


def species_for_step(step: int, species: Sequence[str], seed: int = 0) -> str:
    """The species every rank trains at ``step``: the reference interleaves species with
    ``random.choice`` on identically seeded ranks (multi_modal_loader.py:57-61); here the choice is a
    pure function of (seed, step) so ranks cannot drift apart."""
    return random.Random(seed * 1000003 + step).choice(list(species))


def shard_rows(rows: int, world: int, align: int = 128) -> int:
    """rows of a row-sharded matrix owned by each rank: ceil(rows / world) rounded up to ``align`` (the tile
    height of the tensor-pipe SpMM), so rank r owns rows [r * n, (r + 1) * n) of the matrix padded to
    world * n rows.  The last ranks may own only padding."""
    per = (rows + world - 1) // world
    return (per + align - 1) // align * align


def host_allreduce_max(values: Sequence[int]) -> List[int]:
    """element-wise MAX of a few host integers over the ranks (set-up time only)"""
    if world_size() == 1:
        return [int(v) for v in values]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [int(v) for v in t.tolist()]
