"""Batch sharding of the prototype path across the GPUs of one node.

Pixels are independent given identical prototypes, and the EMA update is a sum over pixels, so
the batch is split into whole images per rank and the only exchange is ONE sum all-reduce per
step of the `sums` buffer ``[C*D class sums | C*D sums of squares | C counts | 8 statistics]``
(39 KB at D=256).  Every rank then applies the same blend, so prototypes stay bit-identical
everywhere, and the statistics tail gives global batch means that a replicated host-side Monitor
turns into identical switch decisions on every rank.  The reference is single-GPU; this module is
the new layer north_star asks for.  Works on any device (the CPU gloo tests exercise it).
"""
from __future__ import annotations

import torch

NUM_STATS = 8
STAT_PROTO_CONF, STAT_PRIOR_CONF, STAT_PL_CONF, STAT_PL_PIXELS, STAT_PIXELS, STAT_ENTROPY = range(6)


def shard_bounds(n_items: int, world: int, rank: int):
    """[start, end) of the items (images) owned by ``rank``: contiguous, sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(tensors, world: int, rank: int):
    """Slices every (B, ...) tensor of ``tensors`` to this rank's images."""
    out = []
    for t in tensors:
        s, e = shard_bounds(t.shape[0], world, rank)
        out.append(t[s:e])
    return out


def sums_numel(C: int, D: int) -> int:
    return 2 * C * D + C + NUM_STATS


def allreduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the per-rank ``sums`` buffers over ``group`` (no-op without a process group)."""
    import torch.distributed as dist
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return sums
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def split_sums(sums: torch.Tensor, C: int, D: int):
    """Views of a ``sums`` buffer: (class sums (C,D), sums of squares (C,D), counts (C,), statistics (8,))."""
    cd = C * D
    return sums[:cd].view(C, D), sums[cd:2 * cd].view(C, D), sums[2 * cd:2 * cd + C], sums[2 * cd + C:]


def stats_from_tail(tail) -> dict:
    """Batch means from the statistics tail (host floats).  ``tail`` is a sequence of 8 numbers."""
    n = tail[STAT_PIXELS]
    inv = 1.0 / n if n > 0 else float("nan")
    return {
        "prototypes": tail[STAT_PROTO_CONF] * inv,               # mean_n max_k softmax(-d/tau)
        "prior": tail[STAT_PRIOR_CONF] * inv,                    # mean_n max_k prior
        "pseudolabel confidence": tail[STAT_PL_CONF] * inv,      # mean_n max_k rectified posterior
        "pseudolabel_pixel_num": tail[STAT_PL_PIXELS],           # pixels whose label is not 255
        "entropy": tail[STAT_ENTROPY] * inv,
        "pixels": n,
    }
