"""Multi-GPU plumbing: instances shard by rank; the only collective is the final statistics reduction.

Instances are independent (one robot, goal and obstacle set each), so the control-cycle path has no
exchange step: rank r of W owns the contiguous instance range ``shard_range(n_total, r, W)`` and runs the
same kernel on it.  ``reduce_stats`` is the one collective (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Tuple


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ``[begin, end)`` of rank's instances (first ``n_total % world`` ranks get one extra)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(int(n_total), int(world))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def reduce_stats(local: Dict[str, float], device="cpu") -> Dict[str, float]:
    """Keys ending in ``_max`` are max-reduced, ``_min`` min-reduced, everything else summed."""
    import torch
    import torch.distributed as dist
    keys = sorted(local)
    out = {}
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {k: float(local[k]) for k in keys}
    for op, sel in ((dist.ReduceOp.MAX, lambda k: k.endswith("_max")), (dist.ReduceOp.MIN, lambda k: k.endswith("_min")),
                    (dist.ReduceOp.SUM, lambda k: not (k.endswith("_max") or k.endswith("_min")))):
        ks = [k for k in keys if sel(k)]
        if not ks:
            continue
        t = torch.tensor([float(local[k]) for k in ks], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=op)
        out.update({k: float(v) for k, v in zip(ks, t.tolist())})
    return out
