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


def bind_to_gpu_numa(device_index: int) -> Dict[str, object]:
    """Pin the calling process to the CPUs (and thereby, through first-touch, its page-locked buffers to the memory) of the
    NUMA node the GPU hangs off.  The host-buffer session moves q / qdot over PCIe every cycle; a rank whose staging memory
    sits on the other socket pays the inter-socket link for every byte, and with one rank per GPU the ranks of the far
    socket then all share that link.  Uses NVML's ideal CPU affinity for the device (``nvmlDeviceGetCpuAffinity``); a no-op
    (reported as such) when NVML or ``sched_setaffinity`` is unavailable or ``VFK_NO_NUMA_BIND`` is set.
    Returns {"bound": bool, "cpus": n, "was": n_before, "why": str}."""
    import os
    info: Dict[str, object] = {"bound": False, "cpus": None, "was": None, "why": ""}
    if os.environ.get("VFK_NO_NUMA_BIND"):
        info["why"] = "VFK_NO_NUMA_BIND set"
        return info
    if not hasattr(os, "sched_setaffinity"):
        info["why"] = "no sched_setaffinity"
        return info
    try:
        import pynvml
        pynvml.nvmlInit()
        phys = int(device_index)
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            try:
                phys = int(visible.split(",")[device_index])
            except (ValueError, IndexError):
                phys = int(device_index)
        handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        before = os.sched_getaffinity(0)
        allowed = cpus & before
        info["was"] = len(before)
        if not allowed:
            info["why"] = "GPU-local CPUs are outside this process's cpuset"
            return info
        if allowed == before:
            info.update(bound=True, cpus=len(allowed), why="already local")
            return info
        os.sched_setaffinity(0, allowed)
        info.update(bound=True, cpus=len(allowed), why="nvmlDeviceGetCpuAffinity")
    except Exception as e:                                           # noqa: BLE001
        info["why"] = "nvml: %r" % (e,)
    return info
