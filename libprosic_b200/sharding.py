"""Region sharding of the hot path across GPUs.

Pairs (part 1) and sites (part 2) are independent, so each rank owns a contiguous slice of the
candidate records; there is no collective on the data path.  The only exchange is the final
gather of per-site result arrays to rank 0 in record order (for the unchanged FDR filtration).
Constraint from the reference: all breakends of one EVENT share results
(calling.rs:415-421, 534-549), so a group must not straddle a shard boundary.
"""
import numpy as np


def shard_bounds(weights, world, group_ids=None):
    """Contiguous partition of records 0..n-1 into `world` slices balanced by `weights`
    (estimated cells or likelihood evaluations).  Returns bounds[world+1].  If `group_ids` is
    given, records with the same non-negative id are contiguous and never split."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    bounds = np.zeros(world + 1, dtype=np.int64)
    bounds[world] = n
    if n == 0:
        return bounds
    csum = np.concatenate([[0.0], np.cumsum(w)])
    total = csum[-1]
    for r in range(1, world):
        b = int(np.searchsorted(csum, total * r / world, side="left"))
        b = min(max(b, int(bounds[r - 1])), n)
        if group_ids is not None:
            g = np.asarray(group_ids)
            while 0 < b < n and g[b] >= 0 and g[b] == g[b - 1]:
                b += 1  # move the cut to the end of the group
        bounds[r] = b
    return np.maximum.accumulate(bounds)


def my_slice(bounds, rank):
    return int(bounds[rank]), int(bounds[rank + 1])


def gather_in_record_order(local, bounds, rank, world, dist=None, dst=0):
    """Gather per-record arrays (first axis = records of this rank's slice) to rank `dst`.
    Returns the concatenated array on `dst`, None elsewhere.  `dist` is torch.distributed (any
    backend); with world == 1 nothing is communicated."""
    local = np.ascontiguousarray(local)
    assert local.shape[0] == bounds[rank + 1] - bounds[rank]
    if world == 1 or dist is None:
        return local
    import torch
    parts = [None] * world if rank == dst else None
    dist.gather_object(local, parts, dst=dst)
    if rank != dst:
        return None
    return np.concatenate(parts, axis=0)


def bind_host_to_device(device_index, pci_bus_id=None):
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index` (its NUMA node),
    so that the pinned host buffers allocated afterwards are first-touched on that node: with one
    process per GPU, eight ranks streaming host buffers no longer pull them all through one socket's
    memory controller and inter-socket link.  Returns a short description, or None when NVML or the
    affinity call is unavailable (nothing is changed then).  `pci_bus_id` ("domain:bus:device.fn") selects
    the NVML device when CUDA_VISIBLE_DEVICES renumbers the CUDA indices."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = (pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode() if isinstance(pci_bus_id, str) else pci_bus_id)
                 if pci_bus_id else pynvml.nvmlDeviceGetHandleByIndex(int(device_index)))
            n_cpu = os.cpu_count() or 1
            words = (n_cpu + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        finally:
            pynvml.nvmlShutdown()
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return "%d of %d cores (NVML affinity of GPU %d)" % (len(cpus), len(allowed), int(device_index))
    except Exception:  # no NVML, no sched_setaffinity, containerised cpuset: leave the process alone
        return None
