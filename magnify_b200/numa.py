"""Bind a rank to the CPUs (and therefore, by first touch, the host memory) closest to its GPU.

The staging loop is PCIe/host-memory bound: with 8 ranks on a two-socket box, pinned buffers that
land on the wrong NUMA node make every H2D/D2H copy cross the socket interconnect.  NVML reports
the ideal CPU set per GPU; setting the affinity before allocating pinned memory keeps it local.
"""
from __future__ import annotations

import os


def bind_to_gpu_numa(device_index: int) -> list[int]:
    """Returns the CPU list the process was bound to ([] when NVML or the affinity call is
    unavailable -- binding is an optimisation, never a requirement)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = device_index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if device_index < len(ids) and ids[device_index].isdigit():
                    index = int(ids[device_index])
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            n_cpu = os.cpu_count() or 1
            words = (n_cpu + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        finally:
            pynvml.nvmlShutdown()
        cpus = [w * 64 + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1 and w * 64 + b < n_cpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []
