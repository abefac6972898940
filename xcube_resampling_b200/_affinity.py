"""Pin the calling thread to the CPUs next to a GPU before it page-locks host memory.

Page-locked buffers are placed by first touch: a thread that runs on the other socket puts the
staging and output buffers of "its" GPU behind the inter-socket link, and every H2D / D2H copy of
the band pipeline then crosses it.  With one process (or thread) per GPU the cure is to bind the
worker to the CPU set NVML reports for the device (``nvmlDeviceGetCpuAffinity``) before the
first allocation.  Binding is best effort: no NVML, a cpuset that does not intersect the device's
CPUs (containers), or a single-node host leave the thread where it is.
"""

from __future__ import annotations

import os


def device_cpus(device_index: int) -> set[int]:
    """CPUs NVML lists as local to the device, intersected with what this process may use."""
    try:
        import pynvml
        pynvml.nvmlInit()
        # CUDA_VISIBLE_DEVICES re-numbers devices; NVML does not
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            ident = ids[device_index]
            handle = (pynvml.nvmlDeviceGetHandleByUUID(ident) if ident.startswith("GPU-")
                      else pynvml.nvmlDeviceGetHandleByIndex(int(ident)))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
    except Exception:  # no NVML, no such device: nothing to bind to
        return set()
    cpus = {64 * k + b for k, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
    return cpus & os.sched_getaffinity(0)


def bind_to_device(device_index: int) -> set[int]:
    """Restrict the CALLING THREAD to the device's local CPUs; returns the set used (empty = not bound)."""
    cpus = device_cpus(device_index)
    if cpus and cpus != os.sched_getaffinity(0):
        try:
            os.sched_setaffinity(0, cpus)  # pid 0 = the calling thread on Linux
        except OSError:
            return set()
    return cpus
