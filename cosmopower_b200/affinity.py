"""CPU / NUMA placement of the host threads that feed one GPU.

The host-buffer path moves 10 KB per spectrum over PCIe into pinned host memory.  On a two-socket box a rank
whose threads (and therefore the pinned pages they first touch) sit on the other socket pays a trip over the
inter-socket link for every DMA write; with eight ranks that link, not PCIe, bounds the aggregate.  Each rank
therefore binds itself to the CPUs NVML reports as local to its GPU before it allocates pinned memory.
"""
import os


def device_cpus(device_index):
    """CPUs local to CUDA device `device_index` (NVML's ideal affinity), or None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if vis:                                   # NVML enumerates the physical devices
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        return [c for c in cpus if c < n_cpu] or None
    except Exception:
        return None


def pin_to_device(device_index):
    """Bind the calling process to the CPUs next to its GPU (intersected with the CPUs it is allowed to use).
    Returns the CPU list applied, or None if nothing was changed."""
    cpus = device_cpus(device_index)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        want = sorted(set(cpus) & set(allowed))
        if not want or set(want) == set(allowed):
            return None
        os.sched_setaffinity(0, want)
        return want
    except Exception:
        return None
