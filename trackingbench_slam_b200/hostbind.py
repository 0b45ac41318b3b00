"""Host-side placement for the end-to-end path: bind the calling process to the CPUs local to its GPU.

The end-to-end leg is bound by pinned-host -> device copies (2 MB per 1080p frame).  On a multi-socket host a pinned buffer
whose pages sit on the other socket crosses the socket interconnect on every upload; binding the process to the GPU's local
CPUs BEFORE it allocates pinned memory puts the pages (first touch) and the submitting thread next to the GPU's PCIe root.
Plumbing only (NVML + sched_setaffinity); no data-path code lives here."""
import os


def gpu_local_cpus(index):
    """CPU ids NVML reports as local to GPU `index` (empty list when NVML is unavailable)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1]
        return [c for c in cpus if c < n_cpu]
    except Exception:
        return []


def bind_to_gpu(index):
    """sched_setaffinity to the GPU's local CPUs (intersected with what the process may use). Returns the CPU list or None."""
    if os.environ.get('TB_BIND_NUMA', '1') == '0':
        return None
    try:
        allowed = os.sched_getaffinity(0)
        cpus = sorted(set(gpu_local_cpus(index)) & allowed)
        if not cpus or len(cpus) == len(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
