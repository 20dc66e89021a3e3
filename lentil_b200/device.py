"""Device plumbing: PyTorch supplies device memory, streams and pinned staging; every kernel is
ours and is reached through the C ABI (lentil_b200._lib).  There is no CPU fallback: calling a
compute entry point without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib

_device = None


class NoDeviceError(RuntimeError):
    pass


def device():
    """The CUDA device this process computes on (cuda:LOCAL_RANK unless set_device was called)."""
    global _device
    if _device is None:
        if not torch.cuda.is_available():
            raise NoDeviceError("lentil_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        _device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
        torch.cuda.set_device(_device)
        _lib.lib()  # fail now, loudly, if the CUDA library has not been built
    return _device


def set_device(index):
    global _device
    if not torch.cuda.is_available():
        raise NoDeviceError("lentil_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    _device = torch.device("cuda", int(index))
    torch.cuda.set_device(_device)
    _lib.lib()
    return _device


def bind_cpu_affinity(index=None):
    """Pin this process to the CPU cores NVML reports as local to GPU `index` (its NUMA node), so that pinned staging
    buffers are first-touched next to the GPU's PCIe root and host<->device copies do not cross sockets.  One process per
    GPU deployments (torchrun) call it once before allocating pinned memory.  Returns the CPU set, or None when NVML or
    sched_setaffinity is unavailable (nothing is changed then)."""
    try:
        import pynvml
        idx = device().index if index is None else int(index)
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or None
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr():
    """cudaStream_t of torch's current stream, as an int for ctypes."""
    dev = device()
    if _raw_stream is not None:
        return _raw_stream(dev.index)
    return torch.cuda.current_stream(dev).cuda_stream


def is_dev(x):
    return isinstance(x, torch.Tensor)


def empty_c128(*shape):
    return torch.empty(*shape, dtype=torch.complex128, device=device())


def zeros_f64(*shape):
    return torch.zeros(*shape, dtype=torch.float64, device=device())


def empty_bytes(nbytes):
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device())


def to_dev(arr, dtype=None):
    """Host array -> device tensor (H2D on the current stream; pinned sources go async)."""
    a = np.ascontiguousarray(arr, dtype=dtype)
    t = torch.from_numpy(a) if a.flags.writeable else torch.from_numpy(a.copy())
    return t.to(device(), non_blocking=True)


def to_host(t):
    """Device tensor -> fresh numpy array.  Large transfers land in a pinned buffer from torch's caching host
    allocator (async D2H at full PCIe rate) that the returned array owns: no second host copy, and the block goes
    back to the allocator's cache when the array is dropped."""
    t = t.detach()
    nbytes = t.numel() * t.element_size()
    if nbytes < (1 << 20) or not t.is_contiguous():
        return t.cpu().numpy()
    pin = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    pin.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return pin.numpy()


def ld_of(t):
    """Leading dimension (elements between rows) of a 2-D device array with unit column stride."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError("device field must be 2-D with contiguous rows")
    return int(t.stride(0)) if t.shape[0] > 1 else int(max(t.shape[1], t.stride(0)))


def launch_count():
    return int(_lib.lib().lfd_launch_count())


def probe_fp64(iters=20000):
    """Measured DMMA / DFMA issue rates of the current device (TFLOP/s) and its SM clock."""
    device()
    out = (C.c_double * 3)()
    _lib.check(_lib.lib().lfd_probe_fp64(out, int(iters)), "lfd_probe_fp64")
    return {"dmma_tflops": out[0], "dfma_tflops": out[1], "clock_mhz": out[2]}
