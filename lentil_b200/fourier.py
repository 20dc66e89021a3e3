"""Matrix-triple-product DFT with independent input/output shape, sampling, sub-pixel output
shift and integer input offset — same call surface as lentil/fourier.py (dft2 :5-103, idft2
:124-198), executed by K2a (lfd_mft_c128, FP64 DMMA with on-the-fly twiddles).

    F[k,l] = sqrt|ar ac| * sum_ij  e^{-2 pi i ar (R_i+off_r)(U_k-shift_r)} f[i,j]
                                   e^{-2 pi i ac (S_j+off_c)(V_l-shift_c)}
    R_i = i - floor(m/2), S_j = j - floor(n/2), U_k = k - floor(M/2), V_l = l - floor(N/2)
"""
import numpy as np

from . import _lib, device


def _pair(v):
    if np.ndim(v) == 0:
        return v, v
    a, b = v
    return a, b


_workspaces = {}


def _workspace(nbytes):
    """Grow-only scratch buffer per (device, CUDA stream): launches on one stream are ordered, so the planes of the next
    call may reuse the intermediates of the previous one (saves an allocation per call of the drop-in loop)."""
    if nbytes > (256 << 20):            # big batches: leave it to torch's caching allocator
        return device.empty_bytes(nbytes)
    key = (device.device().index, device.stream_ptr())      # the default stream is 0 on every device: key by device too
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if len(_workspaces) > 8:
            _workspaces.clear()
        ws = _workspaces[key] = device.empty_bytes(max(nbytes, 1 << 20))
    return ws


EXECUTIONS = {None: 0, 'default': 0, 'direct': 1, 'folded': 2, 'czt': 3, 'auto': 4}    # lfd_mft_desc.execution


def execution_code(execution):
    """lfd_mft_desc.execution for a name: None = the process default (lfd_set_mft_variant / LFD_MFT_VARIANT), else the K2a
    execution this call runs with ('direct' | 'folded' | 'czt' | 'auto'), independent of other threads and streams."""
    try:
        return EXECUTIONS[execution]
    except KeyError:
        raise ValueError(f"execution must be one of {sorted(k for k in EXECUTIONS if k)} or None, not {execution!r}")


def mft_descriptor(desc, f_dev, out_dev, alpha, shift, offset, unitary=True, inverse=False, execution=None):
    """Fill one lfd_mft_desc for device arrays `f_dev` (m x n) -> `out_dev` (M x N)."""
    ar, ac = _pair(alpha)
    sr, sc = _pair(shift)
    orow, ocol = _pair(offset)
    desc.f, desc.ldf = f_dev.data_ptr(), device.ld_of(f_dev)
    desc.out, desc.ldo = out_dev.data_ptr(), device.ld_of(out_dev)
    desc.m, desc.n = int(f_dev.shape[0]), int(f_dev.shape[1])
    desc.M, desc.N = int(out_dev.shape[0]), int(out_dev.shape[1])
    desc.alpha_r, desc.alpha_c = float(ar), float(ac)
    desc.shift_r, desc.shift_c = float(sr), float(sc)
    desc.off_r, desc.off_c = float(orow), float(ocol)
    desc.unitary, desc.inverse = int(bool(unitary)), int(bool(inverse))
    desc.execution = execution_code(execution)
    return desc


# bench.py sets this to a list to get (start event, end event, algorithmic flops) per batched launch
TIMERS = None


def mft_flops(descs, count):
    """Algorithmic flops of a batch, 8*M*n*(m+N) per plane."""
    return float(sum(8.0 * descs[i].M * descs[i].n * (descs[i].m + descs[i].N) for i in range(count)))


def mft_flops_executed(descs, count):
    """Real FP64 flops a batch executes (unpadded) under the execution the library picks for it: the folded form runs
    two real x complex GEMMs of ceil(M/2) x ceil(K/2) per stage, the direct one a complex x complex GEMM of M x K, the
    chirp-z form two FFTs of length L (5 L log2 L each) and three point-wise complex products per row transform."""
    execution = _lib.lib().lfd_mft_execution(descs, count)
    tot = 0.0
    for i in range(count):
        d = descs[i]
        if execution == 2:
            for rows, nin, nout in ((d.m, d.n, d.N), (d.N, d.m, d.M)):
                lg = 6
                while (1 << lg) < nin + nout - 1:
                    lg += 1
                tot += rows * (2 * 5.0 * (1 << lg) * lg + 6.0 * ((1 << lg) + nin + nout))
        elif execution == 1:
            h = lambda v: (v + 1) // 2
            tot += 8.0 * d.n * h(d.M) * h(d.m) + 8.0 * d.M * h(d.N) * h(d.n)
        else:
            tot += 8.0 * d.M * d.n * (d.m + d.N)
    return tot


def run_mft(descs, count, precision='c128', pupil_src=None, intensity_out=False):
    """Launch a batch of planes on the current stream with a torch-owned workspace.
    precision 'c128': K2a (FP64 DMMA); 'c64': K2b (complex64 arrays, 3xTF32 on tcgen05).
    pupil_src: optional lfd_pupil_src table — the fused K1+K2 entry points (folded K2a, or K2b)."""
    L = _lib.lib()
    if precision == 'c64':
        need = L.lfd_mft_c64x3_workspace_bytes(descs, count)
        ws = _workspace(need)
        if pupil_src is not None:
            _lib.check(L.lfd_mft_c64x3_from_pupil(descs, pupil_src, count, int(bool(intensity_out)), ws.data_ptr(), need,
                                                  device.stream_ptr()), "lfd_mft_c64x3_from_pupil")
        else:
            _lib.check(L.lfd_mft_c64x3_batched(descs, count, ws.data_ptr(), need, device.stream_ptr()),
                       "lfd_mft_c64x3_batched")
        return ws
    need = L.lfd_mft_workspace_bytes(descs, count)
    ws = _workspace(need)
    if TIMERS is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if pupil_src is not None:
        _lib.check(L.lfd_mft_c128_from_pupil(descs, pupil_src, count, int(bool(intensity_out)), ws.data_ptr(), need,
                                             device.stream_ptr()), "lfd_mft_c128_from_pupil")
    else:
        _lib.check(L.lfd_mft_c128_batched(descs, count, ws.data_ptr(), need, device.stream_ptr()),
                   "lfd_mft_c128_batched")
    if TIMERS is not None:
        e1.record()
        TIMERS.append((e0, e1, mft_flops(descs, count), mft_flops_executed(descs, count)))
    return ws


def dft2_dev(f_dev, alpha, shape=None, shift=(0, 0), offset=(0, 0), unitary=True, inverse=False,
             out=None, execution=None):
    """dft2 on device arrays; returns a complex128 device tensor (no host transfer)."""
    import torch
    m, n = int(f_dev.shape[0]), int(f_dev.shape[1])
    M, N = (m, n) if shape is None else (int(v) for v in _pair(shape))
    c64 = f_dev.dtype == torch.complex64
    if out is None:
        out = torch.empty(M, N, dtype=f_dev.dtype, device=f_dev.device)
    descs = (_lib.MftDesc * 1)()
    mft_descriptor(descs[0], f_dev, out, alpha, shift, offset, unitary, inverse, execution)
    run_mft(descs, 1, 'c64' if c64 else 'c128')
    return out


def _transform(f, alpha, shape, shift, offset, unitary, out, inverse, execution=None):
    if out is not None and not np.can_cast(complex, out.dtype):
        raise TypeError(f"Cannot cast complex output to dtype('{out.dtype}')")
    f = np.asarray(f)
    m, n = f.shape
    f_dev = device.to_dev(f, dtype=np.complex128)
    F = device.to_host(dft2_dev(f_dev, alpha, shape, shift, offset, unitary, inverse, execution=execution))
    if out is not None:
        out[...] = F
        return out
    return F


def dft2(f, alpha, shape=None, shift=(0, 0), offset=(0, 0), unitary=True, out=None, execution=None):
    """2-D discrete Fourier transform by matrix triple product (lentil/fourier.py:5-103).

    f : array_like (m, n); alpha : float or (row, col); shape : int or (M, N), default f.shape;
    shift : output-plane DC shift in pixels (r, c), may be fractional; offset : input-plane
    offset in pixels (r, c); unitary : scale by sqrt|alpha_r alpha_c|; out : optional result
    array (must accept complex, else TypeError; may alias f).  `execution` (not in the reference) picks the K2a execution
    for this call: None = library default, 'direct' | 'folded' (FP64 tensor cores) | 'czt' | 'auto'."""
    return _transform(f, alpha, shape, shift, offset, unitary, out, inverse=False, execution=execution)


def idft2(F, alpha, shape=None, shift=(0, 0), unitary=True, out=None, execution=None):
    """Inverse transform, conj(dft2(conj F)) / F.size (lentil/fourier.py:124-198): the
    conjugations fold into the twiddle sign, the division into the output scale."""
    return _transform(F, alpha, shape, shift, (0, 0), unitary, out, inverse=True, execution=execution)


def dft2_c64(f, alpha, shape=None, shift=(0, 0), offset=(0, 0), unitary=True, execution=None):
    """dft2 with complex64 input/output (K2b).  Not part of the reference surface (lentil is complex128
    throughout): the optional fast mode of the north star.  execution 'folded' = the 3xTF32 tensor-core
    (tcgen05) form, peak-normalised error ~1e-6 (growing with the input size); 'czt' / 'auto' / None =
    the FP32 chirp-z form wherever the planes fit it, ~3e-7 at any size."""
    f_dev = device.to_dev(np.asarray(f), dtype=np.complex64)
    return device.to_host(dft2_dev(f_dev, alpha, shape, shift, offset, unitary, inverse=False, execution=execution))


def idft2_c64(F, alpha, shape=None, shift=(0, 0), unitary=True, execution=None):
    f_dev = device.to_dev(np.asarray(F), dtype=np.complex64)
    return device.to_host(dft2_dev(f_dev, alpha, shape, shift, (0, 0), unitary, inverse=True, execution=execution))
