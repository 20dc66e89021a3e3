"""Far-field propagation drivers.

`propagate_dft` keeps the call surface of lentil/propagate.py:147-242: per Field it turns the
accumulated tilt into a pixel shift, splits it into an integer window move and a sub-pixel DFT
shift, clips the propagation window against the output (or a detector mask's bounding box) and
transforms the field onto exactly that window.  All Fields of the wavefront go to the device in
ONE batched K2a launch pair (lfd_mft_c128_batched).

`propagate_dft_batch` is the loop the reference leaves to user code
(docs/user/diffraction.rst:154-158) — wavelengths x field points x segments — planned on the
host and executed as K1 -> K2a -> K3 over whole chunks of planes, with the pupil resident in
HBM and the PSF stack accumulated on the device.
"""
import ctypes as C
import math

import numpy as np

from . import _lib, device, extent as _extent, helper
from . import field as _field
from . import fourier as _fourier
import importlib
_pt = importlib.import_module(".ptype", __package__)  # the package attribute `ptype` is the factory function
from .field import Field
from .wavefront import Wavefront


def _dft_alpha(dx, du, wavelength, z, oversample):
    """Output sampling in cycles per input pixel per output pixel (lentil/propagate.py:121-123)."""
    return ((dx[0] * du[0]) / (wavelength * z * oversample),
            (dx[1] * du[1]) / (wavelength * z * oversample))


def _propagate_ptype(ptype, method='fraunhofer'):
    """pupil <-> image (lentil/propagate.py:263-272)."""
    if method == 'fraunhofer':
        if ptype not in (_pt.pupil, _pt.image):
            raise TypeError("Wavefront must have ptype 'pupil' or 'image'")
        return _pt.image if ptype == _pt.pupil else _pt.pupil


def _mask_shape(x, threshold=0):
    rmin, rmax, cmin, cmax = helper.boundary(x, threshold)
    return rmax - rmin + 1, cmax - cmin + 1


def _mask_shift(x, threshold=0):
    """Shift of the masked area's centre from the array centre (lentil/propagate.py:251-260)."""
    rmin, rmax, cmin, cmax = helper.boundary(x, threshold)
    return (rmin + (rmax - rmin + 1) // 2 - x.shape[0] // 2,
            cmin + (cmax - cmin + 1) // 2 - x.shape[1] // 2)


def _output_extent(shape_out, mask):
    """Extent of the region that is actually computed (lentil/propagate.py:182-190)."""
    if mask is None:
        return _extent.array_extent(shape_out, shift=(0, 0))
    mask = np.asarray(mask)
    if np.all(mask.shape != shape_out):   # same (lenient) test as the reference, propagate.py:184
        raise ValueError(f'shape mismatch: mask shape {mask.shape} != output shape {tuple(shape_out)}')
    return _extent.array_extent(_mask_shape(mask), _mask_shift(mask))


def plan_window(shift, prop_shape_out, out_extent):
    """Integer window logic for one Field (lentil/propagate.py:211-230).

    shift : (r, c) float pixel shift of the field centre in the output plane.
    Returns None when the propagation window misses the output region, else
    ``(intersect_shape, intersect_shift, dft_shift)``: the window to compute, the offset of the
    resulting Field, and the shift handed to dft2 (= prop_shift + sub-pixel remainder).
    Plain Python arithmetic (this runs once per Field and wavelength in the drop-in loop);
    math.trunc is np.fix for finite floats."""
    sr, sc = float(shift[0]), float(shift[1])
    fr, fc = float(math.trunc(sr)), float(math.trunc(sc))
    prop_extent = _extent.array_extent(prop_shape_out, (fr, fc))
    if not _extent.intersect(out_extent, prop_extent):
        return None
    intersect_shape = _extent.intersection_shape(out_extent, prop_extent)
    intersect_shift = _extent.intersection_shift(out_extent, prop_extent)
    intersect_extent = _extent.array_extent(intersect_shape, intersect_shift)
    pc, ic = _extent.array_center(prop_extent), _extent.array_center(intersect_extent)
    return intersect_shape, intersect_shift, np.array([pc[0] - ic[0] + (sr - fr), pc[1] - ic[1] + (sc - fc)])


def propagate_dft(wavefront, pixelscale, shape=None, prop_shape=None, oversample=2, mask=None):
    """Propagate a Wavefront in the far field with the matrix DFT (lentil/propagate.py:147-242).

    Same parameters, return value and errors as the reference: TypeError unless the wavefront
    ptype is pupil or image, ValueError on a mask of the wrong shape; returns a NEW Wavefront
    whose `data` holds one Field per input Field that lands on the output (possibly none)."""
    ptype_out = _propagate_ptype(wavefront.ptype, method='fraunhofer')

    shape = np.asarray(wavefront.shape) if shape is None else helper.pair(shape)
    prop_shape = np.asarray(shape) if prop_shape is None else helper.pair(prop_shape)
    shape_out = shape * oversample
    prop_shape_out = prop_shape * oversample
    out_extent = _output_extent(shape_out, mask)

    dx = wavefront.pixelscale
    du = helper.pair(pixelscale)
    z = wavefront.focal_length

    out = Wavefront.empty(wavelength=wavefront.wavelength, pixelscale=du / oversample,
                          focal_length=wavefront.focal_length, shape=shape_out, ptype=ptype_out)

    plans = []
    for field in wavefront.data:
        shift = field.shift(z=z, wavelength=wavefront.wavelength, pixelscale=du,
                            oversample=oversample, indexing='ij')
        plan = plan_window(shift, prop_shape_out, out_extent)
        if plan is not None:
            plans.append((field, plan))
    if not plans:
        return out

    alpha = _dft_alpha(dx=dx, du=du, z=z, wavelength=wavefront.wavelength, oversample=oversample)
    total = sum(int(p[0][0]) * int(p[0][1]) for _, p in plans)
    buf = device.empty_c128(total)
    descs = (_lib.MftDesc * len(plans))()
    pos = 0
    for k, (field, (ishape, ishift, dshift)) in enumerate(plans):
        h, w = int(ishape[0]), int(ishape[1])
        win = buf[pos:pos + h * w].view(h, w)
        pos += h * w
        _fourier.mft_descriptor(descs[k], field.dev, win, alpha, dshift, field.offset, unitary=True)
        out.data.append(Field(data=win, pixelscale=du / oversample, offset=ishift))
    _fourier.run_mft(descs, len(plans))
    return out


def _fft_shape(dx, du, z, wavelength, oversample):
    """Padded size whose FFT bin spacing realises the requested sampling, and the wavelength the
    integer padding really corresponds to (lentil/propagate.py:126-132)."""
    alpha = _dft_alpha(dx, du, wavelength, z, oversample)
    fft_shape = np.round(np.reciprocal(alpha)).astype(int)
    prop_wavelength = np.min((fft_shape / oversample * dx * du) / z)
    return fft_shape, prop_wavelength


def scratch_shape(wavelength, dx, du, z, oversample):
    """Scratch shape an FFT propagation of the longest wavelength needs (lentil/propagate.py:91-118)."""
    dx, du = np.broadcast_to(dx, (2,)), np.broadcast_to(du, (2,))
    fft_shape, _ = _fft_shape(dx, du, z, np.max(wavelength), oversample)
    return tuple(fft_shape)


def _padded_fft_geometry(have, npix):
    """Rows (or columns) of a `have`-long dense field that survive lentil.util.pad to `npix`
    (util.py:31-90: centred pad, centre crop when larger), and the dft2 offset/shift that make
    the matrix transform of just those rows equal to ``ifftshift(fft(fftshift(padded)))``.

    With h = npix // 2 the shifted FFT is  sum_j x[j] e^{-2 pi i (j+h)(k+h)/npix}; in centred
    coordinates R = j - h, U = k - h that is e^{-2 pi i RU/npix} for even npix (2h = npix) and
    e^{-2 pi i (R-1)(U-1)/npix} for odd npix (2h = npix - 1): offset -1, shift +1."""
    if npix - have <= 0:
        src0, keep, dst0 = (have - npix) // 2, npix, 0
    else:
        src0, keep, dst0 = 0, have, (npix - have) // 2
    odd = npix % 2
    offset = dst0 + keep // 2 - npix // 2 - odd
    return src0, keep, offset, odd


def propagate_fft(wavefront, pixelscale, shape=None, oversample=2, scratch=None):
    """Far-field propagation on the FFT's sampling grid (lentil/propagate.py:9-88).

    Same parameters, errors and result as the reference: the dense wavefront field is zero-padded
    to ``round(1/alpha)`` samples, transformed with an orthonormal shifted FFT, and returned as ONE
    Field of the padded size inside a Wavefront of ``shape * oversample`` whose wavelength is the
    one the integer padding really samples.  NotImplementedError for fields that carry tilt,
    ValueError when `shape` exceeds the padded size or `scratch` is too small.

    On the device the padded array never exists: the transform of the non-zero rows and columns is
    a K2a launch with alpha = 1/npix (the same sum the FFT evaluates, term for term), so `scratch`
    is validated and otherwise unused."""
    if any(f.tilt for f in wavefront.data):
        raise NotImplementedError('propagate_fft does not support Wavefronts with fitted tilt. '
                                  'Use propagate_dft instead.')
    ptype_out = _propagate_ptype(wavefront.ptype, method='fraunhofer')
    du = np.broadcast_to(pixelscale, (2,))
    fft_shape, prop_wavelength = _fft_shape(wavefront.pixelscale, du, wavefront.focal_length,
                                            wavefront.wavelength, oversample)
    if shape is None:
        shape_out = tuple(int(v) for v in fft_shape)
    else:
        shape = tuple(np.broadcast_to(shape, (2,)))
        if np.any(shape > fft_shape / oversample):
            raise ValueError(f'requested shape {tuple(shape)} is larger in at least one dimension than '
                             f'maximum propagation shape {tuple(fft_shape // oversample)}')
        shape_out = (shape[0] * oversample, shape[1] * oversample)
    if scratch is not None and not all(np.asarray(scratch.shape) > fft_shape):
        raise ValueError(f'scratch must have shape greater than or equal to {tuple(fft_shape)}')

    out = Wavefront.empty(wavelength=prop_wavelength, pixelscale=du / oversample,
                          focal_length=wavefront.focal_length, shape=shape_out, ptype=ptype_out)
    dense = wavefront._accumulate(device.empty_c128(*wavefront.shape).zero_(), 1, False)   # K3
    r0, nr, off_r, odd_r = _padded_fft_geometry(int(dense.shape[0]), int(fft_shape[0]))
    c0, nc, off_c, odd_c = _padded_fft_geometry(int(dense.shape[1]), int(fft_shape[1]))
    F = _fourier.dft2_dev(dense[r0:r0 + nr, c0:c0 + nc], alpha=(1.0 / fft_shape[0], 1.0 / fft_shape[1]),
                          shape=(int(fft_shape[0]), int(fft_shape[1])), shift=(odd_r, odd_c),
                          offset=(off_r, off_c), unitary=True)                             # K2a
    out.data.append(Field(data=F, pixelscale=du / oversample))
    return out


# ----------------------------------------------------------------------------------------------
# batched driver
# ----------------------------------------------------------------------------------------------


def _shift_for(tilts, z, wavelength, du, oversample):
    """Field.shift for an explicit tilt list (lentil/field.py:183-194, indexing='ij')."""
    x, y = 0, 0
    for t in tilts:
        x, y = t.__shift__(xs=x, ys=y, z=z, wavelength=wavelength)
    return -(y / du[1] * oversample), x / du[0] * oversample


def propagate_dft_batch(plane, wavelengths, pixelscale, shape, prop_shape=None, oversample=2,
                        mask=None, weights=None, tilts=None, opds=None, out=None, chunk_bytes=8 << 30,
                        distributed=False, return_device=False, precision='c128', execution=None):
    """Polychromatic, multi-field-point, multi-realisation PSF stack in one call.

    Equivalent to the reference user loop (docs/user/performance.rst:44-49)::

        for r, opd in enumerate(opds):                # optional Monte-Carlo WFE realisations
            plane.opd = opd
            for p, tilt in enumerate(tilts):          # optional field points
                for wl, wt in zip(wavelengths, weights):
                    w = Wavefront(wl, tilt=tilt) * plane
                    w = propagate_dft(w, pixelscale, shape, prop_shape, oversample, mask)
                    img[r, p] = w.insert(img[r, p], wt)

    plane : Pupil (or Plane with ptype pupil/image and a focal length on the wavefront side)
    wavelengths : (L,) metres;  weights : (L,) default 1
    tilts : None (on-axis) or sequence of [rx, ry] field points -> axis of length P
    opds : None (use plane.opd) or array (R, n, n) of OPD maps (host or device) -> axis of length R
    out : optional float64 device tensor to accumulate into, shape ([R,] [P,] H, W)
    distributed : shard the wavelengths over torch.distributed ranks and all-reduce the stack
    return_device : return the device tensor instead of a numpy array
    precision : 'c128' (default: lentil's arithmetic, FP64 tensor cores) or 'c64' (complex64 fields
        through the 3xTF32 tcgen05 path K2b; intensities are still accumulated in float64;
        peak-normalised PSF error ~1e-6 .. 1e-5)
    execution : K2a execution of THIS call ('direct' | 'folded' = FP64 tensor cores | 'czt' | 'auto'); None = the library
        default (LFD_MFT_AUTO unless lfd_set_mft_variant / LFD_MFT_VARIANT changed it)

    Returns ([R,] [P,] H, W): axes that were not requested are squeezed away.

    Host work is O(field points x segments) window planning plus numpy-vectorised descriptor
    tables (one row per plane); everything per-pixel runs in K1 / K2a / K3.
    """
    import torch
    from .plane import Tilt
    if precision not in ('c128', 'c64'):
        raise ValueError("precision must be 'c128' or 'c64'")
    c64 = precision == 'c64'
    cdtype, esize = (torch.complex64, 8) if c64 else (torch.complex128, 16)
    wavelengths = np.asarray(wavelengths, dtype=float).reshape(-1)
    L = len(wavelengths)
    weights = np.ones(L) if weights is None else np.asarray(weights, dtype=float).reshape(-1)
    if len(weights) != L:
        raise ValueError('weights and wavelengths must have the same length')
    points = [None] if tilts is None else [Tilt(x=t[0], y=t[1]) for t in tilts]
    P = len(points)

    if plane.ptype not in (_pt.pupil, _pt.image):
        raise TypeError("Wavefront must have ptype 'pupil' or 'image'")
    ops = plane._operands()
    if ops['scalar'] is not None:
        raise ValueError('propagate_dft_batch needs a plane with array amplitude or opd')
    if opds is None:
        R, opd_stack = 1, None
    else:
        # the kernels read the stack as dense float64 on this device: normalise whatever the caller handed over
        opd_stack = (opds.to(device.device(), torch.float64).contiguous() if device.is_dev(opds)
                     else device.to_dev(np.asarray(opds), dtype=np.float64))
        if opd_stack.dim() != 3 or tuple(opd_stack.shape[1:]) != tuple(ops['shape']):
            raise ValueError(f"opds must have shape (R, {ops['shape'][0]}, {ops['shape'][1]})")
        R = int(opd_stack.shape[0])

    shape = np.broadcast_to(shape, (2,))
    prop_shape = np.asarray(shape) if prop_shape is None else np.broadcast_to(prop_shape, (2,))
    shape_out = shape * oversample
    prop_shape_out = prop_shape * oversample
    out_extent = _output_extent(shape_out, mask)
    H, W = int(shape_out[0]), int(shape_out[1])
    dx = np.broadcast_to(plane.pixelscale, (2,))
    du = np.broadcast_to(pixelscale, (2,))
    z = getattr(plane, 'focal_length', None)

    my = shard_indices(L, distributed)
    if out is not None:
        if not (device.is_dev(out) and out.dtype == torch.float64 and out.device == device.device()
                and out.is_contiguous() and out.numel() == R * P * H * W):
            raise ValueError(f'out must be a contiguous float64 tensor on {device.device()} with {R * P * H * W} '
                             f'elements ([R,] [P,] {H}, {W})')
    # a distributed call all-reduces what THIS call computed and only then adds it to `out` (whose previous contents
    # would otherwise be summed once per rank)
    reduce_into = out if (out is not None and distributed) else None
    stack = out if (out is not None and reduce_into is None) else device.zeros_f64(R * P, H, W)
    stack3 = stack.view(R * P, H, W)

    nseg = ops['nseg']
    segs = ops['segs']
    seg_tilts = [[plane.tilt[n]] if plane.tilt else [] for n in range(nseg)]
    pairs = [(p, n, ([points[p]] if points[p] is not None else []) + seg_tilts[n])
             for p in range(P) for n in range(nseg)]
    # plain Tilt objects shift by the same number of pixels at every wavelength: plan once
    static = all(type(t) is Tilt for _, _, tl in pairs for t in tl)
    # bytes per wavefront (one realisation at one wavelength): phasors + per field point the folded
    # intermediates and the output window
    per_wf = 16 * ops['total'] + P * sum(
        16 * (2 * int(segs[n].w) * int(prop_shape_out[0]) + int(prop_shape_out[0]) * int(prop_shape_out[1]))
        for n in range(nseg))
    step = max(1, int(chunk_bytes // max(per_wf, 1)))

    seg_off = np.array([segs[n].out_offset for n in range(nseg)], dtype=np.int64)
    seg_h = np.array([segs[n].h for n in range(nseg)], dtype=np.int64)
    seg_w = np.array([segs[n].w for n in range(nseg)], dtype=np.int64)
    offs = np.array(ops['offsets'], dtype=float).reshape(nseg, 2)
    alpha_num = (dx[0] * du[0], dx[1] * du[1])

    # wavefronts of this rank: realisation-major, wavelength-minor
    wf_r = np.repeat(np.arange(R, dtype=np.int64), len(my))
    wf_l = np.tile(np.asarray(my, dtype=np.int64), R)

    for c0 in range(0, len(wf_r), step):
        cr, cl = wf_r[c0:c0 + step], wf_l[c0:c0 + step]
        lam = wavelengths[cl]
        nw = len(lam)
        # With one field point every phasor feeds exactly one transform: K1 is then fused into the fold
        # kernel of the folded K2a / of K2b (lfd_mft_c128_from_pupil, lfd_mft_c64x3_from_pupil) and phasors never exist in HBM; with a single
        # segment the column stage also squares the field itself (no coherent merge to do in K3).
        exec_code = _fourier.execution_code(execution)
        direct = (exec_code == 1) if exec_code else (_lib.lib().lfd_get_mft_variant() == 0)
        fused = P == 1 and (c64 or not direct)                              # every execution but the direct one fuses K1
        intensity_out = fused and nseg == 1
        if not fused:
            phasors = torch.empty(nw, ops['total'], dtype=cdtype, device=device.device())
            for r in np.unique(cr):                                        # K1, one launch per realisation
                sel = np.flatnonzero(cr == r)
                plane._phasors_into(phasors[int(sel[0]):int(sel[-1]) + 1], lam[sel], ops,
                                    None if opd_stack is None else opd_stack[int(r)])
        # ---- window planning: rows (wavefront, p, n) -> window shape / offset / dft shift ----------
        if static:
            plans = [plan_window(_shift_for(tl, z, lam[0], du, oversample), prop_shape_out, out_extent)
                     for _, _, tl in pairs]
            keep = [k for k, pl in enumerate(plans) if pl is not None]
            jp = np.tile(np.array([pairs[k][0] for k in keep], dtype=np.int64), nw)
            jn = np.tile(np.array([pairs[k][1] for k in keep], dtype=np.int64), nw)
            jw = np.repeat(np.arange(nw, dtype=np.int64), len(keep))
            jshape = np.tile(np.array([plans[k][0] for k in keep], dtype=np.int64).reshape(-1, 2), (nw, 1))
            jshift = np.tile(np.array([plans[k][1] for k in keep], dtype=np.int64).reshape(-1, 2), (nw, 1))
            jdft = np.tile(np.array([plans[k][2] for k in keep], dtype=float).reshape(-1, 2), (nw, 1))
        else:
            rows = []
            for wi, wl in enumerate(lam):
                for p, n, tl in pairs:
                    pl = plan_window(_shift_for(tl, z, wl, du, oversample), prop_shape_out, out_extent)
                    if pl is not None:
                        rows.append((wi, p, n, pl))
            jw = np.array([r[0] for r in rows], dtype=np.int64)
            jp = np.array([r[1] for r in rows], dtype=np.int64)
            jn = np.array([r[2] for r in rows], dtype=np.int64)
            jshape = np.array([r[3][0] for r in rows], dtype=np.int64).reshape(-1, 2)
            jshift = np.array([r[3][1] for r in rows], dtype=np.int64).reshape(-1, 2)
            jdft = np.array([r[3][2] for r in rows], dtype=float).reshape(-1, 2)
        nj = len(jw)
        if nj == 0:
            continue
        sizes = jshape[:, 0] * jshape[:, 1]
        pos = np.concatenate(([0], np.cumsum(sizes)[:-1]))
        osize = 8 if intensity_out else esize
        buf = torch.empty(int(sizes.sum()), dtype=torch.float64 if intensity_out else cdtype, device=device.device())
        # ---- K2a / K2b descriptors, one row per plane -------------------------------------------------
        D = np.zeros(nj, dtype=np.dtype(_lib.MftDesc))
        src = None
        if fused:
            n_r, n_c = ops['shape']
            seg_r0 = np.array([segs[n].r0 for n in range(nseg)], dtype=np.int64)
            seg_c0 = np.array([segs[n].c0 for n in range(nseg)], dtype=np.int64)
            seg_mi = np.array([segs[n].mask_index for n in range(nseg)], dtype=np.int64)
            src = np.zeros(nj, dtype=np.dtype(_lib.PupilSrc))
            src['amp'] = ops['amp'].data_ptr()
            src['opd'] = (ops['opd'].data_ptr() if opd_stack is None
                          else opd_stack.data_ptr() + 8 * n_r * n_c * cr[jw])
            src['mask'] = 0 if ops['mask'] is None else ops['mask'].data_ptr() + n_r * n_c * seg_mi[jn]
            src['n_r'], src['n_c'] = n_r, n_c
            src['r0'], src['c0'] = seg_r0[jn], seg_c0[jn]
            src['wavelength'] = lam[jw]
        else:
            D['f'] = phasors.data_ptr() + esize * (jw * ops['total'] + seg_off[jn])
        D['ldf'] = seg_w[jn]
        D['out'] = buf.data_ptr() + osize * pos
        D['ldo'] = jshape[:, 1]
        D['m'], D['n'] = seg_h[jn], seg_w[jn]
        D['M'], D['N'] = jshape[:, 0], jshape[:, 1]
        D['alpha_r'] = alpha_num[0] / (lam[jw] * z * oversample)
        D['alpha_c'] = alpha_num[1] / (lam[jw] * z * oversample)
        D['shift_r'], D['shift_c'] = jdft[:, 0], jdft[:, 1]
        D['off_r'], D['off_c'] = offs[jn, 0], offs[jn, 1]
        D['unitary'] = 1
        D['execution'] = exec_code
        for b0 in range(0, nj, _MAX_PLANES_PER_LAUNCH):                      # K2a
            nb = min(_MAX_PLANES_PER_LAUNCH, nj - b0)
            _fourier.run_mft(D[b0:b0 + nb].ctypes.data_as(C.POINTER(_lib.MftDesc)), nb, precision,
                             None if src is None else src[b0:b0 + nb].ctypes.data_as(C.POINTER(_lib.PupilSrc)),
                             intensity_out)
        # ---- K3: per output image (realisation, field point); groups = wavefronts --------------------
        Wn = np.zeros(nj, dtype=np.dtype(_lib.Window))
        Wn['E'], Wn['ld'] = D['out'], jshape[:, 1]
        Wn['h'], Wn['w'] = jshape[:, 0], jshape[:, 1]
        Wn['r0'] = H // 2 - jshape[:, 0] // 2 + jshift[:, 0]      # lentil/field.py:267-268
        Wn['c0'] = W // 2 - jshape[:, 1] // 2 + jshift[:, 1]
        Wn['group'] = jw
        Wn['c64'] = 2 if intensity_out else (1 if c64 else 0)
        Wn['weight'] = weights[cl][jw]
        img = cr[jw] * P + jp
        for im in np.unique(img):
            sel = Wn[img == im] if (R * P) > 1 else Wn
            if len(sel):
                _field.accumulate_windows(np.ascontiguousarray(sel), stack3[int(im)])

    if distributed:
        reduce_stack(stack)
    if reduce_into is not None:
        reduce_into.view(R * P, H, W).add_(stack3)
        stack3 = reduce_into.view(R * P, H, W)
    result = stack3.view(R, P, H, W)
    if tilts is None:
        result = result[:, 0]
    if opds is None:
        result = result[0]
    return result if return_device else device.to_host(result.contiguous())


_MAX_PLANES_PER_LAUNCH = 16384      # lfd_mft_c128_batched takes at most 32767 planes per call


def shard_indices(L, distributed=False, rank=None, world=None):
    """Wavelength indices this rank owns: a strided deal of range(L) over the ranks of
    torch.distributed (all of them when not distributed).  Strided so that every rank sees the
    whole spectral range (window sizes, and so the work per plane, can vary with wavelength)."""
    if not distributed:
        return np.arange(L)
    if rank is None or world is None:
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
    return np.arange(L)[rank::world]


def reduce_stack(stack):
    """Sum the per-rank PSF stacks in place (NCCL all-reduce over NVLink on GPUs; gloo in the CPU
    tests).  The only collective on the path: planes are independent until the incoherent sum."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stack, op=dist.ReduceOp.SUM)
    return stack
