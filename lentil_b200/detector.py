"""Detector-side sampling of an oversampled PSF on the device: ``rebin`` (lentil/util.py:221-258) and
``pixel`` (lentil/detector.py:167-220), the two steps that follow the PSF in lentil's examples.

``pixel`` is the convolution with the square-pixel aperture done in the Fourier domain,
``|ifft2(fft2(img) * outer(sinc(fx*os), sinc(fy*os)))|``.  The transforms are K2a launches: dft2 with
alpha = 1/n is the centred FFT (lentil/fourier.py notes), so the MTF is built on the centred
frequency grid instead of np.fft.fftfreq's wrapped one — same numbers, different order.
Inputs may be numpy arrays (numpy out) or float64 device tensors (device out)."""
import numpy as np

from . import _lib, device
from . import fourier as _fourier


def _to_dev(img):
    return (img, True) if device.is_dev(img) else (device.to_dev(np.asarray(img), dtype=np.float64), False)


def rebin(img, factor):
    """Rebin an image (or a cube of images along axis 0) by an integer factor (lentil/util.py:206-258).
    Like the reference (whose numpy reshape needs exact division) a shape that `factor` does not divide is a
    ValueError; host input keeps its dtype on the way out, device tensors come back as float64."""
    if np.iscomplexobj(img) if not device.is_dev(img) else img.is_complex():
        raise ValueError('rebin is not defined for complex data')
    in_dtype = None
    if not device.is_dev(img):
        a = np.asarray(img)
        # the reference sums numpy blocks: a cube keeps its dtype (util.py:250), a single image takes numpy's sum dtype
        in_dtype = a.dtype if a.ndim == 3 else np.zeros(1, dtype=a.dtype).sum().dtype
    d, on_dev = _to_dev(img)
    planes = d.reshape((-1,) + tuple(d.shape[-2:]))
    h, w = int(planes.shape[1]), int(planes.shape[2])
    if h % int(factor) or w % int(factor):
        raise ValueError(f'cannot reshape array of size {h * w} into shape '
                         f'({h // factor},{factor},{w // factor},{factor})')
    out = device.zeros_f64(planes.shape[0], h // factor, w // factor)
    L = _lib.lib()
    for k in range(planes.shape[0]):
        _lib.check(L.lfd_rebin(planes[k].data_ptr(), device.ld_of(planes[k]), h, w, int(factor), out[k].data_ptr(),
                               device.stream_ptr()), "lfd_rebin")
    out = out.reshape(tuple(d.shape[:-2]) + (h // factor, w // factor))
    if on_dev:
        return out
    res = device.to_host(out)
    return res if in_dtype is None or in_dtype == res.dtype else res.astype(in_dtype)


def pixel(img, oversample=1):
    """Apply the aperture (MTF) of a square pixel to a discretely sampled image that is
    ``oversample`` times finer than the pixel (lentil/detector.py:167-220)."""
    d, on_dev = _to_dev(img)
    h, w = int(d.shape[0]), int(d.shape[1])
    # centred frequency grids: dft2 puts DC at index n//2, fftfreq puts it at 0
    fy = (np.arange(h) - h // 2) / h
    fx = (np.arange(w) - w // 2) / w
    # the reference builds kernel = outer(sinc(x*os), sinc(y*os)) with x from shape[1] and y from shape[0]:
    # kernel[i, j] = sinc(fx_i * os) * sinc(fy_j * os), i.e. the axes are swapped for non-square images;
    # it only runs for square images in practice (the outer product must match img.shape)
    if h != w:
        raise ValueError('pixel needs a square image (lentil.detector.pixel broadcasts only for square input)')
    my = device.to_dev(np.sinc(fx * oversample), dtype=np.float64)
    mx = device.to_dev(np.sinc(fy * oversample), dtype=np.float64)
    import torch
    f = torch.complex(d, torch.zeros_like(d))
    F = _fourier.dft2_dev(f, (1.0 / h, 1.0 / w), unitary=False)
    L = _lib.lib()
    _lib.check(L.lfd_scale_separable(F.data_ptr(), device.ld_of(F), h, w, my.data_ptr(), mx.data_ptr(),
                                     device.stream_ptr()), "lfd_scale_separable")
    g = _fourier.dft2_dev(F, (1.0 / h, 1.0 / w), unitary=False, inverse=True)
    out = device.zeros_f64(h, w)
    _lib.check(L.lfd_abs_c128(g.data_ptr(), device.ld_of(g), h, w, out.data_ptr(), device.stream_ptr()), "lfd_abs_c128")
    return out if on_dev else device.to_host(out)


# ---- spline rescale (lentil/util.py:261-347) and pixelate (lentil/detector.py:223-249) -----------------

_SPLINE_NPAD = 12            # map_coordinates pads 'nearest' inputs by 12 edge samples before the prefilter
_REFLECT_MODES = ('nearest', 'reflect')          # prefilter boundary: half-sample symmetric
_MIRROR_MODES = ('mirror', 'constant', 'wrap')   # whole-sample symmetric (scipy's legacy 'wrap' included)


def _bspline_weights(order, x):
    """Tap start and B-spline weights, vectorised over the coordinate vector x: (start[n], w[n, order+1])."""
    if order % 2:
        fl = np.floor(x)
    else:
        fl = np.floor(x + 0.5)
    start = fl.astype(np.int64) - order // 2
    y = x - fl
    if order == 0:
        w = np.ones((x.size, 1))
    elif order == 1:
        w = np.stack([1.0 - y, y], axis=1)
    elif order == 2:
        w1 = 0.75 - y * y
        t = 0.5 - y
        w0 = 0.5 * t * t
        w = np.stack([w0, w1, 1.0 - w0 - w1], axis=1)
    elif order == 3:
        z = 1.0 - y
        w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0
        w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
        w0 = z * z * z / 6.0
        w = np.stack([w0, w1, w2, 1.0 - w0 - w1 - w2], axis=1)
    else:
        from math import comb, factorial
        t = x[:, None] - (start[:, None] + np.arange(order + 1)[None, :])      # distance to each tap
        w = np.zeros_like(t)
        for k in range(order + 2):
            u = np.maximum(t + (order + 1) / 2.0 - k, 0.0)
            w += (-1) ** k * comb(order + 1, k) * u ** order
        w /= factorial(order)
    return start, w


def _extend_coordinate(c, n, mode):
    """map_coordinates' coordinate extension for the modes that are not pre-padded; returns (coords, valid)."""
    c = c.copy()
    valid = np.ones(c.shape, dtype=bool)
    if mode == 'nearest':
        return np.clip(c, 0.0, n - 1.0), valid
    if mode == 'constant':
        valid = (c >= 0) & (c <= n - 1)
        return np.where(valid, c, 0.0), valid
    if n <= 1:
        return np.zeros_like(c), valid
    lo, hi = c < 0, c > n - 1
    if mode == 'mirror':
        s2 = 2 * n - 2
        a = s2 * np.trunc(-c / s2) + c
        neg = np.where(a <= 1 - n, a + s2, -a)
        b = c - s2 * np.trunc(c / s2)
        pos = np.where(b > n - 1, s2 - b, b)
    elif mode == 'reflect':
        s2 = 2 * n
        a = np.where(c < -s2, s2 * np.trunc(-c / s2) + c, c)
        neg = np.where(a < -n, a + s2, -a - 1)
        b = c - s2 * np.trunc(c / s2)
        pos = np.where(b >= n, s2 - b - 1, b)
    elif mode == 'wrap':
        sz = n - 1
        neg = c + sz * (np.trunc(-c / sz) + 1)
        pos = c - sz * np.trunc(c / sz)
    else:
        raise ValueError(f"mode {mode!r} is not supported (constant, nearest, reflect, mirror, wrap)")
    return np.where(lo, neg, np.where(hi, pos, c)), valid


def _spline_taps(coords, n, order, mode):
    """(npad, idx int32 [len, order+1], weights float64 [len, order+1]) for one axis."""
    coords = np.asarray(coords, dtype=np.float64)
    taps = np.arange(order + 1)[None, :]
    if mode == 'nearest' and order > 1:
        npad = _SPLINE_NPAD
        start, w = _bspline_weights(order, coords + npad)
        idx = np.clip(start[:, None] + taps, 0, n + 2 * npad - 1)
        return npad, idx.astype(np.int32), w
    cm, valid = _extend_coordinate(coords, n, mode)
    start, w = _bspline_weights(order, cm)
    idx = start[:, None] + taps
    if mode in _REFLECT_MODES:
        s2 = 2 * n
        idx = np.mod(idx, s2)
        idx = np.where(idx >= n, s2 - 1 - idx, idx)
    elif n > 1:
        s2 = 2 * n - 2
        idx = np.mod(np.abs(idx), s2)
        idx = np.where(idx >= n, s2 - idx, idx)
    else:
        idx = np.zeros_like(idx)
    w = np.where(valid[:, None], w, 0.0)            # outside in 'constant' mode: cval = 0
    return 0, idx.astype(np.int32), w


def _map_separable(src, y, x, order, mode, nonzero=False):
    """map_coordinates(src, meshgrid(x, y)[::-1], order, mode) on the device for a real 2-D tensor."""
    L = _lib.lib()
    h, w = int(src.shape[0]), int(src.shape[1])
    npad, iy, wy = _spline_taps(y, h, order, mode)
    _, ix, wx = _spline_taps(x, w, order, mode)
    if order > 1 or npad:
        coef = device.zeros_f64(h + 2 * npad, w + 2 * npad)
        scratch = device.zeros_f64(h + 2 * npad, w + 2 * npad)
        _lib.check(L.lfd_spline_prefilter(src.data_ptr(), device.ld_of(src), h, w, npad, int(order),
                                          1 if mode in _REFLECT_MODES else 0, coef.data_ptr(), scratch.data_ptr(),
                                          device.stream_ptr()), "lfd_spline_prefilter")
    else:
        coef = src
    iyd, wyd = device.to_dev(iy), device.to_dev(wy)
    ixd, wxd = device.to_dev(ix), device.to_dev(wx)
    out = device.zeros_f64(len(y), len(x))
    _lib.check(L.lfd_spline_eval(coef.data_ptr(), device.ld_of(coef), 1 if nonzero else 0, iyd.data_ptr(), wyd.data_ptr(),
                                 len(y), ixd.data_ptr(), wxd.data_ptr(), len(x), order + 1, out.data_ptr(),
                                 device.stream_ptr()), "lfd_spline_eval")
    return out


def _sum_into(x, sums, slot, partials):
    _lib.check(_lib.lib().lfd_sum_f64(x.data_ptr(), x.numel(), partials.data_ptr(), sums[slot:].data_ptr(),
                                      device.stream_ptr()), "lfd_sum_f64")


def rescale(img, scale, shape=None, mask=None, order=3, mode='nearest', unitary=True):
    """Rescale an image by spline interpolation (lentil/util.py:261-347): same arguments, defaults and
    normalisation as ``lentil.rescale``.  The B-spline prefilter, the interpolation, the two sums of the
    unitary normalisation and the mask product run on the device; the per-row / per-column tap tables are
    host vectors.  numpy in -> numpy out, device tensor in -> device tensor out."""
    if order < 0 or order > 5:
        raise RuntimeError('spline order not supported')         # scipy.ndimage.map_coordinates' own error
    import torch
    if device.is_dev(img):
        d, on_dev = img, True
    else:
        a = np.asarray(img)
        d = device.to_dev(a, dtype=np.complex128 if np.iscomplexobj(a) else np.float64)
        on_dev = False
    if d.dim() != 2:
        raise ValueError('rescale needs a 2-D image')
    is_complex = d.is_complex()
    h, w = int(d.shape[0]), int(d.shape[1])
    if shape is None:
        shape = np.ceil((h * scale, w * scale)).astype(int)
    elif np.isscalar(shape):
        shape = np.ceil((shape * scale, shape * scale)).astype(int)
    else:
        shape = np.ceil((shape[0] * scale, shape[1] * scale)).astype(int)
    x = (np.arange(shape[1], dtype=np.float64) - shape[1] / 2.) / scale + w / 2.
    y = (np.arange(shape[0], dtype=np.float64) - shape[0] / 2.) / scale + h / 2.

    # dense planes: the sums below (lfd_sum_f64) read numel() consecutive doubles, so a strided view must be packed first
    planes = [d.real.contiguous(), d.imag.contiguous()] if is_complex else [d.double().contiguous()]
    if mask is None:
        # mask = (img != 0) interpolated bilinearly (util.py:315-319, 334); for complex data a pixel is non-zero when
        # either part is: form |re| + |im| once, the kernel reads it as a 0/1 map
        src = planes[0] if not is_complex else planes[0].abs() + planes[1].abs()
        m = _map_separable(src, y, x, 1, 'nearest', nonzero=True)
    else:
        md = mask if device.is_dev(mask) else device.to_dev(np.asarray(mask), dtype=np.float64)
        m = _map_separable(md.double().contiguous(), y, x, 1, 'nearest')
    outs = [_map_separable(p, y, x, order, mode) for p in planes]

    sums = None
    if unitary:
        sums = device.zeros_f64(4)
        partials = device.zeros_f64(256)
        for k, p in enumerate(planes):
            _sum_into(p, sums, k, partials)
        for k, o in enumerate(outs):
            _sum_into(o, sums, 2 + k, partials)
    n = int(shape[0]) * int(shape[1])
    if is_complex:
        out = device.empty_c128(int(shape[0]), int(shape[1]))
    else:
        out = device.zeros_f64(int(shape[0]), int(shape[1]))
    _lib.check(_lib.lib().lfd_rescale_finish(outs[0].data_ptr(), outs[1].data_ptr() if is_complex else None, m.data_ptr(),
                                             sums.data_ptr() if sums is not None else None, float(np.finfo(np.float64).eps),
                                             n, out.data_ptr(), device.stream_ptr()), "lfd_rescale_finish")
    return out if on_dev else device.to_host(out)


def pixelate(img, oversample):
    """Convolve an image with the pixel MTF and rescale the result to native sampling
    (lentil/detector.py:223-249); the intermediate stays on the device."""
    on_dev = device.is_dev(img)
    d, _ = _to_dev(img)
    out = rescale(pixel(d, oversample), 1 / oversample, order=3, mode='nearest', unitary=True)
    return out if on_dev else device.to_host(out)


def synthesize_opd(basis, coeffs, base=None, return_device=True):
    """OPD maps from a modal basis: ``out[r] = base + sum_k coeffs[r, k] * basis[k]`` — the
    ``np.einsum('ijk,i->jk', basis, coeff)`` of lentil's wavefront-error guide
    (docs/user/wavefront_error.rst:118-135) for R coefficient vectors at once, on the device, so a
    Monte-Carlo run never uploads R full OPD maps.  basis: (K, n, n); coeffs: (R, K) or (K,).
    The result (R, n, n) can be passed as ``opds=`` to ``propagate_dft_batch``."""
    b, _ = _to_dev(basis)
    c = np.atleast_2d(np.asarray(coeffs, dtype=np.float64))
    K, n0, n1 = (int(v) for v in b.shape)
    if c.shape[1] != K:
        raise ValueError(f'coeffs must have {K} columns')
    R = c.shape[0]
    cd = device.to_dev(c)
    base_d = None if base is None else _to_dev(base)[0]
    out = device.zeros_f64(R, n0, n1)
    L = _lib.lib()
    for k0 in range(0, K, 64):            # lfd_opd_synth takes at most 64 basis terms per call
        kk = min(64, K - k0)
        _lib.check(L.lfd_opd_synth(b[k0:k0 + kk].data_ptr(), cd[:, k0:k0 + kk].contiguous().data_ptr(),
                                   base_d.data_ptr() if (base_d is not None and k0 == 0) else None,
                                   n0 * n1, kk, R, 1 if k0 else 0, out.data_ptr(), device.stream_ptr()), "lfd_opd_synth")
    if np.ndim(coeffs) == 1:
        out = out[0]
    return out if return_device else device.to_host(out)
