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
    """Rebin an image (or a cube of images along axis 0) by an integer factor; trailing rows/columns
    that do not fill a block are dropped, as numpy's reshape in the reference requires exact division."""
    if np.iscomplexobj(img) if not device.is_dev(img) else img.is_complex():
        raise ValueError('rebin is not defined for complex data')
    d, on_dev = _to_dev(img)
    planes = d.reshape((-1,) + tuple(d.shape[-2:]))
    h, w = int(planes.shape[1]), int(planes.shape[2])
    out = device.zeros_f64(planes.shape[0], h // factor, w // factor)
    L = _lib.lib()
    for k in range(planes.shape[0]):
        _lib.check(L.lfd_rebin(planes[k].data_ptr(), device.ld_of(planes[k]), h, w, int(factor), out[k].data_ptr(),
                               device.stream_ptr()), "lfd_rebin")
    out = out.reshape(tuple(d.shape[:-2]) + (h // factor, w // factor))
    return out if on_dev else device.to_host(out)


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
