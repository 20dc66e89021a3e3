"""Wavefront-error generators that feed the hot path (lentil/wfe.py).

``power_spectrum`` — PSD-filtered random OPD (lentil/wfe.py:8-70) — is the OPD source of BASELINE config 5.  The random
draw and the filter definition stay on the host (the numpy generator *is* the specification of the noise; the filter is
O(n^2) scalar math that must reproduce the reference's grid conventions bit for bit); the two FFTs of the
noise-shaping step run on the device as K2a transforms with alpha = 1/n.
"""
import numpy as np

from . import device
from . import fourier as _fourier
from .field import _dev_mul


def _psd_filter(shape, pixelscale, half_power_freq, exp):
    """sqrt(PSD) on the reference's frequency grid (lentil/wfe.py:43-58, before the fftshift): zero frequency sits at
    index floor(n/2)+1 (a 1-based centre) and the grid axes follow the reference's (m, n) naming."""
    n, m = shape
    yy, xx = np.mgrid[0:m, 0:n]
    yy = (yy - (np.floor(m / 2) + 1)) / m
    xx = (xx - (np.floor(n / 2) + 1)) / n
    dr = np.sqrt(xx * xx + yy * yy)
    half_power_freq = half_power_freq * pixelscale / np.sqrt(m ** 2 + n ** 2)
    psd = 1 / (1 + (dr / half_power_freq) ** exp)
    psd[dr == 0] = 0
    psd = psd / np.sum(psd)
    return np.sqrt(psd)


def power_spectrum(mask, pixelscale, rms, half_power_freq, exp, seed=None):
    """Wavefront error with an inverse-power-law PSD, masked and scaled to `rms` (lentil/wfe.py:8-70).

    Same parameters and random stream as the reference (``np.random.default_rng(seed).normal``).  The reference
    evaluates ``real(ifft2(fft2(noise) * fftshift(sqrt(psd))))``; here both transforms are centred matrix DFTs on the
    device, for which the same filter is applied in centred order (fftshift of the reference's H) and the centring
    phases of the forward and inverse transforms cancel exactly."""
    mask = np.asarray(mask)
    if mask.ndim != 2 or mask.shape[0] != mask.shape[1]:
        raise ValueError('power_spectrum needs a square 2-D mask (the reference filter only broadcasts for square input)')
    rng = np.random.default_rng(seed)
    root_psd = _psd_filter(mask.shape, pixelscale, half_power_freq, exp)
    H = np.fft.fftshift(root_psd)                       # the reference's filter, in unshifted FFT order
    n, m = mask.shape
    noise = rng.normal(size=[n, m])

    alpha = (1.0 / n, 1.0 / m)
    f = device.to_dev(noise, dtype=np.complex128)
    F = _fourier.dft2_dev(f, alpha, unitary=False)                                   # K2a: centred fft2
    G = _dev_mul(F, device.to_dev(np.fft.fftshift(H), dtype=np.complex128), 1.0)      # filter in centred order
    y = _fourier.dft2_dev(G, alpha, unitary=False, inverse=True)                     # K2a: centred ifft2 (1/(n m) inside)
    opd = np.real(device.to_host(y)) * np.sqrt(m * n)

    opd *= mask
    opd = opd * np.sqrt(np.count_nonzero(opd) / np.sum(np.abs(opd) ** 2)) * rms
    return opd
