"""Synthetic pupils and wavefront-error maps of the BASELINE.json shapes (host, numpy, setup time).

The reference's content generators (lentil/shape.py, zernike.py, segmented.py, wfe.py) are out
of scope; bench.py and the tests only need inputs of the right shape and
character, so these are small independent generators, not re-implementations: hard-edged
apertures, Noll-ordered Zernike polynomials from the textbook radial formula, a flat-to-flat
hexagon tiling.  All take a numpy Generator / seed where random.
"""
from math import factorial

import numpy as np


def _grid(shape, shift=(0, 0)):
    rr, cc = np.meshgrid(np.arange(shape[0]) - shape[0] // 2 - shift[0],
                         np.arange(shape[1]) - shape[1] // 2 - shift[1], indexing='ij')
    return rr.astype(float), cc.astype(float)


def circle(shape, radius, shift=(0, 0)):
    """Hard-edged disc centred on index n//2 (+shift); bbox is (2*radius+1)^2 for integer radius."""
    rr, cc = _grid(shape, shift)
    return (rr * rr + cc * cc <= radius * radius).astype(float)


def annulus(shape, radius, obscuration=1 / 3):
    return circle(shape, radius) - circle(shape, radius * obscuration)


def normalize_power(amp, power=1.0):
    """Scale so that sum |amp|^2 = power (the convention lentil.normalize_power establishes)."""
    amp = np.asarray(amp, dtype=float)
    return amp * np.sqrt(power / np.sum(amp * amp))


def _noll_to_nm(j):
    n = int(np.ceil((-3 + np.sqrt(9 + 8 * (j - 1))) / 2))
    k = j - n * (n + 1) // 2          # 1-based position inside the radial order
    if n % 2 == 0:
        m = 2 * (k // 2)
    else:
        m = 2 * ((k - 1) // 2) + 1
    if m != 0 and j % 2 == 1:
        m = -m
    return n, m


def _radial(n, m, rho):
    out = np.zeros_like(rho)
    for s in range((n - m) // 2 + 1):
        c = (-1) ** s * factorial(n - s) / (factorial(s) * factorial((n + m) // 2 - s)
                                            * factorial((n - m) // 2 - s))
        out += c * rho ** (n - 2 * s)
    return out


def zernike(j, rho, theta):
    """Noll Zernike polynomial j (1-based), unit RMS over the unit disc."""
    n, m = _noll_to_nm(j)
    if m == 0:
        return np.sqrt(n + 1) * _radial(n, 0, rho)
    if m > 0:
        return np.sqrt(2 * (n + 1)) * _radial(n, m, rho) * np.cos(m * theta)
    return np.sqrt(2 * (n + 1)) * _radial(n, -m, rho) * np.sin(-m * theta)


def zernike_opd(mask, coeffs, first=1):
    """OPD = sum_j coeffs[j] Z_{first+j} over the bounding circle of `mask` (zero outside it)."""
    mask = np.asarray(mask)
    rows, cols = np.nonzero(mask)
    r0, c0 = (rows.min() + rows.max()) / 2.0, (cols.min() + cols.max()) / 2.0
    rad = max(rows.max() - rows.min(), cols.max() - cols.min()) / 2.0 + 0.5
    rr, cc = np.meshgrid(np.arange(mask.shape[0]) - r0, np.arange(mask.shape[1]) - c0, indexing='ij')
    rho, theta = np.hypot(rr, cc) / rad, np.arctan2(rr, cc)
    opd = np.zeros(mask.shape)
    for k, c in enumerate(coeffs):
        if c != 0:
            opd += c * zernike(first + k, rho, theta)
    return opd * (mask != 0)


def hexagon_mask(shape, radius, center):
    """Flat-top hexagon of circumradius `radius` centred at (row, col) `center`."""
    rr, cc = np.meshgrid(np.arange(shape[0]) - center[0], np.arange(shape[1]) - center[1], indexing='ij')
    inradius = radius * np.sqrt(3) / 2
    inside = np.abs(rr) <= inradius
    for ang in (np.pi / 3, -np.pi / 3):
        inside &= np.abs(rr * np.cos(ang) + cc * np.sin(ang)) <= inradius
    return inside


def hex_segments(rings, seg_radius, seg_gap, drop_center=True):
    """(nseg, n, n) boolean cube of a hexagonal tiling; rings=2 without the centre = 18 segments
    (the JWST-like aperture of BASELINE config 3)."""
    pitch = seg_radius * np.sqrt(3) + seg_gap            # centre-to-centre distance
    centers = []
    for q in range(-rings, rings + 1):
        for r in range(-rings, rings + 1):
            if max(abs(q), abs(r), abs(q + r)) > rings or (drop_center and q == 0 and r == 0):
                continue
            # neighbours sit along the normals of the flat edges: the row axis and +-60 deg from it
            centers.append((pitch * (q + 0.5 * r), pitch * (np.sqrt(3) / 2) * r))
    ext = max(max(abs(a), abs(b)) for a, b in centers) + seg_radius + 2
    n = 2 * int(np.ceil(ext)) + 1
    cube = np.zeros((len(centers), n, n), dtype=bool)
    for k, (y, x) in enumerate(centers):
        cube[k] = hexagon_mask((n, n), seg_radius, (n // 2 + y, n // 2 + x))
    return cube


def power_law_opd(mask, rms, rng, exponent=3.0):
    """Random OPD with a power-law PSD ~ f^-exponent, scaled to `rms` over the mask."""
    rng = np.random.default_rng(rng)
    n0, n1 = mask.shape
    fr, fc = np.meshgrid(np.fft.fftfreq(n0), np.fft.fftfreq(n1), indexing='ij')
    f = np.hypot(fr, fc)
    f[0, 0] = np.inf
    spec = f ** (-exponent / 2.0) * np.exp(2j * np.pi * rng.random((n0, n1)))
    opd = np.real(np.fft.ifft2(spec))
    m = mask != 0
    opd = (opd - opd[m].mean()) * m
    return opd * (rms / np.sqrt(np.mean(opd[m] ** 2)))
