"""CPU oracle for the far-field diffraction hot path of andykee/lentil v0.8.8.

TEST INFRASTRUCTURE ONLY.  This module is a from-scratch numpy restatement of the reference
algorithm; it exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs can check and time the reference behaviour on a box where
/root/reference does not exist.  Nothing in lentil_b200/ (the product) may import it.

Parity status: PINNED.  `oracle/make_golden.py` imports the real reference (run in the build
container, where /root/reference is mounted), checks every function below against it on seeded
inputs and writes the small golden fixtures under tests/golden/; `tests/test_oracle_golden.py`
re-checks this module against those fixtures on every run.

The innermost arithmetic of the reference lives in numpy (np.exp, np.outer, np.dot -> BLAS
zgemm; declared `numpy>=1.17`, unpinned, lentil pyproject.toml:14-17); the restatement calls
the same numpy primitives in the same order, so it is the reference's arithmetic to rounding.

Data model (functional, no classes): a *field* is a dict
    {"data": complex128 ndarray (2-D, or 0-d for the default planar field),
     "offset": (r, c) ints, "tilt": [(tx, ty), ...]}
where each tilt entry holds the already-swapped attributes of lentil.Tilt (see `tilt_entry`).
"""
import sys
from itertools import combinations

import numpy as np

# --------------------------------------------------------------------------------------------
# lentil/fourier.py
# --------------------------------------------------------------------------------------------


def dft2_coords(m, n, M, N):
    """Centred coordinate vectors, lentil/fourier.py:113-121."""
    R = np.arange(m) - np.floor(m / 2.0)
    S = np.arange(n) - np.floor(n / 2.0)
    U = np.arange(M) - np.floor(M / 2.0)
    V = np.arange(N) - np.floor(N / 2.0)
    return R, S, U, V


def dft2_matrices(m, n, M, N, ar, ac, sr, sc, orow, ocol):
    """The two DFT matrices, lentil/fourier.py:106-110 (E1 is the transposed outer product)."""
    R, S, U, V = dft2_coords(m, n, M, N)
    E1 = np.exp(-2.0 * 1j * np.pi * ar * np.outer(R + orow, U - sr)).T
    E2 = np.exp(-2.0 * 1j * np.pi * ac * np.outer(S + ocol, V - sc))
    return E1, E2


def dft2(f, alpha, shape=None, shift=(0, 0), offset=(0, 0), unitary=True):
    """Matrix-triple-product DFT, lentil/fourier.py:5-103 (argument handling :79-89, GEMMs :97,
    unitary scale :100-101)."""
    ar, ac = np.broadcast_to(alpha, (2,))
    f = np.asarray(f)
    m, n = f.shape
    if shape is None:
        shape = (m, n)
    M, N = np.broadcast_to(shape, (2,))
    sr, sc = np.broadcast_to(shift, (2,))
    orow, ocol = np.broadcast_to(offset, (2,))
    E1, E2 = dft2_matrices(m, n, int(M), int(N), ar, ac, sr, sc, orow, ocol)
    F = np.dot(E1.dot(f), E2)
    if unitary:
        F = F * np.sqrt(np.abs(ar * ac))
    return F


def idft2(F, alpha, shape=None, shift=(0, 0), unitary=True):
    """conj(dft2(conj F)) / F.size, lentil/fourier.py:193-198 (note: input size, no offset)."""
    F = np.asarray(F)
    out = dft2(np.conj(F), alpha, shape, shift, unitary=unitary)
    return np.conj(out) / F.size


# --------------------------------------------------------------------------------------------
# lentil/extent.py  (inclusive integer rectangles (rmin, rmax, cmin, cmax))
# --------------------------------------------------------------------------------------------


def array_extent(shape, shift):
    """lentil/extent.py:5-40 (int() truncation toward zero at :28-29; 0-d shapes act as 1x1)."""
    if len(shape) < 2:
        shape = (1, 1)
    rmin = int(-(shape[0] // 2) + shift[0])
    cmin = int(-(shape[1] // 2) + shift[1])
    return rmin, int(rmin + shape[0] - 1), cmin, int(cmin + shape[1] - 1)


def array_center(extent):
    """lentil/extent.py:43-59."""
    rmin, rmax, cmin, cmax = extent
    return rmin + (rmax - rmin + 1) // 2, cmin + (cmax - cmin + 1) // 2


def intersect(a, b):
    """lentil/extent.py:62-77."""
    return a[0] <= b[1] and a[1] >= b[0] and a[2] <= b[3] and a[3] >= b[2]


def intersection_extent(a, b):
    """lentil/extent.py:80-100."""
    return max(a[0], b[0]), min(a[1], b[1]), max(a[2], b[2]), min(a[3], b[3])


def intersection_shape(a, b):
    """lentil/extent.py:103-124."""
    rmin, rmax, cmin, cmax = intersection_extent(a, b)
    nr, nc = rmax - rmin + 1, cmax - cmin + 1
    return () if (nr <= 0 or nc <= 0) else (nr, nc)


def intersection_slices(a, b):
    """lentil/extent.py:127-150."""
    rmin, rmax, cmin, cmax = intersection_extent(a, b)
    sa = (slice(rmin - a[0], rmax - a[0] + 1), slice(cmin - a[2], cmax - a[2] + 1))
    sb = (slice(rmin - b[0], rmax - b[0] + 1), slice(cmin - b[2], cmax - b[2] + 1))
    return sa, sb


def intersection_shift(a, b):
    """lentil/extent.py:153-169."""
    return array_center(intersection_extent(a, b))


# --------------------------------------------------------------------------------------------
# lentil/util.py:190-218, lentil/helper.py:27-123
# --------------------------------------------------------------------------------------------


def boundary(x, threshold=0):
    """Bounding row/col indices of x > threshold, lentil/util.py:190-218."""
    x = np.asarray(x) > threshold
    rows = np.flatnonzero(np.any(x, axis=1))
    cols = np.flatnonzero(np.any(x, axis=0))
    return rows[0], rows[-1], cols[0], cols[-1]


def boundary_slice(x, threshold=0):
    """lentil/helper.py:27-62 with pad=(0,0)."""
    rmin, rmax, cmin, cmax = boundary(x, threshold)
    return (slice(max(rmin, 0), min(rmax + 1, x.shape[0])),
            slice(max(cmin, 0), min(cmax + 1, x.shape[1])))


def slice_offset(slc, shape):
    """Offset of a slice centre from the array centre, lentil/helper.py:65-123."""
    if slc is Ellipsis:
        return (0, 0)
    h = slc[0].stop - slc[0].start
    w = slc[1].stop - slc[1].start
    return (int(slc[0].start + h // 2 - shape[0] // 2), int(slc[1].start + w // 2 - shape[1] // 2))


# --------------------------------------------------------------------------------------------
# lentil/field.py
# --------------------------------------------------------------------------------------------


def make_field(data, offset=None, tilt=None):
    """lentil/field.py:33-56 (data forced to complex128 at :35)."""
    return {"data": np.asarray(data, dtype=complex),
            "offset": tuple(int(v) for v in (offset if offset is not None else (0, 0))),
            "tilt": list(tilt) if tilt else []}


def field_extent(f):
    return array_extent(f["data"].shape, f["offset"])


def field_mul(a, b):
    """Field product with rectangle intersection, lentil/field.py:80-147 and :464-484.
    Returns None when the result is empty."""
    tilt = a["tilt"] + b["tilt"]
    ad, ao, bd, bo = a["data"], a["offset"], b["data"], b["offset"]
    if ad.size == 1 and bd.size == 1:                       # :118-128
        if tuple(ao) != tuple(bo):
            return None
        return make_field(ad * bd, ao, tilt)
    if ad.shape != bd.shape:                                # :464-484
        if ad.size == 1:
            ad, ao = np.broadcast_to(ad, bd.shape), bo
        if bd.size == 1:
            bd, bo = np.broadcast_to(bd, ad.shape), ao
    ea, eb = array_extent(ad.shape, ao), array_extent(bd.shape, bo)
    if not intersect(ea, eb):
        return None
    sa, sb = intersection_slices(ea, eb)
    return make_field(ad[sa] * bd[sb], intersection_shift(ea, eb), tilt)


def tilt_entry(x, y):
    """lentil.Tilt(x, y) stores self.x = y, self.y = x (lentil/plane.py:898-901)."""
    return (y, x)


def field_shift(f, z, pixelscale, oversample):
    """Tilt -> (row, col) pixel shift, lentil/field.py:149-194 with indexing='ij';
    each tilt applies lentil/plane.py:903-923: x = xs - z*self.x, y = ys - z*self.y."""
    x, y = 0, 0
    for tx, ty in f["tilt"]:
        x, y = x - z * tx, y - z * ty
    ps = np.broadcast_to(pixelscale, (2,))
    px, py = x / ps[0] * oversample, y / ps[1] * oversample
    return -py, px


def fields_boundary(fields):
    """lentil/field.py:196-229 (rmax/cmax seeded with 0 at :220 — reproduced)."""
    rmin, rmax, cmin, cmax = sys.maxsize, 0, sys.maxsize, 0
    for f in fields:
        e = field_extent(f)
        rmin, rmax = min(rmin, e[0]), max(rmax, e[1])
        cmin, cmax = min(cmin, e[2]), max(cmax, e[3])
    return rmin, rmax, cmin, cmax


def insert(field, out, intensity=False, weight=1):
    """Place a field in a dense array, lentil/field.py:231-305."""
    data = field["data"]
    fs, os_ = np.asarray(data.shape), np.asarray(out.shape)
    if data.shape == out.shape and tuple(field["offset"]) == (0, 0):
        osl = fsl = Ellipsis
    else:
        ul = os_ // 2 - fs // 2 + np.asarray(field["offset"])
        fr0, fr1, fc0, fc1 = 0, int(fs[0]), 0, int(fs[1])
        or0, or1 = int(ul[0]), int(ul[0] + fs[0])
        oc0, oc1 = int(ul[1]), int(ul[1] + fs[1])
        if or0 < 0:
            fr0, or0 = -or0, 0
        if or1 > os_[0]:
            fr1, or1 = fr1 - (or1 - os_[0]), int(os_[0])
        if oc0 < 0:
            fc0, oc0 = -oc0, 0
        if oc1 > os_[1]:
            fc1, oc1 = fc1 - (oc1 - os_[1]), int(os_[1])
        osl = (slice(or0, or1), slice(oc0, oc1))
        fsl = (slice(fr0, fr1), slice(fc0, fc1))
    if intensity:
        out[osl] += np.abs(data[fsl] ** 2) * weight      # :301-302 (|z^2|, not re^2+im^2)
    else:
        out[osl] += data[fsl] * weight
    return out


def merge_fields(fields):
    """Coherent sum into the common bounding box, lentil/field.py:331-386."""
    rmin, rmax, cmin, cmax = fields_boundary(fields)
    out = np.zeros((rmax - rmin + 1, cmax - cmin + 1), dtype=complex)
    for f in fields:
        e = field_extent(f)
        out[e[0] - rmin:e[1] - rmin + 1, e[2] - cmin:e[3] - cmin + 1] += f["data"]
    nrow, ncol = rmax - rmin + 1, cmax - cmin + 1
    return make_field(out, (rmin + nrow // 2, cmin + ncol // 2))


def reduce_fields(fields):
    """Disjoint grouping by transitive overlap then merge, lentil/field.py:413-461."""
    groups = [{"field": [f], "extent": field_extent(f)} for f in fields]

    def disjoint(gs):
        for a, b in combinations(range(len(gs)), 2):
            if intersect(gs[a]["extent"], gs[b]["extent"]):
                gs[a]["field"].extend(gs[b]["field"])
                gs[a]["extent"] = fields_boundary(gs[a]["field"])
                gs.pop(b)
                return disjoint(gs)
        return gs

    out = []
    for grp in disjoint(groups):
        out.append(merge_fields(grp["field"]) if len(grp["field"]) > 1 else grp["field"][0])
    return out


def wavefront_intensity(fields, shape):
    """lentil/wavefront.py:114-125."""
    out = np.zeros(shape, dtype=float)
    for f in reduce_fields(fields):
        out = insert(f, out, intensity=True)
    return out


def wavefront_insert(fields, out, weight=1):
    """lentil/wavefront.py:145-165."""
    for f in reduce_fields(fields):
        out = insert(f, out, intensity=True, weight=weight)
    return out


def wavefront_field(fields, shape):
    """lentil/wavefront.py:101-112."""
    out = np.zeros(shape, dtype=complex)
    for f in fields:
        out = insert(f, out)
    return out


# --------------------------------------------------------------------------------------------
# lentil/plane.py:477-516 — Plane.__mul__ for a wavefront
# --------------------------------------------------------------------------------------------


def plane_slices(mask):
    """lentil/plane.py:671-705."""
    mask = np.asarray(mask)
    if mask.ndim < 2:
        return [Ellipsis]
    if mask.ndim == 2:
        return [boundary_slice(mask)]
    return [boundary_slice(m) for m in mask]


def plane_multiply(fields, amplitude, opd, mask, wavelength, plane_tilt=None):
    """Multiply wavefront fields by a plane, lentil/plane.py:494-514.

    `mask` None reproduces plane.py:43-47 (mask = amplitude.astype(bool)); `plane_tilt` is the
    per-segment tilt list fit_tilt leaves on the plane (entries as `tilt_entry`)."""
    amplitude, opd = np.asarray(amplitude), np.asarray(opd)
    if mask is None:
        mask = np.copy(amplitude).astype(bool)
    mask = np.asarray(mask)
    nseg = mask.shape[0] if mask.ndim == 3 else 1
    shape = mask.shape if nseg == 1 else mask.shape[1:]
    out = []
    for field in fields:
        for n, s in enumerate(plane_slices(mask)):
            mk = mask if nseg == 1 else mask[n]
            amp = amplitude if amplitude.size == 1 else amplitude[s] * mk[s]
            od = opd if opd.size == 1 else opd[s]
            phasor = make_field(amp * np.exp(2 * np.pi * 1j * od / wavelength),
                                slice_offset(s, shape),
                                [plane_tilt[n]] if plane_tilt else [])
            res = field_mul(field, phasor)
            if res is not None and res["data"].size > 0:
                out.append(res)
    return out


# --------------------------------------------------------------------------------------------
# lentil/propagate.py:147-260 — propagate_dft
# --------------------------------------------------------------------------------------------


def dft_alpha(dx, du, wavelength, z, oversample):
    """lentil/propagate.py:121-123."""
    return ((dx[0] * du[0]) / (wavelength * z * oversample),
            (dx[1] * du[1]) / (wavelength * z * oversample))


def mask_window(mask):
    """Output-window shape and shift from a detector mask, lentil/propagate.py:245-260."""
    rmin, rmax, cmin, cmax = boundary(mask, 0)
    shape = (rmax - rmin + 1, cmax - cmin + 1)
    shift = (rmin + shape[0] // 2 - mask.shape[0] // 2, cmin + shape[1] // 2 - mask.shape[1] // 2)
    return shape, shift


def plan_window(shift, prop_shape_out, out_extent):
    """Integer window logic of lentil/propagate.py:211-230 for one field.

    Returns None when the propagation window misses the output, else
    (intersect_shape, intersect_shift, dft_shift) with dft_shift = prop_shift + sub-pixel shift."""
    shift = np.asarray(shift, dtype=float)
    fix_shift = np.fix(shift)
    subpx = shift - fix_shift
    prop_extent = array_extent(prop_shape_out, fix_shift)
    if not intersect(out_extent, prop_extent):
        return None
    ishape = intersection_shape(out_extent, prop_extent)
    ishift = intersection_shift(out_extent, prop_extent)
    iextent = array_extent(ishape, ishift)
    prop_shift = np.array(array_center(prop_extent)) - np.array(array_center(iextent))
    return ishape, ishift, prop_shift + subpx


def propagate_dft(fields, wavelength, dx, z, pixelscale, shape, prop_shape=None, oversample=2,
                  mask=None):
    """lentil/propagate.py:147-242 on a list of fields.  `dx` = wavefront pixelscale (2,),
    `z` = focal length.  Returns (list of output fields, shape_out)."""
    shape = np.broadcast_to(shape, (2,))
    prop_shape = np.asarray(shape) if prop_shape is None else np.broadcast_to(prop_shape, (2,))
    shape_out = shape * oversample
    prop_shape_out = prop_shape * oversample
    if mask is not None:
        mask = np.asarray(mask)
        if np.all(mask.shape != shape_out):                  # :184 (quirk: raises only if BOTH differ)
            raise ValueError("shape mismatch: mask shape != output shape")
        mshape, mshift = mask_window(mask)
        out_extent = array_extent(mshape, mshift)
    else:
        out_extent = array_extent(shape_out, (0, 0))
    dx = np.broadcast_to(dx, (2,))
    du = np.broadcast_to(pixelscale, (2,))
    out = []
    for field in fields:
        shift = field_shift(field, z, du, oversample)
        plan = plan_window(shift, prop_shape_out, out_extent)
        if plan is None:
            continue
        ishape, ishift, dshift = plan
        alpha = dft_alpha(dx, du, wavelength, z, oversample)
        data = dft2(field["data"], alpha, shape=ishape, shift=dshift, offset=field["offset"],
                    unitary=True)
        out.append(make_field(data, ishift))
    return out, tuple(int(v) for v in shape_out)


def psf(amplitude, opd, mask, wavelengths, weights, dx, z, pixelscale, shape, prop_shape=None,
        oversample=2, wf_tilt=None, plane_tilt=None, out_mask=None):
    """The canonical user loop (docs/user/diffraction.rst:154-158, performance.rst:44-49):
    for each wavelength  Wavefront(wl[, tilt]) * pupil -> propagate_dft -> insert(img, weight)."""
    shape_out = tuple(int(v) for v in np.broadcast_to(shape, (2,)) * oversample)
    img = np.zeros(shape_out, dtype=float)
    for wl, wt in zip(wavelengths, weights):
        w0 = [make_field(np.array(1, dtype=complex), None,
                         [tilt_entry(*wf_tilt)] if wf_tilt is not None else None)]
        w1 = plane_multiply(w0, amplitude, opd, mask, wl, plane_tilt)
        w2, _ = propagate_dft(w1, wl, dx, z, pixelscale, shape, prop_shape, oversample, out_mask)
        img = wavefront_insert(w2, img, wt)
    return img


# --------------------------------------------------------------------------------------------
# the FFT sibling of propagate_dft (next row): lentil/propagate.py:9-136, lentil/util.py:31-90
# --------------------------------------------------------------------------------------------


def pad(array, shape):
    """Centred zero-pad (or centre crop where the array is larger), lentil/util.py:31-90, 2-D case."""
    array = np.asarray(array)
    src, dst = [], []
    for have, want in zip(array.shape, shape):
        if want - have <= 0:
            lo = (have - want) // 2
            src.append(slice(lo, lo + want))
            dst.append(slice(0, want))
        else:
            lo = (want - have) // 2
            src.append(slice(0, have))
            dst.append(slice(lo, lo + have))
    padded = np.zeros((shape[0], shape[1]), dtype=array.dtype)
    padded[tuple(dst)] = array[tuple(src)]
    return padded


def fft_shape(dx, du, z, wavelength, oversample):
    """Padded size that realises the requested sampling, and the wavelength that the integer
    padding really corresponds to, lentil/propagate.py:126-132."""
    alpha = dft_alpha(dx, du, z, wavelength, oversample)     # z / wavelength swapped as in :129 (a product)
    npix = np.round(np.reciprocal(alpha)).astype(int)
    prop_wavelength = np.min((npix / oversample * dx * du) / z)
    return npix, prop_wavelength


def propagate_fft(fields, wf_shape, wavelength, dx, z, pixelscale, shape=None, oversample=2):
    """lentil/propagate.py:9-88 (no-scratch branch): dense field -> centred pad -> shifted
    orthonormal FFT (:135-136).  Returns (output field, shape_out, propagation wavelength)."""
    if any(f["tilt"] for f in fields):
        raise NotImplementedError('propagate_fft does not support Wavefronts with fitted tilt.')
    dx = np.broadcast_to(dx, (2,))
    du = np.broadcast_to(pixelscale, (2,))
    npix, prop_wavelength = fft_shape(dx, du, z, wavelength, oversample)
    if shape is None:
        shape_out = tuple(int(v) for v in npix)
    else:
        shape = tuple(np.broadcast_to(shape, (2,)))
        if np.any(shape > npix / oversample):
            raise ValueError('requested shape is larger than the maximum propagation shape')
        shape_out = (int(shape[0] * oversample), int(shape[1] * oversample))
    x = pad(wavefront_field(fields, wf_shape), npix)
    F = np.fft.ifftshift(np.fft.fft2(np.fft.fftshift(x), norm='ortho'))
    return make_field(F, (0, 0)), shape_out, prop_wavelength


# --------------------------------------------------------------------------------------------
# detector-side sampling (next rows): lentil/util.py:221-258, lentil/detector.py:167-220
# --------------------------------------------------------------------------------------------


def rebin(img, factor):
    """Integer-factor binning by reshape + sum, lentil/util.py:221-258."""
    img = np.asarray(img)
    if np.iscomplexobj(img):
        raise ValueError('rebin is not defined for complex data')
    if img.ndim == 3:
        return np.stack([rebin(plane, factor) for plane in img])
    return img.reshape(img.shape[0] // factor, factor, img.shape[1] // factor, factor).sum(-1).sum(1)


def pixel(img, oversample=1):
    """Square-pixel MTF applied in the Fourier domain, lentil/detector.py:213-220."""
    img = np.asarray(img)
    mtf_x = np.sinc(np.fft.fftfreq(img.shape[1]) * oversample)
    mtf_y = np.sinc(np.fft.fftfreq(img.shape[0]) * oversample)
    kernel = np.dot(mtf_x[:, np.newaxis], mtf_y[np.newaxis, :])
    return np.abs(np.fft.ifft2(np.fft.fft2(img) * kernel))


# --------------------------------------------------------------------------------------------
# spline rescale to native sampling (next row): lentil/util.py:261-347 (rescale),
# lentil/detector.py:223-249 (pixelate).  The interpolation itself lives in a third-party dependency
# that is not vendored in the reference: scipy.ndimage.map_coordinates (scipy is unpinned in
# pyproject.toml; checked here against scipy 1.18.1).  What follows restates its published algorithm
# (Unser's recursive B-spline prefilter with exact boundary initialisation, then separable B-spline
# evaluation) for the call sites of util.py:334-343, where the coordinates are a meshgrid of one x
# vector and one y vector.
# --------------------------------------------------------------------------------------------

SPLINE_POLES = {
    2: (np.sqrt(8.0) - 3.0,),
    3: (np.sqrt(3.0) - 2.0,),
    4: (np.sqrt(664.0 - np.sqrt(438976.0)) + np.sqrt(304.0) - 19.0,
        np.sqrt(664.0 + np.sqrt(438976.0)) - np.sqrt(304.0) - 19.0),
    5: (np.sqrt(67.5 - np.sqrt(4436.25)) + np.sqrt(26.25) - 6.5,
        np.sqrt(67.5 + np.sqrt(4436.25)) - np.sqrt(26.25) - 6.5),
}
# boundary condition of the prefilter for each map_coordinates mode: 'nearest' runs on an input that
# was padded by 12 edge samples and is then filtered with the half-sample-symmetric ("reflect")
# initialisation; 'constant' and the legacy 'wrap' use the whole-sample-symmetric ("mirror") one.
SPLINE_BOUNDARY = {'nearest': 'reflect', 'reflect': 'reflect', 'mirror': 'mirror', 'constant': 'mirror',
                   'wrap': 'mirror'}
SPLINE_NPAD = 12


def spline_filter_axis0(a, order, boundary):
    """B-spline coefficients of every column of `a` (lines along axis 0), in place semantics on a copy."""
    c = np.array(a, dtype=np.float64)
    n = c.shape[0]
    if order < 2 or n < 2:
        return c
    poles = SPLINE_POLES[order]
    gain = 1.0
    for z in poles:
        gain *= (1.0 - z) * (1.0 - 1.0 / z)
    c *= gain
    for z in poles:
        src = c.copy()
        if boundary == 'mirror':
            zn1 = z ** (n - 1)
            acc = src[0] + zn1 * src[n - 1]
            z_i = z
            for i in range(1, n - 1):
                acc = acc + z_i * (src[i] + zn1 * src[n - 1 - i])
                z_i *= z
            c[0] = acc / (1.0 - zn1 * zn1)
        else:
            zn = z ** n
            acc = src[0] + zn * src[n - 1]
            z_i = z
            for i in range(1, n):
                acc = acc + z_i * (src[i] + zn * src[n - 1 - i])
                z_i *= z
            c[0] = acc * (z / (1.0 - zn * zn)) + src[0]
        for i in range(1, n):
            c[i] = c[i] + z * c[i - 1]
        if boundary == 'mirror':
            c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1.0)
        else:
            c[n - 1] = c[n - 1] * (z / (z - 1.0))
        for i in range(n - 2, -1, -1):
            c[i] = z * (c[i + 1] - c[i])
    return c


def spline_weights(order, x):
    """First tap and the order+1 B-spline weights for coordinate x (a float)."""
    if order % 2:
        start = int(np.floor(x)) - order // 2
    else:
        start = int(np.floor(x + 0.5)) - order // 2
    if order == 0:
        return start, np.array([1.0])
    if order == 1:
        y = x - np.floor(x)
        return start, np.array([1.0 - y, y])
    if order == 2:
        y = x - np.floor(x + 0.5)
        w1 = 0.75 - y * y
        t = 0.5 - y
        w0 = 0.5 * t * t
        return start, np.array([w0, w1, 1.0 - w0 - w1])
    if order == 3:
        y = x - np.floor(x)
        z = 1.0 - y
        w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0
        w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
        w0 = z * z * z / 6.0
        return start, np.array([w0, w1, w2, 1.0 - w0 - w1 - w2])
    # orders 4, 5: the centred cardinal B-spline, beta^n(t) = 1/n! sum_k (-1)^k C(n+1,k) (t + (n+1)/2 - k)_+^n
    from math import comb, factorial
    w = np.zeros(order + 1)
    for j in range(order + 1):
        t = x - (start + j)
        s = 0.0
        for k in range(order + 2):
            u = t + (order + 1) / 2.0 - k
            if u > 0:
                s += (-1) ** k * comb(order + 1, k) * u ** order
        w[j] = s / factorial(order)
    return start, w


def _map_coordinate(c, n, mode):
    """Coordinate extension of map_coordinates for modes without pre-padding; None = outside ('constant')."""
    if mode == 'nearest':
        return min(max(c, 0.0), n - 1.0)
    if mode == 'constant':
        return None if (c < 0 or c > n - 1) else c
    if n <= 1:
        return 0.0
    if mode == 'mirror':
        s2 = 2 * n - 2
        if c < 0:
            c = s2 * int(-c / s2) + c
            return c + s2 if c <= 1 - n else -c
        if c > n - 1:
            c -= s2 * int(c / s2)
            if c > n - 1:
                c = s2 - c
        return c
    if mode == 'reflect':
        s2 = 2 * n
        if c < 0:
            if c < -s2:
                c = s2 * int(-c / s2) + c
            return c + s2 if c < -n else -c - 1
        if c > n - 1:
            c -= s2 * int(c / s2)
            if c >= n:
                c = s2 - c - 1
        return c
    if mode == 'wrap':      # scipy's legacy 'wrap': period n - 1
        sz = n - 1
        if c < 0:
            return c + sz * (int(-c / sz) + 1)
        if c > n - 1:
            return c - sz * int(c / sz)
        return c
    raise ValueError(f'unsupported mode {mode!r}')


def _map_tap(i, n, boundary):
    if 0 <= i < n:
        return i
    if boundary == 'mirror':
        if n <= 1:
            return 0
        s2 = 2 * n - 2
        i = abs(i) % s2
        return s2 - i if i >= n else i
    s2 = 2 * n
    i = i % s2
    return s2 - 1 - i if i >= n else i


def spline_taps(coords, n, order, mode):
    """Tap indices (into the prefiltered line of length n + 2*npad) and weights for every coordinate of a
    1-D coordinate vector: (npad, idx[len, order+1] int64, w[len, order+1] float64).  A coordinate outside
    the array in 'constant' mode gets zero weights (cval = 0, as util.py:334-343 never passes cval)."""
    coords = np.asarray(coords, dtype=np.float64)
    npad = SPLINE_NPAD if (mode == 'nearest' and order > 1) else 0
    N = n + 2 * npad
    idx = np.zeros((coords.size, order + 1), dtype=np.int64)
    w = np.zeros((coords.size, order + 1), dtype=np.float64)
    for k, c in enumerate(coords):
        if npad:
            # padded 'nearest': the coordinate is used as is, taps that leave the padded line are clamped
            st, wk = spline_weights(order, float(c) + npad)
            idx[k] = np.clip(st + np.arange(order + 1), 0, N - 1)
            w[k] = wk
            continue
        cm = _map_coordinate(float(c), n, mode)
        if cm is None:
            continue
        st, wk = spline_weights(order, cm)
        idx[k] = [_map_tap(st + j, N, SPLINE_BOUNDARY[mode]) for j in range(order + 1)]
        w[k] = wk
    return npad, idx, w


def map_coordinates_separable(img, y, x, order, mode):
    """scipy.ndimage.map_coordinates(img, meshgrid(x, y)[::-1], order, mode) for real `img`."""
    img = np.asarray(img, dtype=np.float64)
    npad_y, iy, wy = spline_taps(y, img.shape[0], order, mode)
    npad_x, ix, wx = spline_taps(x, img.shape[1], order, mode)
    c = np.pad(img, npad_y, mode='edge') if npad_y else img
    if order > 1:
        b = SPLINE_BOUNDARY[mode]
        c = spline_filter_axis0(c, order, b)                 # axis 0 first, then axis 1, as spline_filter does
        c = spline_filter_axis0(c.T, order, b).T
    out = np.zeros((len(y), len(x)))
    for a in range(order + 1):                               # same tap order as the C loop: rows outer, columns inner
        for bb in range(order + 1):
            out = out + c[iy[:, a][:, None], ix[:, bb][None, :]] * wy[:, a][:, None] * wx[:, bb][None, :]
    return out


def rescale(img, scale, shape=None, mask=None, order=3, mode='nearest', unitary=True):
    """lentil/util.py:261-347."""
    img = np.asarray(img)
    if mask is None:
        mask = np.zeros_like(img).real
        mask[img != 0] = 1
    if shape is None:
        shape = np.ceil((img.shape[0] * scale, img.shape[1] * scale)).astype(int)
    elif np.isscalar(shape):
        shape = np.ceil((shape * scale, shape * scale)).astype(int)
    else:
        shape = np.ceil((shape[0] * scale, shape[1] * scale)).astype(int)
    x = (np.arange(shape[1], dtype=np.float64) - shape[1] / 2.) / scale + img.shape[1] / 2.
    y = (np.arange(shape[0], dtype=np.float64) - shape[0] / 2.) / scale + img.shape[0] / 2.
    mask = map_coordinates_separable(mask, y, x, 1, 'nearest')
    mask[mask < np.finfo(mask.dtype).eps] = 0
    if np.iscomplexobj(img):
        out = np.zeros(shape, dtype=np.complex128)
        out.real = map_coordinates_separable(img.real, y, x, order, mode)
        out.imag = map_coordinates_separable(img.imag, y, x, order, mode)
    else:
        out = map_coordinates_separable(img, y, x, order, mode)
    if unitary:
        out *= np.sum(img) / np.sum(out)
    out *= mask
    return out


def pixelate(img, oversample):
    """lentil/detector.py:223-249."""
    return rescale(pixel(img, oversample), 1 / oversample, order=3, mode='nearest', unitary=True)


# --------------------------------------------------------------------------------------------
# wavefront-error generator of BASELINE config 5 (next row): lentil/wfe.py:8-70
# --------------------------------------------------------------------------------------------


def power_spectrum(mask, pixelscale, rms, half_power_freq, exp, seed=None):
    """PSD-filtered random OPD, lentil/wfe.py:40-70 (same numpy calls in the same order)."""
    mask = np.asarray(mask)
    rng = np.random.default_rng(seed)
    n, m = mask.shape
    yy, xx = np.mgrid[0:m, 0:n]
    yy = (yy - (np.floor(m / 2) + 1)) / m
    xx = (xx - (np.floor(n / 2) + 1)) / n
    dr = np.sqrt(xx * xx + yy * yy)
    half_power_freq = half_power_freq * pixelscale / np.sqrt(m ** 2 + n ** 2)
    psd = 1 / (1 + (dr / half_power_freq) ** exp)
    psd[dr == 0] = 0
    psd = psd / np.sum(psd)
    H = np.fft.fftshift(np.sqrt(psd))
    noise = rng.normal(size=[n, m])
    opd = np.real(np.fft.ifft2(np.fft.fft2(noise) * H)) * np.sqrt(m * n)
    opd *= mask
    opd = opd * np.sqrt(np.count_nonzero(opd) / np.sum(np.abs(opd) ** 2)) * rms
    return opd
