"""Pin the oracle to the real reference and write the golden fixtures.

Run in the BUILD container only (needs /root/reference, read-only):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

1. imports andykee/lentil v0.8.8 from /root/reference,
2. checks every function of oracle/lentil_oracle.py against it on seeded inputs (hard asserts),
3. writes small input/output vectors produced BY THE REFERENCE to tests/golden/*.npz.

The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)
sys.dont_write_bytecode = True

import lentil  # noqa: E402  (the reference)
import lentil.extent  # noqa: E402
import lentil.field  # noqa: E402
import lentil.fourier  # noqa: E402
import lentil.helper  # noqa: E402
import lentil_oracle as oc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def rel(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def ref_fields(w):
    return [{"data": f.data, "offset": tuple(int(v) for v in f.offset)} for f in w.data]


def pack_fields(prefix, fields, d):
    d[prefix + "_n"] = np.array(len(fields))
    for i, f in enumerate(fields):
        d[f"{prefix}_{i}_data"] = np.asarray(f["data"])
        d[f"{prefix}_{i}_offset"] = np.asarray(f["offset"], dtype=np.int64)


def check_fields(a, b, tol=0.0):
    assert len(a) == len(b), (len(a), len(b))
    for fa, fb in zip(a, b):
        assert tuple(int(v) for v in fa["offset"]) == tuple(int(v) for v in fb["offset"])
        assert fa["data"].shape == fb["data"].shape
        if fa["data"].size:
            assert rel(fa["data"], fb["data"]) <= tol, rel(fa["data"], fb["data"])


# ------------------------------------------------------------------ dft2 / idft2
def golden_dft2():
    rng = np.random.default_rng(1234)
    cases = [
        # m, n, M, N, alpha, shift, offset, unitary
        (10, 10, 10, 10, 1 / 10, (0, 0), (0, 0), False),
        (11, 11, 11, 11, 1 / 11, (0, 0), (0, 0), False),
        (10, 11, 10, 11, (1 / 10, 1 / 11), (0, 0), (0, 0), False),
        (24, 31, 40, 27, (0.013, 0.0071), (1.25, -3.5), (3, -7), True),
        (33, 18, 9, 50, (0.021, 0.017), (-12.75, 20.0), (-40, 11), True),
        (64, 64, 96, 80, 1 / 128, (0.4, 0.6), (0, 0), True),
        (7, 5, 130, 3, (0.002, 0.3), (0, 0.5), (100, -100), False),
    ]
    d = {"ncase": np.array(len(cases))}
    for i, (m, n, M, N, alpha, shift, offset, unitary) in enumerate(cases):
        f = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
        F = lentil.fourier.dft2(f, alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary)
        Fo = oc.dft2(f, alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary)
        assert rel(Fo, F) == 0.0, ("dft2", i, rel(Fo, F))
        g = lentil.fourier.idft2(f, alpha, shape=(M, N), shift=shift, unitary=unitary)
        go = oc.idft2(f, alpha, shape=(M, N), shift=shift, unitary=unitary)
        assert rel(go, g) == 0.0, ("idft2", i)
        d[f"c{i}_f"] = f
        d[f"c{i}_alpha"] = np.broadcast_to(alpha, (2,)).astype(float)
        d[f"c{i}_shape"] = np.array([M, N])
        d[f"c{i}_shift"] = np.array(shift, dtype=float)
        d[f"c{i}_offset"] = np.array(offset, dtype=float)
        d[f"c{i}_unitary"] = np.array(unitary)
        d[f"c{i}_F"] = F
        d[f"c{i}_iF"] = g
    np.savez_compressed(os.path.join(GOLD, "dft2.npz"), **d)
    print("dft2: oracle == reference bit-for-bit on", len(cases), "cases")


# ------------------------------------------------------------------ extent / helper
def golden_extent():
    rng = np.random.default_rng(7)
    rows = []
    for _ in range(200):
        sa = tuple(int(v) for v in rng.integers(1, 40, 2))
        sb = tuple(int(v) for v in rng.integers(1, 40, 2))
        ha = tuple(float(v) for v in rng.integers(-30, 30, 2))
        hb = tuple(float(v) for v in rng.integers(-30, 30, 2))
        ea, eb = lentil.extent.array_extent(sa, ha), lentil.extent.array_extent(sb, hb)
        assert ea == oc.array_extent(sa, ha) and eb == oc.array_extent(sb, hb)
        assert lentil.extent.array_center(ea) == oc.array_center(ea)
        hit = lentil.extent.intersect(ea, eb)
        assert hit == oc.intersect(ea, eb)
        ishape = lentil.extent.intersection_shape(ea, eb)
        assert ishape == oc.intersection_shape(ea, eb)
        ishift = lentil.extent.intersection_shift(ea, eb)
        assert ishift == oc.intersection_shift(ea, eb)
        assert lentil.extent.intersection_slices(ea, eb) == oc.intersection_slices(ea, eb)
        rows.append(list(sa) + list(sb) + list(ha) + list(hb) + list(ea) + list(eb) + [int(hit)]
                    + (list(ishape) if ishape else [0, 0]) + list(ishift))
    np.savez_compressed(os.path.join(GOLD, "extent.npz"), rows=np.array(rows, dtype=np.int64))

    # helper.boundary_slice / slice_offset (tests/test_helper.py:14-41)
    d = {}
    for i, (shape, radius) in enumerate([((64, 64), 16), ((63, 63), 12), ((40, 70), 9)]):
        shift = rng.uniform(-10, 10, 2).astype(int)
        a = lentil.circle(shape=shape, radius=radius, shift=shift, antialias=False)
        slc = lentil.helper.boundary_slice(a)
        off = lentil.helper.slice_offset(slc, a.shape)
        assert slc == oc.boundary_slice(a)
        assert tuple(off) == tuple(oc.slice_offset(slc, a.shape))
        d[f"h{i}_a"] = a.astype(np.uint8)
        d[f"h{i}_slice"] = np.array([slc[0].start, slc[0].stop, slc[1].start, slc[1].stop])
        d[f"h{i}_offset"] = np.array(off, dtype=np.int64)
    d["nh"] = np.array(3)
    np.savez_compressed(os.path.join(GOLD, "helper.npz"), **d)
    print("extent/helper: oracle == reference on 200 random rectangle pairs + 3 masks")


# ------------------------------------------------------------------ Field KATs (tests/test_field.py)
def golden_field():
    kats = [
        (1, [0, 0], 1, [0, 0]),
        (1, [0, 0], 1, [1, 0]),
        (np.ones((2, 2)), [0, 0], np.ones((2, 2)), [0, 0]),
        (np.ones((2, 2)), [0, 0], np.ones((2, 2)), [4, 4]),
        (1, [0, 0], np.ones((3, 3)), [-5, -5]),
        (np.ones((5, 4)), [-2, -2], np.ones((3, 3)), [0, -1]),
    ]
    rng = np.random.default_rng(3)
    for _ in range(20):
        sa, sb = rng.integers(1, 9, 2), rng.integers(1, 9, 2)
        kats.append((rng.normal(size=sa) + 1j * rng.normal(size=sa), list(rng.integers(-6, 6, 2)),
                     rng.normal(size=sb) + 1j * rng.normal(size=sb), list(rng.integers(-6, 6, 2))))
    d = {"n": np.array(len(kats))}
    for i, (ad, ao, bd, bo) in enumerate(kats):
        c = lentil.field.Field(ad, pixelscale=1, offset=ao) * lentil.field.Field(bd, pixelscale=1, offset=bo)
        co = oc.field_mul(oc.make_field(ad, ao), oc.make_field(bd, bo))
        if c.size == 0:
            assert co is None
        else:
            assert np.array_equal(c.data, co["data"]) and tuple(c.offset) == tuple(co["offset"])
        d[f"k{i}_a"], d[f"k{i}_ao"] = np.asarray(ad, dtype=complex), np.asarray(ao)
        d[f"k{i}_b"], d[f"k{i}_bo"] = np.asarray(bd, dtype=complex), np.asarray(bo)
        d[f"k{i}_empty"] = np.array(c.size == 0)
        d[f"k{i}_c"] = np.asarray(c.data, dtype=complex)
        d[f"k{i}_co"] = np.asarray(c.offset if c.size else [0, 0])
    # merge KAT (tests/test_field.py:30-40) + random reduce
    a = lentil.field.Field(np.ones((5, 4)), pixelscale=1, offset=[-2, -2])
    b = lentil.field.Field(np.ones((3, 3)), pixelscale=1, offset=[0, -1])
    c = lentil.field.merge(a, b)
    co = oc.merge_fields([oc.make_field(a.data, a.offset), oc.make_field(b.data, b.offset)])
    assert np.array_equal(c.data, co["data"]) and tuple(c.offset) == tuple(co["offset"])
    flds = []
    for _ in range(7):
        s = rng.integers(2, 9, 2)
        flds.append((rng.normal(size=s) + 1j * rng.normal(size=s), [int(v) for v in rng.integers(-14, 14, 2)]))
    red = lentil.field.reduce([lentil.field.Field(x, pixelscale=1, offset=o) for x, o in flds])
    redo = oc.reduce_fields([oc.make_field(x, o) for x, o in flds])
    check_fields(redo, [{"data": f.data, "offset": f.offset} for f in red])
    out = np.zeros((24, 30))
    for f in red:
        out = lentil.field.insert(f, out, intensity=True, weight=0.7)
    outo = oc.wavefront_insert([oc.make_field(x, o) for x, o in flds], np.zeros((24, 30)), 0.7)
    assert np.array_equal(out, outo)
    d["r_n"] = np.array(len(flds))
    for i, (x, o) in enumerate(flds):
        d[f"r{i}_data"], d[f"r{i}_offset"] = x, np.asarray(o)
    d["r_intensity"] = out
    np.savez_compressed(os.path.join(GOLD, "field.npz"), **d)
    print("field: oracle == reference on", len(kats), "multiply KATs, merge KAT, reduce+insert")


# ------------------------------------------------------------------ plane multiply + propagate_dft
def make_pupil(n, radius, coeffs, rng, nseg=0):
    if nseg:
        cube = lentil.hex_segments(rings=1, seg_radius=radius, seg_gap=2, flatten=False)
        cube = cube[:nseg]
        amp = np.sum(cube, axis=0).astype(float)
        mask = cube > 0
        opd = np.zeros(amp.shape)
        for s in range(nseg):
            c = rng.uniform(-1, 1, 3) * np.array([40e-9, 1.5e-6, 1.5e-6])
            opd += lentil.zernike_compose(cube[s], c)
        return amp / np.sqrt(np.sum(amp ** 2)), opd, mask
    amp = lentil.normalize_power(lentil.circle((n, n), radius))
    opd = lentil.zernike_compose(amp, coeffs) if coeffs is not None else 0
    return amp, opd, None


def golden_propagate():
    rng = np.random.default_rng(99)
    d = {}
    # --- case A: monolithic, field-point tilt, 3 wavelengths accumulated with insert -----------
    n, radius = 64, 28
    amp, opd, _ = make_pupil(n, radius, rng.normal(size=8) * 40e-9, rng)
    dx, z, du = 1.0 / (2 * radius), 10.0, 5e-6
    wls, wts = [550e-9, 650e-9, 800e-9], [0.2, 0.5, 0.3]
    tilt = [3.3e-6, -1.7e-6]
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    img = np.zeros((64, 64))
    for k, (wl, wt) in enumerate(zip(wls, wts)):
        w = lentil.Wavefront(wl, tilt=tilt) * p
        fo = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex), None, [oc.tilt_entry(*tilt)])],
                               amp, opd, None, wl)
        check_fields(fo, ref_fields(w))
        if k == 0:
            pack_fields("A_phasor", ref_fields(w), d)
        w = lentil.propagate_dft(w, pixelscale=du, shape=(32, 32), oversample=2)
        po, _ = oc.propagate_dft(fo, wl, (dx, dx), z, du, (32, 32), None, 2)
        check_fields(po, ref_fields(w))
        if k == 0:
            pack_fields("A_prop", ref_fields(w), d)
        img = w.insert(img, wt)
    imgo = oc.psf(amp, opd, None, wls, wts, (dx, dx), z, du, (32, 32), None, 2, wf_tilt=tilt)
    assert np.array_equal(img, imgo)
    d.update(A_amp=amp, A_opd=opd, A_dx=np.array(dx), A_z=np.array(z), A_du=np.array(du),
             A_wls=np.array(wls), A_wts=np.array(wts), A_tilt=np.array(tilt), A_shape=np.array([32, 32]),
             A_oversample=np.array(2), A_img=img)

    # --- case B: 3 hex segments, fit_tilt, prop_shape smaller than shape --------------------------
    amp, opd, mask = make_pupil(0, 14, None, rng, nseg=3)
    dx, z, du = 1.0 / 80, 12.0, 5e-6
    p = lentil.Pupil(amplitude=amp, opd=opd, mask=mask, pixelscale=dx, focal_length=z)
    p = p.fit_tilt(inplace=False)
    ptilt = [(t.x, t.y) for t in p.tilt]          # already-swapped attributes
    wl = 600e-9
    w = lentil.Wavefront(wl) * p
    fo = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], p.amplitude, p.opd, mask, wl, ptilt)
    check_fields(fo, ref_fields(w))
    w2 = lentil.propagate_dft(w, pixelscale=du, shape=(48, 48), prop_shape=(24, 24), oversample=2)
    po, _ = oc.propagate_dft(fo, wl, (dx, dx), z, du, (48, 48), (24, 24), 2)
    check_fields(po, ref_fields(w2))
    inten = w2.intensity
    assert np.array_equal(inten, oc.wavefront_intensity(po, (96, 96)))
    assert np.array_equal(w2.field, oc.wavefront_field(po, (96, 96)))
    pack_fields("B_phasor", ref_fields(w), d)
    pack_fields("B_prop", ref_fields(w2), d)
    d.update(B_amp=p.amplitude, B_opd=p.opd, B_mask=mask.astype(np.uint8), B_dx=np.array(dx), B_z=np.array(z),
             B_du=np.array(du), B_wl=np.array(wl), B_ptilt=np.array(ptilt), B_shape=np.array([48, 48]),
             B_prop_shape=np.array([24, 24]), B_oversample=np.array(2), B_intensity=inten, B_field=w2.field,
             B_opd_before_fit=opd)

    # --- case C: detector mask (tests/test_propagate_mask.py) ----------------------------------------
    amp, opd, _ = make_pupil(64, 28, np.array([0, 1e-6, 2e-6]), rng)
    dx, z, du, wl = 1.0 / 56, 10.0, 5e-6, 650e-9
    omask = lentil.rectangle((64, 64), 20, 24, shift=(7, -9), antialias=False)
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    w = lentil.Wavefront(wl) * p
    w2 = lentil.propagate_dft(w, shape=32, pixelscale=du, oversample=2, mask=omask)
    fo = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], amp, opd, None, wl)
    po, _ = oc.propagate_dft(fo, wl, (dx, dx), z, du, 32, None, 2, omask)
    check_fields(po, ref_fields(w2))
    pack_fields("C_prop", ref_fields(w2), d)
    d.update(C_amp=amp, C_opd=opd, C_dx=np.array(dx), C_z=np.array(z), C_du=np.array(du), C_wl=np.array(wl),
             C_omask=omask.astype(np.uint8), C_intensity=w2.intensity)

    # --- case D: tilt pushes the PSF off the detector (tests/test_propagate.py:134-141) ------------
    amp, opd, _ = make_pupil(63, 28, np.array([0, 1e-3]), rng)
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1.0 / 56, focal_length=10.0).fit_tilt(inplace=False)
    w = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, shape=(16, 16), pixelscale=5e-6, oversample=2)
    assert len(w.data) == 0 and np.all(w.intensity == 0)
    fo = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], p.amplitude, p.opd, None, 650e-9,
                           [(t.x, t.y) for t in p.tilt])
    po, _ = oc.propagate_dft(fo, 650e-9, (1 / 56, 1 / 56), 10.0, 5e-6, (16, 16), None, 2)
    assert po == []
    d.update(D_amp=p.amplitude, D_opd=p.opd, D_ptilt=np.array([(t.x, t.y) for t in p.tilt]))

    np.savez_compressed(os.path.join(GOLD, "propagate.npz"), **d)
    print("plane multiply / propagate_dft / intensity / insert / field: oracle == reference bit-for-bit "
          "(monolithic+tilt, 3-segment fit_tilt with prop_shape, detector mask, off-detector)")


def golden_detector():
    rng = np.random.default_rng(77)
    img, cube = rng.random((48, 48)), rng.random((3, 30, 42))
    d = dict(img=img, cube=cube, rebin3=lentil.rebin(img, 3), rebin_cube2=lentil.rebin(cube, 2),
             pixel2=lentil.detector.pixel(img, 2), pixel3=lentil.detector.pixel(img[:45, :45], 3))
    assert np.array_equal(oc.rebin(img, 3), d["rebin3"]) and np.array_equal(oc.rebin(cube, 2), d["rebin_cube2"])
    assert np.array_equal(oc.pixel(img, 2), d["pixel2"]) and np.array_equal(oc.pixel(img[:45, :45], 3), d["pixel3"])
    np.savez_compressed(os.path.join(GOLD, "detector.npz"), **d)
    print("detector: oracle == reference bit-for-bit (rebin 2-D / cube, pixel MTF even / odd size)")


def golden_rescale():
    # lentil.rescale / lentil.detector.pixelate go through scipy.ndimage.map_coordinates (a dependency the
    # reference does not vendor): the oracle restates that algorithm, so the pin is numerical (<= 1e-13 of the
    # peak), not bit-for-bit; checked here with the scipy of the build container.
    import scipy
    rng = np.random.default_rng(78)
    img = rng.random((60, 60))
    img[:5] = 0
    img[:, 50:] = 0
    worst = 0.0
    for scale in (1 / 2, 1 / 3, 0.37, 1.7):
        for order in range(6):
            for mode in ('nearest', 'constant', 'reflect', 'wrap'):
                a = lentil.rescale(img, scale, order=order, mode=mode)
                b = oc.rescale(img, scale, order=order, mode=mode)
                worst = max(worst, float(np.max(np.abs(a - b)) / np.max(np.abs(a))))
    assert worst <= 1e-13, worst
    z = img + 1j * rng.random((60, 60))
    psf = rng.random((96, 96)) ** 8
    d = dict(img=img, z=z, psf=psf, scales=np.array([1 / 2, 1 / 3, 0.37, 1.7]))
    cases = {}
    for k, scale in enumerate(d["scales"]):
        cases[f"default_{k}"] = lentil.rescale(img, scale)
    for order in range(6):
        cases[f"order{order}_third"] = lentil.rescale(img, 1 / 3, order=order)
    for mode in ('constant', 'reflect', 'wrap'):
        cases[f"mode_{mode}"] = lentil.rescale(img, 0.37, mode=mode)
    cases["complex_half"] = lentil.rescale(z, 0.5)
    cases["shape40_ones_nonunitary"] = lentil.rescale(img, 0.5, shape=40, mask=np.ones_like(img), unitary=False)
    cases["pixelate3"] = lentil.detector.pixelate(psf, 3)
    cases["pixelate4"] = lentil.detector.pixelate(psf, 4)
    assert rel(oc.rescale(z, 0.5), cases["complex_half"]) <= 1e-13
    assert rel(oc.rescale(img, 0.5, shape=40, mask=np.ones_like(img), unitary=False), cases["shape40_ones_nonunitary"]) <= 1e-13
    assert rel(oc.pixelate(psf, 3), cases["pixelate3"]) <= 1e-13 and rel(oc.pixelate(psf, 4), cases["pixelate4"]) <= 1e-13
    d.update(cases)
    np.savez_compressed(os.path.join(GOLD, "rescale.npz"), **d)
    print(f"rescale / pixelate: oracle == reference to {worst:.1e} of the peak over 96 (scale, order, mode) cases "
          f"(scipy {scipy.__version__}); golden vectors written")


def golden_dispersive_tilt():
    # the reference solves higher-order trace/dispersion polynomials with scipy leastsq / quad to
    # ~1e-8 relative (lentil/plane.py:1037,1050); these vectors pin the host mirror to that level
    wl = np.linspace(500e-9, 900e-9, 9)
    cases = [([2.0, 0.0], [1.0, 650e-9]),                    # both first order (closed form)
             ([3.0, 0.2, 0.0], [5e-6, 650e-9]),              # second-order trace
             ([0.5, 1e-3], [2e-4, 1e-5, 500e-9]),            # second-order dispersion
             ([40.0, -3.0, 0.1, 0.0], [1e-3, 2e-5, 450e-9])]   # third-order trace, second-order dispersion
    d = dict(wl=wl, n=np.array(len(cases)))
    for i, (trace, disp) in enumerate(cases):
        dt = lentil.DispersiveTilt(trace=trace, dispersion=disp)
        xy = np.array([[float(np.ravel(v)[0]) for v in dt.__shift__(wavelength=w, xs=1e-3, ys=-2e-3)] for w in wl])
        d[f"c{i}_trace"], d[f"c{i}_disp"], d[f"c{i}_xy"] = np.array(trace), np.array(disp), xy
    np.savez_compressed(os.path.join(GOLD, "dispersive_tilt.npz"), **d)
    print("dispersive tilt: reference shifts for", len(cases), "polynomial pairs x", len(wl), "wavelengths")


def golden_propagate_fft():
    """propagate_fft (lentil/propagate.py:9-88): even and odd padded sizes, default and explicit
    shape, one segmented pupil; plus the scratch_shape helper and the two error cases."""
    rng = np.random.default_rng(5)
    d = {}
    cases = [  # n, radius, dx, z, du, wl, shape, oversample
        (40, 18, 1 / 36, 10.0, 5e-6, 650e-9, (24, 24), 2),     # alpha -> 94 samples (even)
        (40, 18, 1 / 36, 10.0, 5e-6, 655e-9, None, 2),          # -> 94, default shape
        (41, 19, 1 / 38, 8.0, 6e-6, 600e-9, (20, 30), 1),       # -> odd padded size (30.4 -> 30; 8*600e-9*38/6e-6)
        (36, 16, 1 / 32, 10.0, 5e-6, 555e-9, (15, 15), 3),      # odd padded size 107
    ]
    for i, (n, radius, dx, z, du, wl, shape, os_) in enumerate(cases):
        amp, opd, _ = make_pupil(n, radius, rng.normal(size=6) * 40e-9, rng)
        p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
        w = lentil.Wavefront(wl) * p
        w2 = lentil.propagate_fft(w, pixelscale=du, shape=shape, oversample=os_)
        fo = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], amp, opd, None, wl)
        Fo, shape_out, pw = oc.propagate_fft(fo, w.shape, wl, (dx, dx), z, du, shape, os_)
        assert np.array_equal(Fo["data"], w2.data[0].data) and tuple(shape_out) == tuple(w2.shape)
        assert pw == w2.wavelength
        assert np.array_equal(oc.wavefront_intensity([Fo], shape_out), w2.intensity)
        d.update({f"c{i}_amp": amp, f"c{i}_opd": opd, f"c{i}_dx": np.array(dx), f"c{i}_z": np.array(z),
                  f"c{i}_du": np.array(du), f"c{i}_wl": np.array(wl),
                  f"c{i}_shape": np.array(shape if shape is not None else (-1, -1)), f"c{i}_os": np.array(os_),
                  f"c{i}_shape_out": np.array(w2.shape),
                  f"c{i}_prop_wl": np.array(w2.wavelength), f"c{i}_intensity": w2.intensity})
        if i != 1:                                   # same padded size as case 0: keep the fixture small
            d[f"c{i}_F"] = w2.data[0].data
        print("  fft case", i, "padded", w2.data[0].data.shape, "shape_out", tuple(w2.shape))
    d["n"] = np.array(len(cases))
    d["scratch"] = np.array(lentil.propagate.scratch_shape([500e-9, 700e-9], 1 / 36, 5e-6, 10.0, 2))
    np.savez_compressed(os.path.join(GOLD, "propagate_fft.npz"), **d)
    print("propagate_fft: oracle == reference bit-for-bit on", len(cases), "cases")


def golden_power_spectrum():
    d = {}
    for i, (n, radius, seed) in enumerate([(48, 20, 3), (45, 19, 7)]):
        mask = lentil.circle((n, n), radius, antialias=False)
        args = dict(pixelscale=1 / (2 * radius), rms=30e-9, half_power_freq=5, exp=3, seed=seed)
        ref = lentil.power_spectrum(mask, **args)
        assert np.array_equal(oc.power_spectrum(mask, **args), ref)
        d[f"c{i}_mask"], d[f"c{i}_opd"], d[f"c{i}_seed"], d[f"c{i}_radius"] = mask.astype(np.uint8), ref, np.array(seed), np.array(radius)
    d["n"] = np.array(2)
    np.savez_compressed(os.path.join(GOLD, "power_spectrum.npz"), **d)
    print("power_spectrum: oracle == reference bit-for-bit (even and odd grid)")


if __name__ == "__main__":
    assert lentil.__version__ == "0.8.8", lentil.__version__
    golden_dft2()
    golden_extent()
    golden_field()
    golden_propagate()
    golden_detector()
    golden_rescale()
    golden_dispersive_tilt()
    golden_propagate_fft()
    golden_power_spectrum()
    sizes = {f: os.path.getsize(os.path.join(GOLD, f)) for f in sorted(os.listdir(GOLD))}
    print("fixtures:", sizes)
