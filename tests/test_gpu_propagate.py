"""End-to-end parity of the hot path: Plane.__mul__ (K1) -> propagate_dft (K2a) ->
Wavefront.intensity / insert / field (K3), against the reference's golden vectors, the oracle on
seeded inputs at the BASELINE shapes, and the reference's own property tests."""
import numpy as np
import pytest

import lentil_b200 as lentil
import lentil_oracle as oc
from lentil_b200 import synth
from conftest import peak_err, unpack_fields, TOL64

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures('mft_variant')]


def _check_fields(wavefront, gold, tol=TOL64):
    assert len(wavefront.data) == len(gold)
    for f, (data, off) in zip(wavefront.data, gold):
        assert tuple(int(v) for v in f.offset) == off
        assert tuple(f.shape) == data.shape
        assert peak_err(f.data, data) <= tol


def test_golden_case_A_tilted_polychromatic(golden):
    d = golden("propagate")
    dx, z, du = float(d["A_dx"]), float(d["A_z"]), float(d["A_du"])
    p = lentil.Pupil(amplitude=d["A_amp"], opd=d["A_opd"], pixelscale=dx, focal_length=z)
    tilt = list(d["A_tilt"])
    w = lentil.Wavefront(d["A_wls"][0], tilt=tilt) * p
    _check_fields(w, unpack_fields(d, "A_phasor"))
    w = lentil.propagate_dft(w, pixelscale=du, shape=tuple(d["A_shape"]), oversample=int(d["A_oversample"]))
    _check_fields(w, unpack_fields(d, "A_prop"))
    img = np.zeros(d["A_img"].shape)
    for wl, wt in zip(d["A_wls"], d["A_wts"]):
        w = lentil.Wavefront(wl, tilt=tilt) * p
        w = lentil.propagate_dft(w, pixelscale=du, shape=tuple(d["A_shape"]), oversample=int(d["A_oversample"]))
        img = w.insert(img, wt)
    assert peak_err(img, d["A_img"]) <= TOL64
    # the batched driver computes the same stack in one call
    img_b = lentil.propagate_dft_batch(p, d["A_wls"], du, tuple(d["A_shape"]), oversample=int(d["A_oversample"]),
                                       weights=d["A_wts"], tilts=[tilt])
    assert img_b.shape == (1,) + d["A_img"].shape
    assert peak_err(img_b[0], d["A_img"]) <= TOL64


def test_golden_case_B_segments_fit_tilt_prop_shape(golden):
    d = golden("propagate")
    dx = float(d["B_dx"])
    p = lentil.Pupil(amplitude=d["B_amp"], opd=d["B_opd"], mask=d["B_mask"].astype(bool), pixelscale=dx,
                     focal_length=float(d["B_z"]))
    for tx, ty in d["B_ptilt"]:          # stored attributes are already swapped: Tilt(x=ty, y=tx)
        p.tilt.append(lentil.Tilt(x=ty, y=tx))
    w = lentil.Wavefront(float(d["B_wl"])) * p
    _check_fields(w, unpack_fields(d, "B_phasor"))
    w2 = lentil.propagate_dft(w, pixelscale=float(d["B_du"]), shape=tuple(d["B_shape"]),
                              prop_shape=tuple(d["B_prop_shape"]), oversample=int(d["B_oversample"]))
    _check_fields(w2, unpack_fields(d, "B_prop"))
    assert peak_err(w2.intensity, d["B_intensity"]) <= TOL64
    assert peak_err(w2.field, d["B_field"]) <= TOL64
    img_b = lentil.propagate_dft_batch(p, [float(d["B_wl"])], float(d["B_du"]), tuple(d["B_shape"]),
                                       prop_shape=tuple(d["B_prop_shape"]), oversample=int(d["B_oversample"]))
    assert peak_err(img_b, d["B_intensity"]) <= TOL64


def test_golden_fit_tilt_matches_reference(golden):
    d = golden("propagate")
    p = lentil.Pupil(amplitude=d["B_amp"], opd=d["B_opd_before_fit"], mask=d["B_mask"].astype(bool),
                     pixelscale=float(d["B_dx"]), focal_length=float(d["B_z"]))
    q = p.fit_tilt()
    got = np.array([(t.x, t.y) for t in q.tilt])
    assert np.allclose(got, d["B_ptilt"], rtol=1e-9, atol=1e-18)
    assert peak_err(q.opd, d["B_opd"]) <= 1e-9


def test_golden_case_C_detector_mask(golden):
    d = golden("propagate")
    p = lentil.Pupil(amplitude=d["C_amp"], opd=d["C_opd"], pixelscale=float(d["C_dx"]), focal_length=float(d["C_z"]))
    w = lentil.Wavefront(float(d["C_wl"])) * p
    wm = lentil.propagate_dft(w, shape=32, pixelscale=float(d["C_du"]), oversample=2, mask=d["C_omask"])
    _check_fields(wm, unpack_fields(d, "C_prop"))
    assert peak_err(wm.intensity, d["C_intensity"]) <= TOL64
    # reference tests/test_propagate_mask.py:4-37: psf_mask == psf * mask
    full = lentil.propagate_dft(lentil.Wavefront(float(d["C_wl"])) * p, shape=32, pixelscale=float(d["C_du"]),
                                oversample=2).intensity
    assert np.allclose(wm.intensity, full * d["C_omask"])
    with pytest.raises(ValueError):
        lentil.propagate_dft(w, shape=32, pixelscale=5e-6, oversample=2, mask=np.ones((10, 10)))


def test_golden_case_D_no_output(golden):
    d = golden("propagate")
    p = lentil.Pupil(amplitude=d["D_amp"], opd=d["D_opd"], pixelscale=1 / 56, focal_length=10.0)
    for tx, ty in d["D_ptilt"]:
        p.tilt.append(lentil.Tilt(x=ty, y=tx))
    w = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, shape=(16, 16), pixelscale=5e-6, oversample=2)
    assert w.data == [] and np.all(w.intensity == 0) and w.intensity.shape == (32, 32)


def test_field_kats_on_device(golden):
    # reference tests/test_field.py:7-40 (exact small known answers) + random rectangles
    d = golden("field")
    for i in range(int(d["n"])):
        c = lentil.Field(d[f"k{i}_a"], pixelscale=1, offset=list(d[f"k{i}_ao"])) * \
            lentil.Field(d[f"k{i}_b"], pixelscale=1, offset=list(d[f"k{i}_bo"]))
        if bool(d[f"k{i}_empty"]):
            assert c.size == 0
        else:
            assert peak_err(c.data, d[f"k{i}_c"]) <= 1e-15
            assert tuple(int(v) for v in c.offset) == tuple(d[f"k{i}_co"])
    a = lentil.Field(np.ones((5, 4)), pixelscale=1, offset=[-2, -2])
    b = lentil.Field(np.ones((3, 3)), pixelscale=1, offset=[0, -1])
    c = lentil.field.merge(a, b)
    expect = np.zeros((6, 5))
    expect[0:5, 0:4] += 1
    expect[3:6, 2:5] += 1
    assert np.array_equal(c.data, expect) and tuple(c.offset) == (-1, -2)
    with pytest.raises(ValueError):
        lentil.field.merge(a, lentil.Field(np.ones((2, 2)), pixelscale=1, offset=[40, 40]))
    flds = [lentil.Field(d[f"r{i}_data"], pixelscale=1, offset=list(d[f"r{i}_offset"])) for i in range(int(d["r_n"]))]
    out = np.zeros(d["r_intensity"].shape)
    for f in lentil.field.reduce(flds):
        out = lentil.field.insert(f, out, intensity=True, weight=0.7)
    assert peak_err(out, d["r_intensity"]) <= 1e-14


def test_plane_multiply_phasor():
    # reference tests/test_plane.py:83-116
    rng = np.random.default_rng(0)
    mask = synth.circle((256, 256), 64, shift=(10, -20))
    amp, opd = rng.uniform(size=(256, 256)) * mask, rng.normal(size=(256, 256)) * 200e-9 * mask
    p = lentil.Plane(amplitude=amp, opd=opd, mask=mask)
    w1 = lentil.Wavefront(650e-9) * p
    slc = lentil.helper.boundary_slice(mask)
    assert peak_err(w1.data[0].data, amp[slc] * np.exp(2 * np.pi * 1j * opd[slc] / 650e-9)) <= 1e-14
    assert tuple(w1.data[0].offset) == (10, -20)
    w2 = p * lentil.Wavefront(650e-9)
    assert peak_err(w2.data[0].data, w1.data[0].data) == 0.0


def test_overlapping_segment_bboxes():
    # reference tests/test_plane.py:119-133: neighbours intruding into a bbox are masked out
    cube = synth.hex_segments(1, 20, 1)[:2]
    mask = cube.sum(axis=0).astype(float)
    pupil = lentil.Pupil(amplitude=mask, mask=cube, pixelscale=1 / 256, focal_length=10)
    w = lentil.Wavefront(500e-9) * pupil
    assert len(w.data) == 2
    assert np.array_equal(mask, w.intensity)


def test_pupil_hands_over_focal_length_and_frozen_cache():
    amp = synth.normalize_power(synth.circle((64, 64), 20))
    p = lentil.Pupil(amplitude=amp, opd=np.zeros((64, 64)), pixelscale=1 / 40, focal_length=7.0)
    w = lentil.Wavefront(650e-9) * p
    assert w.focal_length == 7.0 and w.ptype == lentil.pupil and tuple(w.shape) == (64, 64)
    a = lentil.propagate_dft(w, 5e-6, shape=32).intensity
    p.freeze()
    b = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, 5e-6, shape=32).intensity
    c = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, 5e-6, shape=32).intensity   # cached operands
    p.thaw()
    assert np.array_equal(a, b) and np.array_equal(b, c)


def _cfg(n, radius, ncoef, seed, obsc=None):
    rng = np.random.default_rng(seed)
    mask = synth.circle((n, n), radius) if obsc is None else synth.annulus((n, n), radius, obsc)
    amp = synth.normalize_power(mask)
    opd = synth.zernike_opd(mask, rng.normal(size=ncoef) * 30e-9) if ncoef else np.zeros((n, n))
    return amp, opd


def test_cfg1_monochromatic_airy_against_oracle():
    # BASELINE config 1: 256^2 circular pupil (bbox 241^2) -> 128^2 detector, os 2, 650 nm
    amp, opd = _cfg(256, 120, 0, 0)
    dx, z, du = 1 / 240, 10.0, 5e-6
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    w = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, pixelscale=du, shape=(128, 128), oversample=2)
    assert tuple(w.data[0].shape) == (256, 256)
    ref = oc.psf(amp, opd, None, [650e-9], [1.0], (dx, dx), z, du, (128, 128), None, 2)
    assert peak_err(w.intensity, ref) <= TOL64
    assert 0.95 <= w.intensity.sum() <= 1.0          # reference tests/test_propagate.py:8-23


def test_airy_analytic():
    # reference tests/test_propagate.py:26-42 (atol 1e-3 peak-normalised against the Airy pattern)
    from scipy.special import jn
    amp, _ = _cfg(512, 250, 0, 0)
    p = lentil.Pupil(amplitude=amp, pixelscale=1 / 500, focal_length=10.0, opd=np.zeros((512, 512)))
    w = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, shape=511, pixelscale=5e-6, oversample=1)
    psf = w.intensity
    psf = psf / psf.max()
    y, x = np.indices((511, 511), dtype=float)
    q = np.hypot((x - 255) * 5e-6, (y - 255) * 5e-6)
    X = np.pi * q / (650e-9 * 10.0 / p.diameter)
    X[X == 0] = np.finfo(float).eps
    airy = (2 * jn(1, X) / X) ** 2
    assert np.all(np.isclose(psf, airy / airy.max(), atol=1e-3))


def test_cfg2_polychromatic_obscured_against_oracle():
    # BASELINE config 2 at full shape (1024^2 annular pupil -> 512^2 det x os 2), 3 of the 100 wavelengths
    amp, opd = _cfg(1024, 500, 15, 0, obsc=1 / 3)
    dx, z, du = 1 / 1000, 20.0, 5e-6
    wls = np.linspace(500e-9, 900e-9, 100)[[0, 49, 99]]
    wts = np.array([0.2, 0.5, 0.3])
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    img = lentil.propagate_dft_batch(p, wls, du, (512, 512), oversample=2, weights=wts)
    ref = oc.psf(amp, opd, None, wls, wts, (dx, dx), z, du, (512, 512), None, 2)
    assert img.shape == (1024, 1024)
    assert peak_err(img, ref) <= TOL64
    # the drop-in loop gives the same image
    loop = np.zeros((1024, 1024))
    p.freeze()
    for wl, wt in zip(wls, wts):
        w = lentil.propagate_dft(lentil.Wavefront(wl) * p, du, (512, 512), oversample=2)
        loop = w.insert(loop, wt)
    assert peak_err(loop, ref) <= TOL64


def test_cfg3_like_segmented_against_oracle():
    # BASELINE config 3 geometry at 1/4 linear scale: 18 hexes, per-segment piston/tip/tilt, fit_tilt,
    # sparse windows (prop_shape < shape) merged coherently
    rng = np.random.default_rng(1)
    cube = synth.hex_segments(2, 58, 2)
    n = cube.shape[1]
    amp = synth.normalize_power(cube.sum(axis=0).astype(float))
    opd = np.zeros((n, n))
    for s in range(18):
        opd += synth.zernike_opd(cube[s], rng.uniform(-1, 1, 3) * np.array([50e-9, 5e-7, 5e-7]))
    dx, z, du = 1 / 500, 20.0, 5e-6
    p = lentil.Pupil(amplitude=amp, opd=opd, mask=cube, pixelscale=dx, focal_length=z).fit_tilt()
    ptilt = [(t.x, t.y) for t in p.tilt]
    wls, wts = [550e-9, 800e-9], [0.6, 0.4]
    img = lentil.propagate_dft_batch(p, wls, du, (128, 128), prop_shape=(64, 64), oversample=2, weights=wts)
    ref = oc.psf(p.amplitude, p.opd, cube, wls, wts, (dx, dx), z, du, (128, 128), (64, 64), 2, plane_tilt=ptilt)
    assert peak_err(img, ref) <= TOL64
    w = lentil.propagate_dft(lentil.Wavefront(wls[0]) * p, du, (128, 128), prop_shape=(64, 64), oversample=2)
    assert len(w.data) == 18
    ref0 = oc.psf(p.amplitude, p.opd, cube, wls[:1], [1.0], (dx, dx), z, du, (128, 128), (64, 64), 2, plane_tilt=ptilt)
    assert peak_err(w.intensity, ref0) <= TOL64


def test_field_points_stack_against_oracle():
    # BASELINE config 5 structure at small scale: field points x wavelengths -> (P, H, W) stack
    amp, opd = _cfg(128, 60, 10, 3, obsc=1 / 3)
    dx, z, du = 1 / 120, 20.0, 5e-6
    tilts = [[rx, ry] for rx in (-8e-6, 8e-6) for ry in (-5e-6, 12e-6)]
    wls = np.linspace(500e-9, 900e-9, 5)
    wts = np.full(5, 0.2)
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    stack = lentil.propagate_dft_batch(p, wls, du, (64, 64), oversample=2, weights=wts, tilts=tilts)
    assert stack.shape == (4, 128, 128)
    for k, t in enumerate(tilts):
        ref = oc.psf(amp, opd, None, wls, wts, (dx, dx), z, du, (64, 64), None, 2, wf_tilt=t)
        assert peak_err(stack[k], ref) <= TOL64
    # tiny chunks exercise the chunked path
    stack2 = lentil.propagate_dft_batch(p, wls, du, (64, 64), oversample=2, weights=wts, tilts=tilts, chunk_bytes=1)
    assert peak_err(stack2, stack) <= 1e-14          # chunking only changes the summation association


def test_insert_accumulates_on_device_and_is_deterministic():
    amp, opd = _cfg(128, 60, 6, 4)
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 120, focal_length=20.0)
    acc = lentil.device.zeros_f64(128, 128)
    host = np.zeros((128, 128))
    for wl in (5e-7, 6e-7, 7e-7):
        w = lentil.propagate_dft(lentil.Wavefront(wl) * p, 5e-6, (64, 64), oversample=2)
        acc = w.insert(acc, 0.3)
        host = w.insert(host, 0.3)
    a = lentil.device.to_host(acc)
    assert peak_err(a, host) <= 1e-15
    b = lentil.propagate_dft_batch(p, [5e-7, 6e-7, 7e-7], 5e-6, (64, 64), weights=[0.3] * 3)
    c = lentil.propagate_dft_batch(p, [5e-7, 6e-7, 7e-7], 5e-6, (64, 64), weights=[0.3] * 3)
    assert np.array_equal(b, c)                       # owner-computes accumulation: bit-identical reruns
    assert peak_err(b, a) <= 1e-14


def test_tilt_plane_and_product_of_planes():
    # a Tilt plane multiplies by 1 and books a shift; two array planes multiply on the overlap
    amp, opd = _cfg(96, 40, 5, 5)
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 80, focal_length=10.0)
    w = lentil.Wavefront(650e-9) * p
    w = w * lentil.Tilt(x=4e-6, y=-3e-6)
    a = lentil.propagate_dft(w, 5e-6, (48, 48), oversample=2).intensity
    ref = oc.psf(amp, opd, None, [650e-9], [1.0], (1 / 80, 1 / 80), 10.0, 5e-6, (48, 48), None, 2,
                 wf_tilt=[4e-6, -3e-6])
    assert peak_err(a, ref) <= TOL64
    stop = lentil.Pupil(amplitude=synth.circle((96, 96), 25, shift=(5, -7)), pixelscale=1 / 80, focal_length=10.0,
                        opd=np.zeros((96, 96)))
    w2 = (lentil.Wavefront(650e-9) * p) * stop
    f0 = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], amp, opd, None, 650e-9)
    f1 = oc.plane_multiply(f0, stop.amplitude, stop.opd, None, 650e-9)
    assert tuple(int(v) for v in w2.data[0].offset) == tuple(f1[0]["offset"])
    assert peak_err(w2.data[0].data, f1[0]["data"]) <= 1e-14


def test_monte_carlo_opd_stack_against_oracle():
    # BASELINE config 4 structure at small scale: WFE realisations x wavelengths (x focus diversity folded
    # into the realisation axis) -> one polychromatic PSF per realisation
    rng = np.random.default_rng(11)
    mask = synth.circle((96, 96), 45)
    amp = synth.normalize_power(mask)
    opds = np.stack([synth.zernike_opd(mask, rng.normal(size=12) * 20e-9, first=4) for _ in range(5)])
    dx, z, du = 1 / 90, 20.0, 5e-6
    wls = np.linspace(600e-9, 700e-9, 4)
    wts = np.array([0.1, 0.4, 0.3, 0.2])
    p = lentil.Pupil(amplitude=amp, opd=np.zeros((96, 96)), pixelscale=dx, focal_length=z)
    stack = lentil.propagate_dft_batch(p, wls, du, (48, 48), oversample=2, weights=wts, opds=opds)
    assert stack.shape == (5, 96, 96)
    for r in range(5):
        ref = oc.psf(amp, opds[r], None, wls, wts, (dx, dx), z, du, (48, 48), None, 2)
        assert peak_err(stack[r], ref) <= TOL64
    # realisations x field points, tiny chunks
    tilts = [[0.0, 0.0], [6e-6, -4e-6]]
    s2 = lentil.propagate_dft_batch(p, wls, du, (48, 48), oversample=2, weights=wts, opds=opds[:2], tilts=tilts,
                                    chunk_bytes=1)
    assert s2.shape == (2, 2, 96, 96)
    assert peak_err(s2[:, 0], stack[:2]) <= 1e-13
    ref = oc.psf(amp, opds[1], None, wls, wts, (dx, dx), z, du, (48, 48), None, 2, wf_tilt=tilts[1])
    assert peak_err(s2[1, 1], ref) <= TOL64


def test_fit_tilt_recovers_a_plane():
    rng = np.random.default_rng(8)
    n, dx = 48, 1 / 40
    amp = synth.circle((n, n), 18)
    rr, cc = lentil.helper.mesh((n, n))
    opd = (2e-7 + 3e-6 * rr * dx - 1.5e-6 * (-cc) * dx) * amp
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=5.0)
    q = p.fit_tilt(inplace=False)
    assert q is not p and len(p.tilt) == 0 and len(q.tilt) == 1
    assert np.allclose(q.opd[amp > 0], 2e-7, atol=1e-15)
    assert np.isclose(q.tilt[0].y, 3e-6) and np.isclose(q.tilt[0].x, -1.5e-6)
    assert p.fit_tilt(inplace=True) is p
    assert lentil.Image(amplitude=amp).fit_tilt() is not None


def test_fit_tilt_equals_tilt_in_opd():
    # reference tests/test_propagate.py:45-70: PSF centroid with fit_tilt == with the tilt left in the OPD
    rng = np.random.default_rng(12)
    mask = synth.circle((256, 256), 120)
    amp = synth.normalize_power(mask)
    coeffs = np.concatenate(([0.0], 5e-6 * rng.uniform(-0.5, 0.5, 2)))
    opd = synth.zernike_opd(mask, coeffs)
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 240, focal_length=10.0)
    a = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, shape=128, pixelscale=5e-6, oversample=2).intensity
    q = p.fit_tilt()
    b = lentil.propagate_dft(lentil.Wavefront(650e-9) * q, shape=128, pixelscale=5e-6, oversample=2).intensity
    a[a < 1e-5] = 0
    b[b < 1e-5] = 0
    rr, cc = np.indices(a.shape)
    ca = np.array([np.sum(rr * a), np.sum(cc * a)]) / np.sum(a)
    cb = np.array([np.sum(rr * b), np.sum(cc * b)]) / np.sum(b)
    assert np.all(np.abs(ca - cb) <= 1e-6)


def test_phase_only_plane_and_explicit_nonbool_mask():
    # scalar amplitude + array opd: the phasor covers the whole array and the plane has no shape of its own
    rng = np.random.default_rng(21)
    opd = rng.normal(size=(40, 50)) * 1e-7
    w = lentil.Wavefront(600e-9) * lentil.Plane(amplitude=2.0, opd=opd)
    assert len(w.data) == 1 and tuple(w.data[0].shape) == (40, 50) and tuple(w.data[0].offset) == (0, 0)
    assert peak_err(w.data[0].data, 2.0 * np.exp(2j * np.pi * opd / 600e-9)) <= 1e-14
    # explicit integer (0/1) mask smaller than the amplitude support: amp*mask and the mask's bbox are used
    amp = np.ones((64, 64))
    mask = np.zeros((64, 64), dtype=int)
    mask[10:30, 20:55] = 1
    p = lentil.Pupil(amplitude=amp, opd=np.zeros((64, 64)), mask=mask, pixelscale=1 / 60, focal_length=5.0)
    w = lentil.Wavefront(600e-9) * p
    f = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], amp, np.zeros((64, 64)), mask, 600e-9)
    assert tuple(w.data[0].shape) == f[0]["data"].shape == (20, 35)
    assert tuple(int(v) for v in w.data[0].offset) == tuple(f[0]["offset"])
    assert peak_err(w.data[0].data, f[0]["data"]) <= 1e-15
    with pytest.raises(IndexError):
        lentil.Wavefront(600e-9) * lentil.Pupil(amplitude=np.zeros((8, 8)), opd=np.zeros((8, 8)), pixelscale=1, focal_length=1)


# ---- propagate_fft (lentil/propagate.py:9-88): the FFT sibling, executed as a K2a launch -----------
def test_propagate_fft_golden(golden):
    d = golden("propagate_fft")
    for i in range(int(d["n"])):
        dx, z, du, os_ = float(d[f"c{i}_dx"]), float(d[f"c{i}_z"]), float(d[f"c{i}_du"]), int(d[f"c{i}_os"])
        shape = None if d[f"c{i}_shape"][0] < 0 else tuple(int(v) for v in d[f"c{i}_shape"])
        p = lentil.Pupil(amplitude=d[f"c{i}_amp"], opd=d[f"c{i}_opd"], pixelscale=dx, focal_length=z)
        w = lentil.propagate_fft(lentil.Wavefront(float(d[f"c{i}_wl"])) * p, pixelscale=du, shape=shape,
                                 oversample=os_)
        assert tuple(int(v) for v in w.shape) == tuple(int(v) for v in d[f"c{i}_shape_out"])
        assert w.wavelength == float(d[f"c{i}_prop_wl"])
        assert w.ptype == lentil.image and len(w.data) == 1
        if f"c{i}_F" in d:
            assert tuple(w.data[0].shape) == d[f"c{i}_F"].shape
            assert peak_err(w.data[0].data, d[f"c{i}_F"]) <= TOL64
        assert peak_err(w.intensity, d[f"c{i}_intensity"]) <= TOL64


def test_propagate_fft_errors_and_scratch():
    amp = synth.normalize_power(synth.annulus((32, 32), 14))
    w = lentil.Wavefront(650e-9) * lentil.Pupil(amplitude=amp, pixelscale=1 / 28, focal_length=10.0)
    with pytest.raises(ValueError):
        lentil.propagate_fft(w, pixelscale=5e-6, shape=(4000, 4000), oversample=2)
    with pytest.raises(ValueError):
        lentil.propagate_fft(w, pixelscale=5e-6, shape=(8, 8), scratch=np.zeros((16, 16), dtype=complex))
    big = np.zeros(tuple(v + 1 for v in lentil.scratch_shape(650e-9, 1 / 28, 5e-6, 10.0, 2)), dtype=complex)
    a = lentil.propagate_fft(w, pixelscale=5e-6, shape=(8, 8), scratch=big).intensity
    b = lentil.propagate_fft(w, pixelscale=5e-6, shape=(8, 8)).intensity
    assert np.array_equal(a, b)
    pl = lentil.Pupil(amplitude=amp, opd=np.fromfunction(lambda r, c: 1e-7 * r, (32, 32)), pixelscale=1 / 28,
                      focal_length=10.0).fit_tilt()
    with pytest.raises(NotImplementedError):
        lentil.propagate_fft(lentil.Wavefront(650e-9) * pl, pixelscale=5e-6, shape=(8, 8))


def test_propagate_fft_equals_dft_at_the_fft_sampling():
    # at the wavelength the integer padding really samples, both propagators evaluate the same sum
    amp = synth.normalize_power(synth.annulus((64, 64), 30, 0.3))
    p = lentil.Pupil(amplitude=amp, pixelscale=1 / 60, focal_length=10.0)     # flat: phasor independent of wavelength
    wf = lentil.propagate_fft(lentil.Wavefront(640e-9) * p, pixelscale=5e-6, shape=(24, 24), oversample=2)
    wd = lentil.propagate_dft(lentil.Wavefront(wf.wavelength) * p, pixelscale=5e-6, shape=(24, 24), oversample=2)
    assert wf.field.shape == wd.field.shape == (48, 48)
    assert peak_err(wf.field, wd.field) <= 1e-9
    assert peak_err(wf.intensity, wd.intensity) <= 1e-9


def test_fourier_level_patch_reroutes_a_numpy_propagation(golden):
    # rebinding <module>.dft2 sends a host-side propagate_dft (here the oracle's restatement of
    # lentil/propagate.py:147-242, which like the reference resolves dft2 through its module at call time) through K2a
    import lentil_oracle as oc
    from lentil_b200 import patch
    d = golden("propagate")
    dx, z, du = float(d["A_dx"]), float(d["A_z"]), float(d["A_du"])
    args = (d["A_amp"], d["A_opd"], None, d["A_wls"], d["A_wts"], (dx, dx), z, du, tuple(d["A_shape"]), None, int(d["A_oversample"]))
    kw = dict(wf_tilt=list(d["A_tilt"]))
    ref = oc.psf(*args, **kw)
    import types
    shim = types.SimpleNamespace(fourier=oc)            # the oracle module plays lentil.fourier: it owns dft2 / idft2
    n0 = lentil.device.launch_count()
    patch.enable(shim)
    try:
        assert oc.dft2 is lentil.fourier.dft2
        got = oc.psf(*args, **kw)
    finally:
        patch.disable(shim)
    assert oc.dft2 is not lentil.fourier.dft2
    assert lentil.device.launch_count() - n0 >= len(d["A_wls"])      # one K2a launch group per wavelength
    assert peak_err(got, ref) <= TOL64 and peak_err(got, d["A_img"]) <= TOL64


def test_batch_normalises_device_inputs_and_validates_out():
    # caller-supplied device tensors are normalised (dtype / layout) instead of being read as dense float64
    import torch
    rng = np.random.default_rng(21)
    mask = synth.circle((96, 96), 45)
    amp = synth.normalize_power(mask)
    opds = np.stack([synth.zernike_opd(mask, rng.normal(size=8) * 20e-9, first=4) for _ in range(3)])
    dx, z, du = 1 / 90, 20.0, 5e-6
    wls, wts = [6e-7, 7e-7], [0.5, 0.5]
    p = lentil.Pupil(amplitude=amp, opd=np.zeros((96, 96)), pixelscale=dx, focal_length=z)
    want = lentil.propagate_dft_batch(p, wls, du, (48, 48), oversample=2, weights=wts, opds=opds)
    dev = lentil.device.device()
    f32 = torch.from_numpy(opds).to(dev, torch.float32)                         # wrong dtype
    got = lentil.propagate_dft_batch(p, wls, du, (48, 48), oversample=2, weights=wts, opds=f32)
    assert peak_err(got, want) <= 1e-5                                           # float32 OPDs: rounded inputs, not garbage
    wide = torch.zeros(3, 96, 200, dtype=torch.float64, device=dev)
    wide[:, :, :96] = torch.from_numpy(opds).to(dev)
    got = lentil.propagate_dft_batch(p, wls, du, (48, 48), oversample=2, weights=wts, opds=wide[:, :, :96])   # non-contiguous view
    assert peak_err(got, want) <= 1e-14
    # out: accumulate-into semantics, and rejection of buffers the kernels could not address
    out = torch.full((3, 96, 96), 2.0, dtype=torch.float64, device=dev)
    res = lentil.propagate_dft_batch(p, wls, du, (48, 48), oversample=2, weights=wts, opds=opds, out=out, return_device=True)
    assert res.data_ptr() == out.data_ptr()
    assert peak_err(lentil.device.to_host(out) - 2.0, want) <= 1e-13
    for bad in (torch.zeros(3, 96, 96, dtype=torch.float32, device=dev), torch.zeros(3, 96, 192, dtype=torch.float64, device=dev)[:, :, ::2],
                torch.zeros(2, 96, 96, dtype=torch.float64, device=dev), np.zeros((3, 96, 96))):
        with pytest.raises(ValueError):
            lentil.propagate_dft_batch(p, wls, du, (48, 48), oversample=2, weights=wts, opds=opds, out=bad)
