"""Detector-side sampling on the device (next rows of SURVEY.md section 8(f)): rebin and the pixel MTF,
against golden vectors produced by the reference (lentil.rebin, lentil.detector.pixel)."""
import numpy as np
import pytest

import lentil_b200 as lentil
from conftest import peak_err

pytestmark = pytest.mark.gpu


def test_rebin_golden(golden):
    d = golden("detector")
    assert np.array_equal(lentil.rebin(d["img"], 3), d["rebin3"])                 # same summation order: exact
    assert np.array_equal(lentil.rebin(d["cube"], 2), d["rebin_cube2"])
    dev = lentil.device.to_dev(d["img"])
    out = lentil.rebin(dev, 3)
    assert lentil.device.is_dev(out) and np.array_equal(lentil.device.to_host(out), d["rebin3"])
    with pytest.raises(ValueError):
        lentil.rebin(d["img"].astype(complex), 2)


def test_pixel_mtf_golden(golden):
    d = golden("detector")
    assert peak_err(lentil.detector.pixel(d["img"], 2), d["pixel2"]) <= 1e-12
    assert peak_err(lentil.detector.pixel(d["img"][:45, :45], 3), d["pixel3"]) <= 1e-12


def test_pixel_preserves_radiometry_and_psf_chain():
    # docs: pixel MTF and rebinning conserve the total (lentil/detector.py:193-199)
    from lentil_b200 import synth
    mask = synth.circle((128, 128), 60)
    p = lentil.Pupil(amplitude=synth.normalize_power(mask), opd=np.zeros((128, 128)), pixelscale=1 / 120, focal_length=20.0)
    psf = lentil.propagate_dft_batch(p, [650e-9], 5e-6, (64, 64), oversample=4, return_device=True)
    mtf = lentil.detector.pixel(psf, oversample=4)
    det = lentil.rebin(mtf, 4)
    tot = float(psf.sum())
    assert abs(float(mtf.sum()) - tot) <= 1e-12 * tot and abs(float(det.sum()) - tot) <= 1e-12 * tot
    assert tuple(det.shape) == (64, 64)


def test_opd_synthesis_matches_einsum_and_feeds_the_batch():
    # docs/user/wavefront_error.rst:118-135: z = np.einsum('ijk,i->jk', basis, coeff)
    from lentil_b200 import synth
    import lentil_oracle as oc
    rng = np.random.default_rng(31)
    mask = synth.circle((96, 96), 45)
    rows, cols = np.nonzero(mask)
    rr, cc = np.meshgrid(np.arange(96) - 47.5, np.arange(96) - 47.5, indexing='ij')
    rho, theta = np.hypot(rr, cc) / 46.0, np.arctan2(rr, cc)
    basis = np.stack([synth.zernike(j, rho, theta) * mask for j in range(4, 37)])       # 33 modes
    coeffs = rng.normal(size=(6, 33)) * 20e-9
    base = rng.normal(size=(96, 96)) * 1e-9 * mask
    got = lentil.detector.synthesize_opd(basis, coeffs, base=base, return_device=False)
    ref = np.einsum('ijk,ri->rjk', basis, coeffs) + base
    assert got.shape == (6, 96, 96)
    assert peak_err(got, ref) <= 1e-14
    one = lentil.detector.synthesize_opd(basis, coeffs[2], return_device=False)
    assert peak_err(one, np.einsum('ijk,i->jk', basis, coeffs[2])) <= 1e-14
    # more than 64 terms are chained
    big_basis = rng.normal(size=(70, 16, 16))
    big_c = rng.normal(size=(3, 70))
    assert peak_err(lentil.detector.synthesize_opd(big_basis, big_c, return_device=False),
                    np.einsum('ijk,ri->rjk', big_basis, big_c)) <= 1e-13
    # device-resident OPD stack straight into the Monte-Carlo batch
    opds = lentil.detector.synthesize_opd(basis, coeffs)
    p = lentil.Pupil(amplitude=synth.normalize_power(mask), opd=np.zeros((96, 96)), pixelscale=1 / 90, focal_length=20.0)
    wls, wts = np.array([6e-7, 7e-7]), np.array([0.5, 0.5])
    stack = lentil.propagate_dft_batch(p, wls, 5e-6, (48, 48), oversample=2, weights=wts, opds=opds)
    ref = oc.psf(p.amplitude, np.einsum('ijk,i->jk', basis, coeffs[4]), None, wls, wts, (1 / 90, 1 / 90), 20.0, 5e-6,
                 (48, 48), None, 2)
    assert peak_err(stack[4], ref) <= 1e-10


def test_power_spectrum_golden(golden):
    # lentil/wfe.py:8-70: the two FFTs of the noise shaping run as centred K2a transforms (even and odd grids)
    d = golden("power_spectrum")
    for i in range(int(d["n"])):
        mask, ref = d[f"c{i}_mask"], d[f"c{i}_opd"]
        opd = lentil.power_spectrum(mask, 1 / (2 * int(d[f"c{i}_radius"])), 30e-9, 5, 3, seed=int(d[f"c{i}_seed"]))
        assert opd.shape == ref.shape
        assert np.max(np.abs(opd - ref)) <= 1e-12 * np.max(np.abs(ref))
        rms = np.sqrt(np.sum(opd ** 2) / np.count_nonzero(opd))
        assert abs(rms - 30e-9) <= 1e-15
    with pytest.raises(ValueError):
        lentil.power_spectrum(np.ones((8, 9)), 1.0, 1e-9, 5, 3, seed=0)


def test_rescale_golden(golden):
    # lentil.rescale (lentil/util.py:261-347): every spline order, the four documented extension modes, complex
    # input, explicit shape/mask, non-unitary — against vectors produced by the reference
    from test_oracle_golden import rescale_golden_cases
    d = golden("rescale")
    for name, key, args, kw in rescale_golden_cases(d):
        got = lentil.rescale(d[key], *args, **kw)
        assert got.dtype == d[name].dtype, name
        assert peak_err(got, d[name]) <= 1e-12, name
    got = lentil.rescale(d["img"], 0.5, shape=40, mask=np.ones_like(d["img"]), unitary=False)
    assert peak_err(got, d["shape40_ones_nonunitary"]) <= 1e-12


def test_pixelate_golden_and_device_chain(golden):
    d = golden("rescale")
    for os_ in (3, 4):
        assert peak_err(lentil.detector.pixelate(d["psf"], os_), d[f"pixelate{os_}"]) <= 1e-12
    dev = lentil.device.to_dev(d["psf"])
    out = lentil.detector.pixelate(dev, 3)
    assert lentil.device.is_dev(out) and peak_err(lentil.device.to_host(out), d["pixelate3"]) <= 1e-12


def test_rescale_matches_oracle_at_psf_size():
    # a 1024^2 oversampled PSF brought to native sampling (BASELINE configs[1] detector), against the oracle
    import lentil_oracle as oc
    rng = np.random.default_rng(12)
    img = rng.random((1024, 1024)) ** 6
    img[:, :17] = 0
    for scale, kw in ((0.5, {}), (1 / 3, {}), (0.25, dict(order=5)), (0.3, dict(mode="constant"))):
        assert peak_err(lentil.rescale(img, scale, **kw), oc.rescale(img, scale, **kw)) <= 1e-12


def test_rescale_small_and_degenerate_shapes():
    import lentil_oracle as oc
    rng = np.random.default_rng(13)
    for shape in ((2, 2), (3, 7), (1, 5), (13, 2)):
        img = rng.random(shape) + 0.1
        for scale in (0.5, 1.0, 2.5):
            for mode in ("nearest", "reflect", "constant"):
                with np.errstate(all="ignore"):
                    ref = oc.rescale(img, scale, mode=mode)
                got = lentil.rescale(img, scale, mode=mode)
                if not np.all(np.isfinite(ref)):        # a one-sample axis in 'constant' mode: sum(out) = 0, NaN in the reference too
                    assert np.array_equal(np.isfinite(got), np.isfinite(ref))
                    continue
                assert peak_err(got, ref) <= 1e-12, (shape, scale, mode)
