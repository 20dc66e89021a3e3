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
