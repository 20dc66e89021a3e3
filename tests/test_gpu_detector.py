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
