"""The reference's own property tests for the path (lentil tests/test_propagate*.py, test_plane.py, test_wavefront.py),
restated against lentil_b200 with seeded inputs.  Each test names the reference test it mirrors."""
import numpy as np
import pytest

import lentil_b200 as lentil
import lentil_oracle as oc
from lentil_b200 import synth, helper
from conftest import peak_err, TOL64

pytestmark = pytest.mark.gpu


def _pupil(focal_length, diameter, shape, radius, coeffs=None):
    # tests/fixtures: normalize_power(circle), pixelscale = diameter / (2 radius), Zernike OPD
    amp = synth.normalize_power(synth.circle(shape, radius))
    opd = synth.zernike_opd(amp > 0, coeffs) if coeffs is not None else 0
    return lentil.Pupil(amplitude=amp, opd=opd, pixelscale=diameter / (2 * radius), focal_length=focal_length)


@pytest.mark.parametrize("oversample", [1, 2])
def test_amplitude_normalize_power(oversample):
    # test_propagate.py:8-23 — a unit-power pupil puts (almost) unit power on a big enough detector
    p = _pupil(10, 1, (256, 256), 120)
    w = lentil.Wavefront(wavelength=650e-9)
    w *= p                                                   # test_plane.py:107 (imul)
    w = lentil.propagate_dft(w, shape=(64, 64), pixelscale=5e-6, oversample=oversample)
    total = np.sum(w.intensity)
    assert 0.95 <= total <= 1


@pytest.mark.parametrize("coeffs", [None, [0, 1e-6, 2e-6]])
def test_propagate_mask(coeffs):
    # test_propagate_mask.py:4-37 — computing only the masked window == masking the full PSF
    p = _pupil(10, 1, (256, 256), 120, coeffs)
    mask = np.zeros((256, 256))
    mask[128 + 20 - 32:128 + 20 + 32, 128 - 30 - 32:128 - 30 + 32] = 1      # rectangle((256,256), 64, 64, shift=(20,-30))
    psf = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, shape=128, pixelscale=5e-6, oversample=2).intensity
    w_mask = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, shape=128, pixelscale=5e-6, oversample=2, mask=mask)
    assert len(w_mask.data) == 1 and tuple(w_mask.data[0].shape) == (64, 64)
    assert peak_err(w_mask.intensity, psf * mask) <= TOL64
    with pytest.raises(ValueError):
        lentil.propagate_dft(lentil.Wavefront(650e-9) * p, shape=128, pixelscale=5e-6, oversample=2, mask=np.ones((7, 9)))


def test_propagate_slice_one_and_multi():
    # test_propagate_slice.py:9-100 — transforming the bounding boxes with their offsets == transforming the full array
    rng = np.random.default_rng(11)
    n, wavelength, du, z = 256, 500e-9, 5e-6, 20.0
    dx = 1 / 102
    alpha, oversample, shape = (dx * du) / (wavelength * z), 5, 64
    amps, opds = [], []
    for shift in [(0, int(-0.3 * n)), (0, int(0.3 * n))]:
        a = synth.circle((n, n), n // 5, shift=shift)
        amps.append(a)
        opds.append(synth.zernike_opd(a > 0, np.r_[0, 0, 0, rng.uniform(-1, 1, 8) * 100e-9]))
    amp, opd = amps[0] + amps[1], opds[0] + opds[1]
    F = lentil.fourier.dft2(amp * np.exp(-2j * np.pi * opd / wavelength), alpha / oversample, shape=shape * oversample)
    F_slc = 0
    for a, o in zip(amps, opds):
        slc = helper.boundary_slice(a)
        assert slc == oc.boundary_slice(a)
        ofst = helper.slice_offset(slc, a.shape)
        F_slc = F_slc + lentil.fourier.dft2(a[slc] * np.exp(-2j * np.pi * o[slc] / wavelength), alpha / oversample,
                                            shape=shape * oversample, offset=ofst)
    assert peak_err(F_slc, F) <= TOL64
    # one aperture, shifted off centre (test_propagate_slice_one)
    a, o = amps[0], opds[0]
    slc = helper.boundary_slice(a)
    F1 = lentil.fourier.dft2(a * np.exp(-2j * np.pi * o / wavelength), alpha / oversample, shape=shape * oversample)
    F1s = lentil.fourier.dft2(a[slc] * np.exp(-2j * np.pi * o[slc] / wavelength), alpha / oversample,
                              shape=shape * oversample, offset=helper.slice_offset(slc, a.shape))
    assert peak_err(F1s, F1) <= TOL64


def test_propagate_tilt_moves_the_centroid_like_the_reference_chain():
    # test_propagate.py:69-131 — a tilt in the OPD moves the PSF by z * tilt / du * oversample pixels, and fit_tilt
    # (which turns that tilt into a window shift) lands the PSF on the same spot
    oversample, du, npix, z = 10, 5e-6, 64, 10.0
    amp = synth.normalize_power(synth.circle((256, 256), 120))
    dx = 1 / 240
    rr, cc = np.meshgrid(np.arange(256) - 128, np.arange(256) - 128, indexing='ij')
    tx, ty = 2.1e-6, -1.3e-6                                       # OPD slopes (m per m) along +c and +r
    opd = (tx * cc * dx + ty * rr * dx) * (amp > 0)
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    shifts = []
    for plane in (p, p.fit_tilt(inplace=False)):
        psf = lentil.propagate_dft(lentil.Wavefront(650e-9) * plane, shape=npix, pixelscale=du, oversample=oversample).intensity
        psf = psf / psf.max()
        psf[psf < 0.2] = 0
        r, c = np.indices(psf.shape)
        centroid = np.array([np.sum(r * psf), np.sum(c * psf)]) / np.sum(psf)
        shifts.append(centroid - npix // 2 * oversample)
    expect = np.array([ty, tx]) * z / du * oversample
    assert np.all(np.abs(np.abs(shifts[0]) - np.abs(expect)) / oversample < 0.2)       # magnitude as in the reference test
    assert np.all(np.abs(shifts[0] - shifts[1]) / oversample < 0.05)                  # OPD tilt == fitted Tilt


def test_fit_tilt_inplace_semantics():
    # test_plane.py:74-80
    amp = synth.circle((64, 64), 28)
    opd = np.fromfunction(lambda r, c: 1e-7 * r - 2e-7 * c, (64, 64)) * (amp > 0)
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 56, focal_length=10.0)
    q = p.fit_tilt(inplace=False)
    assert q is not p and len(p.tilt) == 0 and len(q.tilt) == 1
    assert np.array_equal(p.opd, opd)
    r = p.fit_tilt(inplace=True)
    assert r is p and len(p.tilt) == 1
    resid = p.opd[amp > 0]                                 # tip/tilt removed, piston stays (lentil/plane.py:597-609)
    assert np.max(resid) - np.min(resid) < 1e-12


def test_wavefront_plane_products():
    # test_plane.py:83-116, test_wavefront.py:11-13 — w * p, p * w and w *= p are the same product
    amp = synth.circle((32, 32), 12)
    opd = synth.zernike_opd(amp > 0, np.array([0, 0, 0, 3e-8, -2e-8]))
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 24, focal_length=5.0)
    w1 = lentil.Wavefront(650e-9) * p
    w2 = p * lentil.Wavefront(650e-9)
    w3 = lentil.Wavefront(650e-9)
    w3 *= p
    slc = helper.boundary_slice(amp)
    ref = amp[slc] * np.exp(2j * np.pi * opd[slc] / 650e-9)
    for w in (w1, w2, w3):
        assert len(w.data) == 1 and w.focal_length == 5.0 and w.ptype == lentil.pupil
        assert peak_err(w.data[0].data, ref) <= 1e-14
        assert tuple(int(v) for v in w.data[0].offset) == tuple(helper.slice_offset(slc, amp.shape))


def test_plane_defaults_alias_and_hooks():
    # test_plane.py:18-58
    p = lentil.Plane()
    assert p.pixelscale is None and np.all(p.amplitude == 1) and np.all(p.opd == 0) and p.mask == p.amplitude
    assert lentil.Plane(amp=10).amplitude == 10
    with pytest.raises(AttributeError):
        lentil.Plane(amplitude=10, amp=10)

    class Hooked(lentil.Plane):
        def __amp__(self):
            return np.array(1)

        def __opd__(self):
            return np.array(2)

        def __mask__(self):
            return np.array(3)

    h = Hooked()
    assert h.amplitude == 1 and h.opd == 2 and h.mask == 3


def test_propagate_no_output_and_wrong_ptype():
    # test_propagate.py:134-141, test_wavefront.py:16-19
    amp = synth.normalize_power(synth.circle((63, 63), 28))
    opd = np.fromfunction(lambda r, c: 3e-5 * c / 56, (63, 63)) * (amp > 0)
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 56, focal_length=10.0).fit_tilt(inplace=False)
    w = lentil.propagate_dft(lentil.Wavefront(650e-9) * p, shape=(16, 16), pixelscale=5e-6, oversample=2)
    assert len(w.data) == 0 and np.all(w.intensity == 0) and w.intensity.shape == (32, 32)
    with pytest.raises(TypeError):
        lentil.propagate_dft(lentil.Wavefront(650e-9), pixelscale=5e-6, shape=(8, 8))
