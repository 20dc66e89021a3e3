"""The REAL andykee/lentil (staged under oracle/_ref by oracle/build_ref.sh; it travels to the GPU box with the
snapshot) as the checker on hardware: lentil_b200.patch.enable() reroutes the real package through the CUDA library
and its own user loop (docs/user/diffraction.rst:154-158, lentil/propagate.py:147-242) must give the same PSF;
lentil_b200's mirror classes are compared with lentil's on identical inputs (fit_tilt supports, rebin errors)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import ref_loader  # noqa: E402
from conftest import TOL64, peak_err  # noqa: E402

pytestmark = pytest.mark.gpu

import lentil_b200  # noqa: E402
from lentil_b200 import patch, synth  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    mod = ref_loader.reference()
    if mod is None:
        pytest.skip("no staged reference under oracle/_ref (run oracle/build_ref.sh where /root/reference exists)")
    return mod


def _model(n=256, radius=120, seed=3):
    mask = synth.annulus((n, n), radius, 0.3)
    amp = synth.normalize_power(mask)
    opd = synth.zernike_opd(mask, np.random.default_rng(seed).normal(size=10) * 40e-9)
    return amp, opd


def _ref_psf(ref, amp, opd, wls, wts, dx, z, du, shape, oversample, tilt=None):
    p = ref.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    img = np.zeros((shape[0] * oversample, shape[1] * oversample))
    for wl, wt in zip(wls, wts):
        w = ref.Wavefront(wl) * p
        if tilt is not None:
            w = w * ref.Tilt(x=tilt[0], y=tilt[1])
        w = ref.propagate_dft(w, pixelscale=du, shape=shape, oversample=oversample)
        img = w.insert(img, wt)
    return img


def test_patch_fourier_level_runs_lentils_own_loop_on_the_gpu(ref):
    amp, opd = _model()
    dx, z, du = 1 / 240, 20.0, 5e-6
    wls, wts = [550e-9, 700e-9, 850e-9], [0.3, 0.5, 0.2]
    want = _ref_psf(ref, amp, opd, wls, wts, dx, z, du, (128, 128), 2, tilt=(3e-6, -2e-6))      # numpy lentil, untouched
    n0 = lentil_b200.device.launch_count()
    patch.enable(ref)                                                                      # level='fourier'
    try:
        assert ref.fourier.dft2 is lentil_b200.fourier.dft2
        got = _ref_psf(ref, amp, opd, wls, wts, dx, z, du, (128, 128), 2, tilt=(3e-6, -2e-6))     # same code, dft2 on the GPU
    finally:
        patch.disable(ref)
    assert ref.fourier.dft2 is not lentil_b200.fourier.dft2
    assert lentil_b200.device.launch_count() - n0 >= 3 * 2                                 # the transforms did run in our kernels
    assert peak_err(got, want) <= TOL64


def test_patch_path_level_swaps_propagate_dft(ref):
    amp, opd = _model(seed=4)
    dx, z, du = 1 / 240, 20.0, 5e-6
    want = _ref_psf(ref, amp, opd, [650e-9], [1.0], dx, z, du, (96, 128), 3)
    patch.enable(ref, level='path')
    try:
        assert ref.propagate_dft is lentil_b200.propagate_dft and ref.propagate.propagate_dft is lentil_b200.propagate_dft
        p = lentil_b200.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)       # path level: lentil_b200 objects
        w = ref.propagate_dft(lentil_b200.Wavefront(650e-9) * p, pixelscale=du, shape=(96, 128), oversample=3)
        got = w.insert(np.zeros((288, 384)), 1.0)
        small = ref.rebin(np.arange(36.0).reshape(6, 6), 3)
    finally:
        patch.disable(ref)
    assert peak_err(got, want) <= TOL64
    assert np.array_equal(small, ref.rebin(np.arange(36.0).reshape(6, 6), 3))


def test_mirror_classes_match_the_real_package(ref):
    """the same user code against lentil and against lentil_b200: Field offsets, shapes and the PSF agree"""
    amp, opd = _model(n=200, radius=90, seed=6)
    dx, z, du = 1 / 180, 15.0, 4e-6
    out = {}
    for name, mod in (("ref", ref), ("ours", lentil_b200)):
        p = mod.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
        w = mod.Wavefront(600e-9) * p
        w = w * mod.Tilt(x=-5e-6, y=2.5e-6)
        w = mod.propagate_dft(w, pixelscale=du, shape=(64, 80), prop_shape=(48, 48), oversample=2)
        out[name] = (w.intensity, [tuple(int(v) for v in f.offset) for f in w.data], w.shape)
    assert out["ours"][1] == out["ref"][1] and tuple(out["ours"][2]) == tuple(out["ref"][2])
    assert peak_err(out["ours"][0], out["ref"][0]) <= TOL64


@pytest.mark.parametrize("case", ["mask_scalar_amplitude", "segments_scalar_amplitude", "segments_array_amplitude"])
def test_fit_tilt_supports_match_the_real_package(ref, case):
    """lentil/plane.py:522-611: the fit runs over the MASK, also when the amplitude is a scalar"""
    rng = np.random.default_rng(12)
    n = 96
    rr, cc = np.mgrid[:n, :n]
    if case == "mask_scalar_amplitude":
        mask = synth.circle((n, n), 40).astype(bool)
        amp = 1.0
    else:
        left = synth.circle((n, n), 20, shift=(0, -24)).astype(bool)
        right = synth.circle((n, n), 18, shift=(5, 22)).astype(bool)
        mask = np.stack([left, right])
        amp = 1.0 if case == "segments_scalar_amplitude" else mask.sum(axis=0).astype(float)
    flat = mask if mask.ndim == 2 else mask.any(axis=0)
    opd = (3e-7 * (rr - n / 2) / n - 5e-7 * (cc - n / 2) / n + 2e-8 * rng.normal(size=(n, n))) * flat
    if mask.ndim == 3:
        opd = opd + 4e-7 * (rr - n / 2) / n * mask[1]
    a = ref.Pupil(amplitude=amp, opd=opd.copy(), mask=mask, pixelscale=1 / 80, focal_length=10.0).fit_tilt()
    b = lentil_b200.Pupil(amplitude=amp, opd=opd.copy(), mask=mask, pixelscale=1 / 80, focal_length=10.0).fit_tilt()
    assert len(a.tilt) == len(b.tilt) == (1 if mask.ndim == 2 else 2)
    for ta, tb in zip(a.tilt, b.tilt):
        assert abs(ta.x - tb.x) <= 1e-9 * max(abs(ta.x), 1e-12) + 1e-18
        assert abs(ta.y - tb.y) <= 1e-9 * max(abs(ta.y), 1e-12) + 1e-18
    assert np.max(np.abs(np.asarray(a.opd) - np.asarray(b.opd))) <= 1e-9 * np.max(np.abs(opd))


def test_rebin_rejects_shapes_the_factor_does_not_divide(ref):
    img = np.arange(35.0).reshape(5, 7)
    with pytest.raises(ValueError):
        ref.rebin(img, 2)
    with pytest.raises(ValueError):
        lentil_b200.rebin(img, 2)
    ints = np.arange(36, dtype=np.int32).reshape(6, 6)
    assert lentil_b200.rebin(ints, 2).dtype == ref.rebin(ints, 2).dtype
    assert np.array_equal(lentil_b200.rebin(ints, 2), ref.rebin(ints, 2))


def test_oracle_port_is_the_real_package_bit_for_bit(ref):
    """the oracle port that the other GPU tests use, re-pinned on the GPU box's own numpy / BLAS"""
    import lentil_oracle as oc
    amp, opd = _model(seed=9)
    dx, z, du = 1 / 240, 20.0, 5e-6
    want = _ref_psf(ref, amp, opd, [650e-9], [1.0], dx, z, du, (64, 64), 2, tilt=(4e-6, -2e-6))
    got = oc.psf(amp, opd, None, [650e-9], [1.0], (dx, dx), z, du, (64, 64), None, 2, wf_tilt=[4e-6, -2e-6])
    assert np.array_equal(got, want)
