"""K2b parity: complex64 / 3xTF32 (tcgen05) matrix Fourier transform against the oracle.
Gate: <= 1e-5 peak-normalised (BASELINE.json north_star); observed ~1e-6."""
import numpy as np
import pytest

import lentil_b200 as lentil
import lentil_oracle as oc
from conftest import peak_err

pytestmark = pytest.mark.gpu
TOL32 = 1e-5


@pytest.fixture(params=["tcgen05", "czt"], autouse=True)
def c64_execution(request):
    """Every test of this file runs under both executions of the complex64 mode: the folded 3xTF32 transform on tcgen05
    (process default LFD_MFT_FOLDED) and the FP32 chirp-z row transform (LFD_MFT_AUTO, the default)."""
    from lentil_b200 import _lib
    L = _lib.lib()
    L.lfd_set_mft_variant(1 if request.param == "tcgen05" else 3)
    d = (_lib.MftDesc * 1)()
    d[0].m = d[0].n = d[0].M = d[0].N = 64
    assert L.lfd_mft_c64_execution(d, 1) == (1 if request.param == "tcgen05" else 2)
    yield request.param
    L.lfd_set_mft_variant(3)

CASES = [
    (10, 10, 10, 10, 0.1, (0, 0), (0, 0), True),
    (11, 13, 17, 9, (1 / 11, 1 / 13), (0.3, -1.7), (2, -3), True),
    (33, 47, 64, 20, (0.004, 0.003), (1.5, -0.25), (3, -4), False),
    (241, 241, 256, 256, 1.3e-3, (0.4, 0.6), (0, 0), True),
    (300, 200, 130, 260, (0.002, 0.0031), (13.4, -7.6), (-40, 25), True),
    (501, 501, 486, 499, 3.846e-4, (13.4, 7.6), (0, 0), True),
    (1001, 1001, 1024, 1024, 1 / 2048, (0.3, -0.4), (0, 0), True),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}to{c[2]}x{c[3]}")
def test_dft2_c64_against_oracle(case):
    m, n, M, N, alpha, shift, offset, unitary = case
    rng = np.random.default_rng(m + 3 * n + 5 * M)
    f = (rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))).astype(np.complex64)
    F = lentil.fourier.dft2_c64(f, alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary)
    assert F.shape == (M, N) and F.dtype == np.complex64
    ref = oc.dft2(f.astype(np.complex128), alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary)
    assert peak_err(F, ref) <= TOL32
    g = lentil.fourier.idft2_c64(f, alpha, shape=(M, N), shift=shift, unitary=unitary)
    assert peak_err(g, oc.idft2(f.astype(np.complex128), alpha, shape=(M, N), shift=shift, unitary=unitary)) <= TOL32


def test_c64_psf_error_on_a_pupil():
    # the north-star metric: peak-normalised PSF (intensity) error of the 3xTF32 path
    from lentil_b200 import synth
    mask = synth.annulus((512, 512), 250)
    amp = synth.normalize_power(mask)
    opd = synth.zernike_opd(mask, np.random.default_rng(0).normal(size=15) * 30e-9)
    slc = lentil.helper.boundary_slice(mask)
    f = amp[slc] * np.exp(2j * np.pi * opd[slc] / 650e-9)
    alpha = (1 / 500) * 5e-6 / (650e-9 * 20.0 * 2)
    F = lentil.fourier.dft2_c64(f.astype(np.complex64), alpha, shape=(512, 512))
    ref = oc.dft2(f, alpha, shape=(512, 512))
    I, Iref = np.abs(F.astype(np.complex128)) ** 2, np.abs(ref) ** 2
    assert np.max(np.abs(I - Iref)) / np.max(Iref) <= TOL32


def test_batch_pipeline_c64_against_oracle():
    """K1 (complex64 phasors) -> K2b -> K3 (float64 accumulation) through propagate_dft_batch."""
    from lentil_b200 import synth
    rng = np.random.default_rng(5)
    mask = synth.annulus((256, 256), 120)
    amp = synth.normalize_power(mask)
    opd = synth.zernike_opd(mask, rng.normal(size=12) * 30e-9)
    dx, z, du = 1 / 240, 20.0, 5e-6
    wls = np.linspace(500e-9, 900e-9, 6)
    wts = np.full(6, 1 / 6)
    tilts = [[0.0, 0.0], [7e-6, -3e-6]]
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    stack = lentil.propagate_dft_batch(p, wls, du, (128, 128), oversample=2, weights=wts, tilts=tilts, precision='c64')
    ref64 = lentil.propagate_dft_batch(p, wls, du, (128, 128), oversample=2, weights=wts, tilts=tilts)
    assert stack.shape == (2, 256, 256) and stack.dtype == np.float64
    for k, t in enumerate(tilts):
        ref = oc.psf(amp, opd, None, wls, wts, (dx, dx), z, du, (128, 128), None, 2, wf_tilt=t)
        assert peak_err(stack[k], ref) <= TOL32
        assert peak_err(ref64[k], ref) <= 1e-10
    with pytest.raises(ValueError):
        lentil.propagate_dft_batch(p, wls, du, (128, 128), precision='fp16')


def test_fused_pupil_prep_c64_single_field_point_and_segments():
    """With one field point K1 is fused into K2b's fold kernel (lfd_mft_c64x3_from_pupil): monolithic pupils also take
    the |F|^2 epilogue, segmented ones keep complex64 windows for the coherent merge in K3; a Monte-Carlo OPD stack
    selects the OPD plane per realisation."""
    from lentil_b200 import synth
    rng = np.random.default_rng(9)
    mask = synth.annulus((200, 200), 95, 0.25)
    amp = synth.normalize_power(mask)
    opd = synth.zernike_opd(mask, rng.normal(size=10) * 30e-9)
    dx, z, du = 1 / 190, 15.0, 5e-6
    wls, wts = np.linspace(550e-9, 800e-9, 5), np.array([0.1, 0.2, 0.4, 0.2, 0.1])
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    got = lentil.propagate_dft_batch(p, wls, du, (100, 100), oversample=2, weights=wts, precision='c64')
    ref = oc.psf(amp, opd, None, wls, wts, (dx, dx), z, du, (100, 100), None, 2)
    assert got.shape == ref.shape and peak_err(got, ref) <= TOL32
    # Monte-Carlo realisations (one OPD plane each)
    opds = np.stack([synth.zernike_opd(mask, rng.normal(size=10) * 25e-9) for _ in range(3)])
    mc = lentil.propagate_dft_batch(p, wls[:2], du, (100, 100), oversample=2, weights=wts[:2], opds=opds, precision='c64')
    for r in range(3):
        assert peak_err(mc[r], oc.psf(amp, opds[r], None, wls[:2], wts[:2], (dx, dx), z, du, (100, 100), None, 2)) <= TOL32
    # segmented pupil: windows overlap and must interfere
    cube = synth.hex_segments(1, 30, 2)
    sa = synth.normalize_power(cube.sum(axis=0).astype(float))
    so = np.zeros(sa.shape)
    for s in range(cube.shape[0]):
        so += synth.zernike_opd(cube[s], rng.uniform(-1, 1, 3) * np.array([40e-9, 3e-7, 3e-7]))
    ps = lentil.Pupil(amplitude=sa, opd=so, mask=cube, pixelscale=1 / 160, focal_length=12.0)
    gs = lentil.propagate_dft_batch(ps, wls[:3], du, (64, 64), oversample=2, weights=wts[:3], precision='c64')
    rs = oc.psf(sa, so, cube, wls[:3], wts[:3], (1 / 160, 1 / 160), 12.0, du, (64, 64), None, 2)
    assert peak_err(gs, rs) <= TOL32
