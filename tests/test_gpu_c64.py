"""K2b parity: complex64 / 3xTF32 (tcgen05) matrix Fourier transform against the oracle.
Gate: <= 1e-5 peak-normalised (BASELINE.json north_star); observed ~1e-6."""
import numpy as np
import pytest

import lentil_b200 as lentil
import lentil_oracle as oc
from conftest import peak_err

pytestmark = pytest.mark.gpu
TOL32 = 1e-5

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
