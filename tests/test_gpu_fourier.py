"""K2a parity: lentil_b200.fourier.dft2/idft2 (C ABI -> CUDA) against the oracle, the reference's
golden vectors and the reference's own property tests (tests/test_fourier.py).  Tolerance: the
FP64 gate of BASELINE.json, 1e-10 peak-normalised (observed ~1e-14)."""

import numpy as np
import pytest

import lentil_b200 as lentil
import lentil_oracle as oc
from lentil_b200 import _lib
from conftest import peak_err, TOL64

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures('mft_variant')]


def test_golden_vectors(golden):
    d = golden("dft2")
    for i in range(int(d["ncase"])):
        f, alpha = d[f"c{i}_f"], tuple(d[f"c{i}_alpha"])
        shape, shift, offset = tuple(d[f"c{i}_shape"]), tuple(d[f"c{i}_shift"]), tuple(d[f"c{i}_offset"])
        unitary = bool(d[f"c{i}_unitary"])
        F = lentil.fourier.dft2(f, alpha, shape=shape, shift=shift, offset=offset, unitary=unitary)
        assert peak_err(F, d[f"c{i}_F"]) <= TOL64
        iF = lentil.fourier.idft2(f, alpha, shape=shape, shift=shift, unitary=unitary)
        assert peak_err(iF, d[f"c{i}_iF"]) <= TOL64


@pytest.mark.parametrize("m,n", [(10, 10), (11, 11), (10, 11)])
def test_dft2_is_shifted_fft(m, n):
    rng = np.random.default_rng(m * 100 + n)
    f = rng.random((m, n)) + 1j * rng.random((m, n))
    F = lentil.fourier.dft2(f, [1 / m, 1 / n], unitary=False)
    assert np.allclose(F, np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(f))))
    g = lentil.fourier.idft2(f, [1 / m, 1 / n], unitary=False)
    assert np.allclose(g, np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(f))))


def test_dft2_inverse_roundtrip():
    rng = np.random.default_rng(1)
    f = rng.random((10, 10)) + 1j * rng.random((10, 10))
    F = lentil.fourier.dft2(f, 1 / 10, unitary=False)
    assert np.allclose(f, lentil.fourier.idft2(F, 1 / 10, unitary=False))


def test_dft2_unitary_parseval():
    rng = np.random.default_rng(2)
    f = rng.random((10, 10)) + 1j * rng.random((10, 10))
    F = lentil.fourier.dft2(f, 1 / 10, unitary=True)
    assert np.allclose(np.sum(np.abs(f) ** 2), np.sum(np.abs(F) ** 2))


def test_dft2_shift_moves_centroid():
    # reference tests/test_fourier.py:68-78 (integer shifts, exact)
    rng = np.random.default_rng(3)
    n = 100
    shift = np.round(rng.uniform(-25, 25, 2))
    F = lentil.fourier.dft2(np.ones((n, n)), 1 / n, shift=shift)
    I = np.abs(F) ** 2
    rr, cc = np.indices(I.shape)
    cen = np.array([np.sum(rr * I), np.sum(cc * I)]) / np.sum(I)
    assert np.allclose(cen - n // 2, shift, atol=1e-9)


def test_dft2_offset_equals_zero_embedding():
    rng = np.random.default_rng(4)
    n, m = 10, 3
    f = np.zeros((n, n), dtype=complex)
    r, c = rng.integers(0, n - m, 2)
    f[r:r + m, c:c + m] = rng.random((m, m)) + 1j * rng.random((m, m))
    slc = lentil.helper.boundary_slice(f)
    off = lentil.helper.slice_offset(slc, f.shape)
    F = lentil.fourier.dft2(f, alpha=1 / m, shape=10)
    FF = lentil.fourier.dft2(f[slc], alpha=1 / m, shape=10, offset=off)
    assert np.allclose(F, FF)


def test_dft2_out_aliases_input_and_dtype_check():
    rng = np.random.default_rng(5)
    f = rng.normal(size=(10, 10)).astype(complex)
    expect = oc.dft2(f.copy(), 1 / 10)
    F = lentil.fourier.dft2(f, 1 / 10, out=f)
    assert F is f and peak_err(F, expect) <= TOL64
    with pytest.raises(TypeError):
        lentil.fourier.dft2(np.ones((4, 4)), 0.25, out=np.zeros((4, 4)))


def test_real_input_is_promoted():
    f = np.arange(12.0).reshape(3, 4)
    assert peak_err(lentil.fourier.dft2(f, 0.1, shape=(5, 6)), oc.dft2(f, 0.1, shape=(5, 6))) <= TOL64


CASES = [
    # m, n, M, N, alpha, shift, offset, unitary
    (1, 1, 1, 1, 0.5, (0, 0), (0, 0), True),                       # degenerate
    (1, 7, 9, 1, (0.3, 0.11), (0.5, 0), (2, -1), True),
    (3, 5, 2, 200, (0.05, 0.002), (0, 17.25), (0, 0), False),      # K < one MMA step
    (17, 4, 131, 67, 0.004, (-3.5, 2.25), (5, 6), True),           # ragged tiles everywhere
    (127, 129, 128, 64, 1 / 128, (0, 0), (0, 0), True),            # just under / over tile sizes
    (129, 65, 129, 65, 1 / 200, (0.5, -0.5), (-64, 32), True),
    (241, 241, 256, 256, 1.3e-3, (0.4, 0.6), (0, 0), True),        # cfg1 plane
    (406, 468, 251, 504, 7.7e-4, (132.4, -97.7), (-821, 640), True),  # cfg3-like window
    (501, 501, 486, 499, 3.846e-4, (13.4, 7.6), (0, 0), True),
    (300, 1, 1, 300, 1e-3, (0, 0), (0, 0), True),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}to{c[2]}x{c[3]}")
def test_against_oracle(case):
    m, n, M, N, alpha, shift, offset, unitary = case
    rng = np.random.default_rng(m * 7 + n * 13 + M)
    f = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
    F = lentil.fourier.dft2(f, alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary)
    assert F.shape == (M, N) and F.dtype == np.complex128
    assert peak_err(F, oc.dft2(f, alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary)) <= TOL64
    g = lentil.fourier.idft2(f, alpha, shape=(M, N), shift=shift, unitary=unitary)
    assert peak_err(g, oc.idft2(f, alpha, shape=(M, N), shift=shift, unitary=unitary)) <= TOL64


def test_cfg2_plane_against_oracle():
    # BASELINE config 2 plane: 1001^2 -> 1024^2, alpha = 1/2048
    rng = np.random.default_rng(0)
    f = rng.normal(size=(1001, 1001)) + 1j * rng.normal(size=(1001, 1001))
    F = lentil.fourier.dft2(f, 1 / 2048, shape=1024, shift=(0.3, -0.4))
    assert peak_err(F, oc.dft2(f, 1 / 2048, shape=1024, shift=(0.3, -0.4))) <= TOL64


def test_full_size_properties_2k_headline():
    """BASELINE '2k' headline shape (dense 1024^2 -> 1024^2, alpha 1/2048) through size-independent
    properties: linearity, Parseval on the critically sampled grid, idft2(dft2) = identity."""
    rng = np.random.default_rng(6)
    n = 1024
    a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    b = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
    Fa = lentil.fourier.dft2(a, 1 / 2048, shape=n, shift=(0.3, -0.4))
    Fb = lentil.fourier.dft2(b, 1 / 2048, shape=n, shift=(0.3, -0.4))
    Fab = lentil.fourier.dft2(2.0 * a - 3.0j * b, 1 / 2048, shape=n, shift=(0.3, -0.4))
    assert peak_err(Fab, 2.0 * Fa - 3.0j * Fb) <= TOL64
    G = lentil.fourier.dft2(a, 1 / n, shape=n, unitary=True)          # critically sampled: unitary
    assert abs(np.sum(np.abs(G) ** 2) / np.sum(np.abs(a) ** 2) - 1.0) <= 1e-12
    back = lentil.fourier.idft2(G, 1 / n, shape=n, unitary=True) * n * n   # undo the two sqrt(alpha^2) scales
    assert peak_err(back, a) <= TOL64


def test_batched_descriptors_heterogeneous_planes():
    """One lfd_mft_c128_batched call over planes of different shapes equals plane-by-plane calls."""
    import torch
    rng = np.random.default_rng(7)
    shapes = [(33, 47, 64, 20), (100, 100, 130, 257), (5, 300, 12, 9), (200, 64, 128, 128)]
    dev = lentil.device.device()
    ins, outs, refs = [], [], []
    descs = (_lib.MftDesc * len(shapes))()
    for k, (m, n, M, N) in enumerate(shapes):
        f = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
        alpha, shift, off = (0.003 + 0.001 * k, 0.002), (k * 0.7, -k * 1.3), (k, -2 * k)
        refs.append(oc.dft2(f, alpha, shape=(M, N), shift=shift, offset=off))
        ins.append(lentil.device.to_dev(f))
        outs.append(lentil.device.empty_c128(M, N))
        lentil.fourier.mft_descriptor(descs[k], ins[k], outs[k], alpha, shift, off)
    lentil.fourier.run_mft(descs, len(shapes))
    for o, r in zip(outs, refs):
        assert peak_err(lentil.device.to_host(o), r) <= TOL64


def test_strided_input_and_output_windows():
    """ldf/ldo larger than the row length: transform a sub-window in place of a bigger buffer."""
    rng = np.random.default_rng(8)
    big = rng.normal(size=(60, 80)) + 1j * rng.normal(size=(60, 80))
    sub = big[7:40, 11:61]
    dbig = lentil.device.to_dev(big)
    dout = lentil.device.empty_c128(50, 90).zero_()
    win = dout[3:43, 5:75]
    descs = (_lib.MftDesc * 1)()
    lentil.fourier.mft_descriptor(descs[0], dbig[7:40, 11:61], win, 0.01, (0.5, 0.25), (3, -4))
    lentil.fourier.run_mft(descs, 1)
    host = lentil.device.to_host(dout)
    assert peak_err(host[3:43, 5:75], oc.dft2(sub, 0.01, shape=(40, 70), shift=(0.5, 0.25), offset=(3, -4))) <= TOL64
    host[3:43, 5:75] = 0
    assert not host.any()                      # nothing written outside the window


def test_host_buffer_entry_point():
    """lfd_ctx_dft2_host: numpy pointers in, numpy pointers out (the e2e leg of bench.py)."""
    L = _lib.lib()
    ctx = L.lfd_ctx_create(lentil.device.device().index or 0)
    assert ctx
    rng = np.random.default_rng(9)
    f = rng.normal(size=(70, 90)) + 1j * rng.normal(size=(70, 90))
    F = np.empty((64, 48), dtype=np.complex128)
    rc = L.lfd_ctx_dft2_host(ctx, f.ctypes.data, 90, 70, 90, 0.01, 0.008, 64, 48, 0.5, -1.5, 2.0, 3.0, 1, 0,
                             F.ctypes.data, 48)
    _lib.check(rc)
    assert peak_err(F, oc.dft2(f, (0.01, 0.008), shape=(64, 48), shift=(0.5, -1.5), offset=(2, 3))) <= TOL64
    L.lfd_ctx_destroy(ctx)


SPARSE = ["corners", "ring", "single", "zero", "stripe"]


@pytest.mark.parametrize("kind", SPARSE)
def test_sparse_support_patterns(kind):
    """The folded kernel skips K tiles that hold no data (support map built by the fold kernel):
    inputs whose non-zeros sit in awkward places must still transform exactly."""
    rng = np.random.default_rng(hash(kind) % 1000)
    m, n, M, N = 150, 171, 96, 140
    f = np.zeros((m, n), dtype=complex)
    if kind == "corners":
        for r, c in ((0, 0), (0, n - 1), (m - 1, 0), (m - 1, n - 1)):
            f[r, c] = rng.normal() + 1j * rng.normal()
    elif kind == "ring":
        rr, cc = np.indices((m, n))
        rad = np.hypot(rr - m // 2, cc - n // 2)
        f[(rad > 55) & (rad < 70)] = 1.0
        f *= np.exp(1j * rng.normal(size=(m, n)))
    elif kind == "single":
        f[m // 2 + 37, n // 2 - 5] = 2.0 - 1.0j
    elif kind == "stripe":
        f[:, 3] = rng.normal(size=m)
        f[m // 2, :] += 1j
    F = lentil.fourier.dft2(f, (0.004, 0.0031), shape=(M, N), shift=(2.5, -7.25), offset=(3, -11))
    ref = oc.dft2(f, (0.004, 0.0031), shape=(M, N), shift=(2.5, -7.25), offset=(3, -11))
    if kind == "zero":
        assert not F.any()
    else:
        assert peak_err(F, ref) <= TOL64


def test_chirpz_length_boundaries_and_fallback():
    # the chirp-z execution pads each axis to the power of two >= n_in + n_out - 1 (64 .. 8192); beyond 8192 the library runs
    # the folded DMMA form instead.  Rectangular planes keep the long axis cheap.
    L = _lib.lib()
    rng = np.random.default_rng(41)
    saved = L.lfd_get_mft_variant()
    L.lfd_set_mft_variant(2)
    try:
        for (m, n), (M, N), expect in [((33, 32), (32, 33), 2),        # 64 exactly on both axes
                                       ((32, 5), (34, 7), 2),          # 65 -> 128, 11 -> 64
                                       ((2049, 8), (2048, 8), 2),      # 4096 exactly
                                       ((2100, 8), (2048, 8), 2),      # 4147 -> 8192 (one buffer, in place)
                                       ((8, 4097), (8, 4096), 2),      # 8192 exactly, on the column axis
                                       ((4200, 8), (4096, 8), 1)]:     # 8295 > 8192: folded
            d = (_lib.MftDesc * 1)()
            d[0].m, d[0].n, d[0].M, d[0].N = m, n, M, N
            assert L.lfd_mft_execution(d, 1) == expect, (m, n, M, N)
            f = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
            alpha = (0.9 / max(m, M), 0.8 / max(n, N))
            for kw in (dict(shift=(0.25, -1.5), offset=(3, -2)), dict(unitary=False)):
                got = lentil.fourier.dft2(f, alpha, shape=(M, N), **kw)
                ref = oc.dft2(f, alpha, shape=(M, N), **kw)
                assert peak_err(got, ref) <= TOL64, (m, n, M, N, kw)
            got = lentil.fourier.idft2(f, alpha, shape=(M, N), shift=(0.5, 0.0))
            assert peak_err(got, oc.idft2(f, alpha, shape=(M, N), shift=(0.5, 0.0))) <= TOL64
    finally:
        L.lfd_set_mft_variant(saved)
