"""CPU emulation of the index algebra of the chirp-z kernel (lentil_b200/csrc/mft_czt.cu), so that host-only CI
exercises the math the GPU tests cover on hardware: the radix-16 Stockham pass plan (scatter / own-element
addresses, per-pass twiddle tables, the 4 x 4 split of the 16-point butterfly with its constants), the "turn"
(last forward pass -> x H -> first adjoint pass in registers), the adjoint passes in reverse order, the padded
shared-memory slots, and the pre / post / lag-wrapped chirps of the Bluestein convolution.  Every line mirrors
a line of the kernel; numpy's FFT and the oracle's dft2 (lentil/fourier.py:95-101) are the checkers."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import lentil_oracle as oc  # noqa: E402

H = 0.70710678118654752440
C1, S1 = 0.92387953251128675613, 0.38268343236508977173


def nreg(lg):
    return (lg - 1) // 4


def turn_radix(lg):
    return 1 << (lg - 4 * nreg(lg))


def tw_offset(p):
    return ((1 << (4 * p)) - 16) // 15


def roots(lg):
    """g_tw[lg] as roots_kernel fills it."""
    tab = np.zeros(4400, complex)
    for p in range(1, nreg(lg) + 1):
        Ns = 1 << (4 * p)
        den = 16.0 * Ns if p < nreg(lg) else float(1 << lg)
        k = np.arange(Ns)
        tab[tw_offset(p):tw_offset(p) + Ns] = np.exp(-2j * np.pi * k / den)
    return tab


def mul_mi(a, S):
    return a * (-1j if S > 0 else 1j)


def mul_root(a, wr, wi, S):
    return a * (wr - 1j * wi if S > 0 else wr + 1j * wi)


def dft4(x0, x1, x2, x3, S):
    s0, s1, s2, s3 = x0 + x2, x0 - x2, x1 + x3, mul_mi(x1 - x3, S)
    return s0 + s2, s1 + s3, s0 - s2, s1 - s3


def dft16(v, S):
    """v: (16, ...) -> (16, ...), the register-level 4 x 4 form of the kernel."""
    v = [np.array(x) for x in v]
    for n2 in range(4):
        v[n2], v[4 + n2], v[8 + n2], v[12 + n2] = dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2], S)
    v[5] = mul_root(v[5], C1, S1, S)
    v[6] = mul_root(v[6], H, H, S)
    v[7] = mul_root(v[7], S1, C1, S)
    v[9] = mul_root(v[9], H, H, S)
    v[10] = mul_mi(v[10], S)
    v[11] = mul_root(v[11], -H, H, S)
    v[13] = mul_root(v[13], S1, C1, S)
    v[14] = mul_root(v[14], -H, H, S)
    v[15] = mul_root(v[15], -C1, -S1, S)
    for k1 in range(4):
        v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3] = dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3], S)
    for a in range(4):
        for b in range(a + 1, 4):
            v[4 * a + b], v[4 * b + a] = v[4 * b + a], v[4 * a + b]
    return np.array(v)


def dftR(v, S):
    R = len(v)
    if R == 16:
        return dft16(v, S)
    k = np.arange(R)
    F = np.exp(-S * 2j * np.pi * np.outer(k, k) / R)       # the small butterflies are plain DFTs (dft2p/dft4/dft8 unchanged from round 1)
    v = np.array(v)
    return np.tensordot(F.astype(v.dtype), v, axes=(1, 0))


def twiddle_powers(v, w1, conj):
    """four interleaved chains stepping by w^4, as in the kernel"""
    R = len(v)
    v = [np.array(x) for x in v]
    if conj:
        w1 = np.conj(w1)
    v[1] = v[1] * w1
    if R >= 4:
        w2 = w1 * w1
        w3 = w2 * w1
        v[2] = v[2] * w2
        v[3] = v[3] * w3
        if R >= 8:
            w4 = w2 * w2
            q0, q1, q2, q3 = w4, w1, w2, w3
            for a in range(4, R, 4):
                q1, q2, q3 = q1 * w4, q2 * w4, q3 * w4
                v[a], v[a + 1], v[a + 2], v[a + 3] = v[a] * q0, v[a + 1] * q1, v[a + 2] * q2, v[a + 3] * q3
                if a + 4 < R:
                    q0 = q0 * w4
    return np.array(v)


def slot(i):
    return i + (i >> 4)


def czt_row_emulated(x, Hf, lg):
    """y = IFFT_unnormalised(FFT(x) * Hf) exactly as czt_row schedules it (one buffer, in place)."""
    L = 1 << lg
    T, NREG, RT = L // 16, nreg(lg), turn_radix(lg)
    NB, NS = 16 // RT, L // RT
    tw = roots(lg).astype(x.dtype)                      # the complex64 build reads the same roots rounded once (g_twf)
    X = np.full(L + L // 16 + 1, np.nan, x.dtype)
    t = np.arange(T)
    s = np.arange(16)[:, None]
    v = dft16(x[t[None, :] + s * T], 1)
    X[(17 * t)[None, :] + s] = v
    Ns = 16
    for p in range(1, NREG):
        k = t & (Ns - 1)
        j0 = (t - k) * 16 + k
        w1 = tw[tw_offset(p) + k]
        v = X[slot(t[None, :] + s * T)]
        v = dft16(twiddle_powers(v, w1, False), 1)
        X[:] = np.nan                                   # in place: every element is read before any is rewritten
        X[slot(j0[None, :] + s * Ns)] = v
        Ns *= 16
    for q in range(NB):
        j = t + q * T
        r = np.arange(RT)[:, None]
        idx = slot(t[None, :] + (q + r * NB) * T)
        u = X[idx]
        w1 = tw[tw_offset(NREG) + j]
        u = dftR(twiddle_powers(u, w1, False), 1)
        u = u * Hf[j[None, :] + r * NS]
        u = twiddle_powers(dftR(u, -1), w1, True)
        X[idx] = u
    for p in range(NREG - 1, 0, -1):
        Ns //= 16
        k = t & (Ns - 1)
        j0 = (t - k) * 16 + k
        w1 = tw[tw_offset(p) + k]
        v = X[slot(j0[None, :] + s * Ns)]
        v = twiddle_powers(dft16(v, -1), w1, True)
        X[:] = np.nan
        X[slot(t[None, :] + s * T)] = v
    v = dft16(X[(17 * t)[None, :] + s], -1)
    y = np.empty(L, x.dtype)
    y[t[None, :] + s * T] = v
    return y


def forward_fft_emulated(x, lg):
    """the plain forward transform of czt_tables_kernel (chirp filter H)"""
    L = 1 << lg
    T, NREG, RT = L // 16, nreg(lg), turn_radix(lg)
    NB, NS = 16 // RT, L // RT
    tw = roots(lg)
    X = np.full(L + L // 16 + 1, np.nan, complex)
    t = np.arange(T)
    s = np.arange(16)[:, None]
    X[(17 * t)[None, :] + s] = dft16(x[t[None, :] + s * T], 1)
    Ns = 16
    for p in range(1, NREG):
        k = t & (Ns - 1)
        j0 = (t - k) * 16 + k
        v = dft16(twiddle_powers(X[slot(t[None, :] + s * T)], tw[tw_offset(p) + k], False), 1)
        X[:] = np.nan
        X[slot(j0[None, :] + s * Ns)] = v
        Ns *= 16
    out = np.empty(L, complex)
    for q in range(NB):
        j = t + q * T
        r = np.arange(RT)[:, None]
        u = dftR(twiddle_powers(X[slot(t[None, :] + (q + r * NB) * T)], tw[tw_offset(NREG) + j], False), 1)
        out[j[None, :] + r * NS] = u
    return out


def test_dft16_register_form():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(16, 5)) + 1j * rng.normal(size=(16, 5))
    assert np.allclose(dft16(x, 1), np.fft.fft(x, axis=0), atol=1e-13)
    assert np.allclose(dft16(x, -1), np.fft.ifft(x, axis=0) * 16, atol=1e-13)


@pytest.mark.parametrize("R", [2, 4, 8, 16])
def test_twiddle_power_chains(R):
    rng = np.random.default_rng(R)
    v = rng.normal(size=(R, 7)) + 1j * rng.normal(size=(R, 7))
    w = np.exp(-2j * np.pi * rng.uniform(size=7))
    r = np.arange(R)[:, None]
    assert np.allclose(twiddle_powers(v, w, False), v * w[None, :] ** r, atol=1e-13)
    assert np.allclose(twiddle_powers(v, w, True), v * np.conj(w[None, :]) ** r, atol=1e-13)


@pytest.mark.parametrize("lg", range(6, 14))
def test_forward_plan_is_an_fft(lg):
    rng = np.random.default_rng(lg)
    x = rng.normal(size=1 << lg) + 1j * rng.normal(size=1 << lg)
    got = forward_fft_emulated(x, lg)
    assert np.max(np.abs(got - np.fft.fft(x))) <= 1e-11 * np.max(np.abs(got))


@pytest.mark.parametrize("lg", range(6, 14))
def test_row_is_a_circular_convolution(lg):
    L = 1 << lg
    rng = np.random.default_rng(100 + lg)
    x = rng.normal(size=L) + 1j * rng.normal(size=L)
    h = rng.normal(size=L) + 1j * rng.normal(size=L)
    got = czt_row_emulated(x, np.fft.fft(h), lg) / L
    want = np.fft.ifft(np.fft.fft(x) * np.fft.fft(h))
    assert np.max(np.abs(got - want)) <= 1e-11 * np.max(np.abs(want))


def chirpz_axis(x, alpha, x0, y0, nout, sgn, lg):
    """One axis of the transform through the emulated row: y[u] = sum_i x[i] exp(sgn 2 pi i alpha (i + x0)(u + y0))."""
    nin, L = len(x), 1 << lg
    i = np.arange(nin) + x0
    u = np.arange(nout) + y0
    pre = np.exp(sgn * 1j * np.pi * alpha * i * i)
    post = np.exp(sgn * 1j * np.pi * alpha * u * u)
    q = np.arange(L)
    p = np.where(q < nout, q, q - L)
    D = p + (y0 - x0)
    h = np.where((p > -nin) & (p < nout), np.exp(-sgn * 1j * np.pi * alpha * D * D), 0)
    xp = np.zeros(L, complex)
    xp[:nin] = x * pre
    return czt_row_emulated(xp, forward_fft_emulated(h, lg), lg)[:nout] * post / L


@pytest.mark.parametrize("m,n,M,N,alpha,shift,offset", [
    (30, 27, 35, 38, (0.013, 0.017), (0.3, -1.2), (2, -3)),
    (100, 120, 129, 90, (0.004, 0.0033), (-4.75, 2.5), (0, 0)),
    (501, 40, 512, 25, (0.0011, 0.02), (0.0, 0.0), (-7, 4)),
])
def test_chirpz_equals_oracle_dft2(m, n, M, N, alpha, shift, offset):
    """both stages (rows, then the columns of the transposed intermediate) against the oracle's matrix triple product"""
    rng = np.random.default_rng(m + n)
    f = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
    want = oc.dft2(f, alpha, shape=(M, N), shift=shift, offset=offset, unitary=True)

    def lg_for(a, b):
        lg = 6
        while (1 << lg) < a + b - 1:
            lg += 1
        return lg
    x0r, y0r = -np.floor(m / 2.0) + offset[0], -np.floor(M / 2.0) - shift[0]
    x0c, y0c = -np.floor(n / 2.0) + offset[1], -np.floor(N / 2.0) - shift[1]
    G = np.array([chirpz_axis(f[i], alpha[1], x0c, y0c, N, -1.0, lg_for(n, N)) for i in range(m)])           # stage A
    F = np.array([chirpz_axis(G[:, v], alpha[0], x0r, y0r, M, -1.0, lg_for(m, M)) for v in range(N)]).T      # stage B
    F *= np.sqrt(abs(alpha[0] * alpha[1]))
    assert np.max(np.abs(F - want)) <= 1e-10 * np.max(np.abs(want))


@pytest.mark.parametrize("lg", [9, 11, 13])
def test_complex64_build_accuracy(lg):
    """The FP32 build of the row transform (complex64 mode): complex64 data and twiddles, the chirp filter H built in float64
    and rounded once.  Peak-normalised error of the convolution against float64 stays ~1e-6, well inside the 1e-5 gate."""
    L = 1 << lg
    rng = np.random.default_rng(lg)
    x = np.zeros(L, complex)
    x[:L // 2 - 3] = rng.normal(size=L // 2 - 3) + 1j * rng.normal(size=L // 2 - 3)
    h = np.exp(1j * np.pi * 0.37e-3 * (np.arange(L) - L / 2) ** 2)            # a chirp, like the real filter
    Hf = np.fft.fft(h)
    want = np.fft.ifft(np.fft.fft(x) * Hf)
    got = czt_row_emulated(x.astype(np.complex64), Hf.astype(np.complex64), lg)
    assert got.dtype == np.complex64
    err = np.max(np.abs(got / L - want)) / np.max(np.abs(want))
    assert err <= 3e-6, err


def dft16_pruned(v, S, in8=False, out8=False):
    """the IN8 / OUT8 forms of the kernel's dft16: zero upper half of the inputs not read, only outputs 0..7 formed"""
    v = [np.array(x) for x in v]
    for n2 in range(4):
        if in8:
            x0, x1 = v[n2], v[4 + n2]
            m = mul_mi(x1, S)
            v[n2], v[4 + n2], v[8 + n2], v[12 + n2] = x0 + x1, x0 + m, x0 - x1, x0 - m
        else:
            v[n2], v[4 + n2], v[8 + n2], v[12 + n2] = dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2], S)
    v[5] = mul_root(v[5], C1, S1, S)
    v[6] = mul_root(v[6], H, H, S)
    v[7] = mul_root(v[7], S1, C1, S)
    v[9] = mul_root(v[9], H, H, S)
    v[10] = mul_mi(v[10], S)
    v[11] = mul_root(v[11], -H, H, S)
    v[13] = mul_root(v[13], S1, C1, S)
    v[14] = mul_root(v[14], -H, H, S)
    v[15] = mul_root(v[15], -C1, -S1, S)
    if not out8:
        return dft16_tail(v, S)
    lo = [(v[4 * k1] + v[4 * k1 + 2]) + (v[4 * k1 + 1] + v[4 * k1 + 3]) for k1 in range(4)]
    hi = [(v[4 * k1] - v[4 * k1 + 2]) + mul_mi(v[4 * k1 + 1] - v[4 * k1 + 3], S) for k1 in range(4)]
    return np.array(lo + hi)


def dft16_tail(v, S):
    for k1 in range(4):
        v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3] = dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3], S)
    for a in range(4):
        for b in range(a + 1, 4):
            v[4 * a + b], v[4 * b + a] = v[4 * b + a], v[4 * a + b]
    return np.array(v)


@pytest.mark.parametrize("S", [1, -1])
def test_pruned_butterflies_of_the_half_length_path(S):
    rng = np.random.default_rng(7)
    x = rng.normal(size=(16, 6)) + 1j * rng.normal(size=(16, 6))
    full = dft16(x, S)
    assert np.allclose(dft16_pruned(x, S, out8=True), full[:8], atol=1e-13)
    xz = x.copy()
    xz[8:] = 0
    garbage = x.copy()
    garbage[8:] = np.nan                                # IN8 must not read the upper half
    assert np.allclose(dft16_pruned(garbage, S, in8=True), dft16(xz, S), atol=1e-13)
