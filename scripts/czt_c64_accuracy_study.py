"""Would a complex64 chirp-z execution meet the 1e-5 gate of the complex64 mode?  CPU study (numpy / scipy.fft in single
precision): the Bluestein convolution of mft_czt.cu with every array rounded to complex64 and the FFTs run in float32, chirp
phases formed exactly in float64 and rounded once — against the float64 oracle.  Development aid (no GPU needed)."""
import os
import sys

import numpy as np
import scipy.fft as sfft

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lentil_oracle as oc  # noqa: E402
from lentil_b200 import synth  # noqa: E402


def cis(x):
    x = x - np.rint(x)
    return np.exp(2j * np.pi * x)


def czt_axis_c64(x, alpha, nout, shift, offset):
    """rows of x (axis 1) -> nout outputs, complex64 arithmetic."""
    nin = x.shape[1]
    L = 1 << int(np.ceil(np.log2(nin + nout - 1)))
    R = np.arange(nin) - nin // 2 + offset
    U = np.arange(nout) - nout // 2 - shift
    pre = cis(-0.5 * alpha * R * R).astype(np.complex64)
    post = cis(-0.5 * alpha * U * U).astype(np.complex64)
    h = np.zeros(L, dtype=np.complex64)
    p = np.arange(-(nin - 1), nout)
    D = p + ((-(nout // 2) - shift) - (-(nin // 2) + offset))
    h[p % L] = cis(0.5 * alpha * D * D).astype(np.complex64)
    H = sfft.fft(h)                                           # complex64 in -> single-precision transform
    a = np.zeros((x.shape[0], L), dtype=np.complex64)
    a[:, :nin] = x.astype(np.complex64) * pre
    y = sfft.ifft(sfft.fft(a, axis=1) * H, axis=1)[:, :nout]
    assert y.dtype == np.complex64
    return y * post


def dft2_czt_c64(f, alpha, M, N, shift=(0, 0)):
    g = czt_axis_c64(f, alpha, N, shift[1], 0.0)              # along axis 1
    F = czt_axis_c64(g.T.copy(), alpha, M, shift[0], 0.0).T   # along axis 0
    return F * np.float32(np.sqrt(alpha * alpha))


rng = np.random.default_rng(0)
for m, M in ((241, 256), (501, 512), (1001, 1024)):
    alpha = 1.0 / (2 * M)
    f = rng.normal(size=(m, m)) + 1j * rng.normal(size=(m, m))
    ref = oc.dft2(f, alpha, shape=(M, M), shift=(0.3, -0.4))
    got = dft2_czt_c64(f, alpha, M, M, shift=(0.3, -0.4))
    e_field = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    # coherent case: a filled aperture (the PSF peak is the sum of all samples)
    mask = synth.circle((m, m), m // 2 - 1).astype(float)
    ref = oc.dft2(mask, alpha, shape=(M, M))
    got = dft2_czt_c64(mask, alpha, M, M)
    I, Ir = np.abs(got.astype(np.complex128)) ** 2, np.abs(ref) ** 2
    print(f"{m}^2 -> {M}^2: random field error {e_field:.2e}; filled-aperture PSF error {np.max(np.abs(I - Ir)) / Ir.max():.2e}")
