"""Seeded random planes through every execution of the transform: shapes from 1 x 1 to a few hundred in each direction
(all four of m, n, M, N independent, so every chirp-z length 64 .. 2048 and both the half-length and the general path
occur), random sampling, fractional shifts, integer offsets, unitary on / off, forward and inverse — against the
oracle's matrix triple product (lentil/fourier.py:5-198).  FP64 gate 1e-10, complex64 gate 1e-5."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import lentil_oracle as oc  # noqa: E402
from conftest import TOL64, peak_err  # noqa: E402

pytestmark = pytest.mark.gpu

import lentil_b200 as lentil  # noqa: E402


def _cases(seed, count, hi):
    rng = np.random.default_rng(seed)
    out = [(1, 1, 1, 1), (1, 7, 5, 1), (2, 1, 1, 3), (64, 1, 1, 64), (33, 32, 32, 33)]
    while len(out) < count:
        out.append(tuple(int(v) for v in rng.integers(1, hi, size=4)))
    res = []
    for (m, n, M, N) in out:
        alpha = (float(rng.uniform(0.2, 1.0)) / max(m, M), float(rng.uniform(0.2, 1.0)) / max(n, N))
        shift = (float(rng.uniform(-0.3, 0.3)) * M, float(rng.uniform(-0.3, 0.3)) * N)
        offset = (int(rng.integers(-m, m + 1)), int(rng.integers(-n, n + 1)))
        res.append((m, n, M, N, alpha, shift, offset, bool(rng.integers(0, 2)), int(rng.integers(0, 1 << 30))))
    return res


CASES = _cases(2024, 28, 700)


@pytest.mark.parametrize("execution", ["czt", "folded", "direct"])
def test_random_planes_fp64(execution):
    for (m, n, M, N, alpha, shift, offset, unitary, seed) in CASES:
        rng = np.random.default_rng(seed)
        f = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
        got = lentil.fourier.dft2(f, alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary, execution=execution)
        ref = oc.dft2(f, alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary)
        assert peak_err(got, ref) <= TOL64, (execution, m, n, M, N)
        got = lentil.fourier.idft2(f, alpha, shape=(M, N), shift=shift, unitary=unitary, execution=execution)
        ref = oc.idft2(f, alpha, shape=(M, N), shift=shift, unitary=unitary)
        assert peak_err(got, ref) <= TOL64, ("inverse", execution, m, n, M, N)


@pytest.mark.parametrize("execution", ["czt", "folded"])
def test_random_planes_c64(execution):
    for (m, n, M, N, alpha, shift, offset, unitary, seed) in CASES:
        rng = np.random.default_rng(seed)
        f = (rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))).astype(np.complex64)
        got = lentil.fourier.dft2_c64(f, alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary, execution=execution)
        ref = oc.dft2(f.astype(np.complex128), alpha, shape=(M, N), shift=shift, offset=offset, unitary=unitary)
        assert got.dtype == np.complex64
        assert peak_err(got, ref) <= 1e-5, (execution, m, n, M, N)


def test_mixed_batch_of_all_lengths_in_one_launch():
    """one batched launch holding planes of every chirp-z length (64 .. 2048) and shape: the per-length unit tables"""
    import torch
    from lentil_b200 import _lib, device, fourier
    rng = np.random.default_rng(77)
    shapes = [(5, 9, 30, 20), (40, 70, 60, 50), (100, 130, 120, 90), (250, 200, 260, 300), (500, 40, 512, 30), (1001, 12, 1024, 10),
              (12, 1001, 10, 1024), (33, 32, 32, 33)]
    fs, outs, refs = [], [], []
    descs = (_lib.MftDesc * len(shapes))()
    for k, (m, n, M, N) in enumerate(shapes):
        f = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
        alpha = (0.7 / max(m, M), 0.6 / max(n, N))
        fd = device.to_dev(f, dtype=np.complex128)
        od = torch.empty(M, N, dtype=torch.complex128, device=fd.device)
        fourier.mft_descriptor(descs[k], fd, od, alpha, (0.4, -1.3), (1, -2), True, False, "czt")
        fs.append(fd); outs.append(od)
        refs.append(oc.dft2(f, alpha, shape=(M, N), shift=(0.4, -1.3), offset=(1, -2)))
    fourier.run_mft(descs, len(shapes))
    for od, ref in zip(outs, refs):
        assert peak_err(device.to_host(od), ref) <= TOL64
