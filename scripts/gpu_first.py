"""First GPU contact: FP64 probe, dft2 correctness against the closed-form numpy expression, and
a batched 1024^2 -> 1024^2 timing.  Development aid (run under gpurun), not part of the product."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lentil_b200 import _lib  # noqa: E402


def np_dft2(f, alpha, shape, shift=(0, 0), offset=(0, 0), unitary=True, inverse=False):
    ar, ac = np.broadcast_to(alpha, (2,))
    m, n = f.shape
    M, N = np.broadcast_to(shape, (2,))
    R = np.arange(m) - np.floor(m / 2.0) + offset[0]
    S = np.arange(n) - np.floor(n / 2.0) + offset[1]
    U = np.arange(M) - np.floor(M / 2.0) - shift[0]
    V = np.arange(N) - np.floor(N / 2.0) - shift[1]
    sg = 1.0 if inverse else -1.0
    E1 = np.exp(sg * 2j * np.pi * ar * np.outer(U, R))
    E2 = np.exp(sg * 2j * np.pi * ac * np.outer(S, V))
    F = E1 @ f @ E2
    if unitary:
        F = F * np.sqrt(abs(ar * ac))
    if inverse:
        F = F / f.size
    return F


def main():
    L = _lib.lib()
    out = (C.c_double * 3)()
    _lib.check(L.lfd_probe_fp64(out, 20000), "probe")
    print(json.dumps({"probe_dmma_tflops": out[0], "probe_dfma_tflops": out[1], "clock_mhz": out[2]}))

    variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    L.lfd_set_mft_variant(variant)
    print("variant", L.lfd_get_mft_variant())
    ctx = L.lfd_ctx_create(0)
    assert ctx, L.lfd_last_error()
    rng = np.random.default_rng(0)
    cases = [
        (10, 10, 10, 10, (0.1, 0.1), (0, 0), (0, 0), True, False),
        (11, 13, 17, 9, (1 / 11, 1 / 13), (0.3, -1.7), (2, -3), True, False),
        (241, 241, 256, 256, (0.0013, 0.0013), (0.4, 0.6), (0, 0), True, False),
        (300, 200, 130, 260, (0.002, 0.0031), (13.4, -7.6), (-40, 25), False, False),
        (64, 64, 64, 64, (1 / 64, 1 / 64), (0, 0), (0, 0), False, True),
        (501, 501, 486, 499, (3.846e-4, 3.846e-4), (13.4, 7.6), (0, 0), True, False),
        (1001, 1001, 1024, 1024, (1 / 2048, 1 / 2048), (0.3, -0.4), (0, 0), True, False),
    ]
    worst = 0.0
    for (m, n, M, N, alpha, shift, off, unitary, inverse) in cases:
        f = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
        F = np.empty((M, N), dtype=np.complex128)
        rc = L.lfd_ctx_dft2_host(ctx, f.ctypes.data, n, m, n, alpha[0], alpha[1], M, N,
                                 shift[0], shift[1], off[0], off[1], int(unitary), int(inverse),
                                 F.ctypes.data, N)
        _lib.check(rc, "dft2_host")
        ref = np_dft2(f, alpha, (M, N), shift, off, unitary, inverse)
        err = np.max(np.abs(F - ref)) / np.max(np.abs(ref))
        worst = max(worst, err)
        print(f"dft2 {m}x{n}->{M}x{N} inv={inverse} rel_err={err:.3e}")
    print("worst", worst)

    import torch
    dev = torch.device("cuda:0")
    for (m, M, B) in [(1024, 1024, 32), (1024, 512, 32), (512, 512, 64), (4096, 2048, 2)]:
        f = torch.randn(B, m, m, 2, dtype=torch.float64, device=dev)
        o = torch.empty(B, M, M, 2, dtype=torch.float64, device=dev)
        descs = (_lib.MftDesc * B)()
        for b in range(B):
            d = descs[b]
            d.f = f[b].data_ptr(); d.ldf = m; d.out = o[b].data_ptr(); d.ldo = M
            d.m = m; d.n = m; d.M = M; d.N = M
            d.alpha_r = d.alpha_c = 1.0 / 2048
            d.shift_r = 0.3; d.shift_c = -0.4
            d.unitary = 1
        need = L.lfd_mft_workspace_bytes(descs, B)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            _lib.check(L.lfd_mft_c128_batched(descs, B, ws.data_ptr(), need, st))
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            _lib.check(L.lfd_mft_c128_batched(descs, B, ws.data_ptr(), need, st))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 8.0 * M * m * (m + M) * B
        print(json.dumps({"m": m, "M": M, "batch": B, "ms_per_plane": ms / B,
                          "planes_per_s": B / (ms * 1e-3), "tflops": flops / (ms * 1e-3) / 1e12}))
        del f, o, ws
    L.lfd_ctx_destroy(ctx)


if __name__ == "__main__":
    main()
