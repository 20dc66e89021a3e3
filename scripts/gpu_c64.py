"""K2b timing + error report (complex64 / 3xTF32 on tcgen05).  Development aid."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
from lentil_b200 import _lib, device  # noqa: E402
import lentil_b200 as lentil  # noqa: E402
import lentil_oracle as oc  # noqa: E402

L = _lib.lib()
dev = device.device()
rng = np.random.default_rng(0)
for (m, M) in [(241, 256), (501, 512), (1001, 1024)]:
    f = (rng.normal(size=(m, m)) + 1j * rng.normal(size=(m, m))).astype(np.complex64)
    F = lentil.fourier.dft2_c64(f, 1 / (2 * M), shape=(M, M), shift=(0.3, -0.4))
    ref = oc.dft2(f.astype(np.complex128), 1 / (2 * M), shape=(M, M), shift=(0.3, -0.4))
    print(f"random {m}->{M}: field err {np.max(np.abs(F - ref)) / np.max(np.abs(ref)):.2e}")
    g = np.ones((m, m), np.complex64)
    F = lentil.fourier.dft2_c64(g, 1 / (2 * M), shape=(M, M))
    ref = oc.dft2(g.astype(np.complex128), 1 / (2 * M), shape=(M, M))
    I, Ir = np.abs(F.astype(np.complex128)) ** 2, np.abs(ref) ** 2
    print(f"coherent {m}->{M}: PSF err {np.max(np.abs(I - Ir)) / np.max(Ir):.2e}")

for (m, M, B) in [(1024, 1024, 64), (1001, 1024, 64), (512, 512, 128), (4096, 2048, 4)]:
    f = torch.randn(B, m, m, 2, dtype=torch.float32, device=dev)
    o = torch.empty(B, M, M, 2, dtype=torch.float32, device=dev)
    descs = (_lib.MftDesc * B)()
    for b in range(B):
        d = descs[b]
        d.f = f[b].data_ptr(); d.ldf = m; d.out = o[b].data_ptr(); d.ldo = M
        d.m = m; d.n = m; d.M = M; d.N = M
        d.alpha_r = d.alpha_c = 1.0 / 2048; d.shift_r = 0.3; d.shift_c = -0.4; d.unitary = 1
    need = L.lfd_mft_c64x3_workspace_bytes(descs, B)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        _lib.check(L.lfd_mft_c64x3_batched(descs, B, ws.data_ptr(), need, st))
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        _lib.check(L.lfd_mft_c64x3_batched(descs, B, ws.data_ptr(), need, st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flops = 8.0 * M * m * (m + M) * B
    print(json.dumps({"m": m, "M": M, "batch": B, "ms_per_plane": ms / B, "planes_per_s": B / (ms * 1e-3),
                      "tflops_algorithmic": flops / (ms * 1e-3) / 1e12}))
    del f, o, ws
