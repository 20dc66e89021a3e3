"""Where does a K2b CTA spend its cycles?  Diagnostic build only.  Development aid."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lentil_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "lentil_b200", "liblentil_b200_tt.so")
L = _lib.lib()
L.lfd_debug_c64_timing.argtypes = [C.c_void_p]
B, m, M = 16, 1001, 1024
dev = torch.device("cuda:0")
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
buf = torch.zeros(B * 64 * 8, dtype=torch.int64, device=dev)
L.lfd_debug_c64_timing(buf.data_ptr())
_lib.check(L.lfd_mft_c64x3_batched(descs, B, ws.data_ptr(), need, st))
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 8).astype(float)
t = t[t[:, 7] > 0]
names = ["mma wait fullA", "mma wait fullB", "mma issue+commit", "gen wait emptyA", "gen compute+store", "gen fence+arrive", "-", "mma loop total"]
nkb = 64
for i, n in enumerate(names):
    print(f"{n:20s} {t[:, i].mean() / nkb:9.1f} cycles per k-block")
