"""Per-CTA timeline of the folded MFT kernel from the diagnostic build (liblentil_b200_tt.so).
Development aid."""
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
L.lfd_debug_tile_timing.argtypes = [C.c_void_p]
B, m, M = 16, 1001, 1024
dev = torch.device("cuda:0")
f = torch.randn(B, m, m, 2, dtype=torch.float64, device=dev)
o = torch.empty(B, M, M, 2, dtype=torch.float64, device=dev)
descs = (_lib.MftDesc * B)()
for b in range(B):
    d = descs[b]
    d.f = f[b].data_ptr(); d.ldf = m; d.out = o[b].data_ptr(); d.ldo = M
    d.m = m; d.n = m; d.M = M; d.N = M
    d.alpha_r = d.alpha_c = 1.0 / 2048; d.shift_r = 0.3; d.shift_c = -0.4; d.unitary = 1
need = L.lfd_mft_workspace_bytes(descs, B)
ws = torch.empty(need, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.check(L.lfd_mft_c128_batched(descs, B, ws.data_ptr(), need, st))
ntile = 8 * 32     # both stages: 256 tiles per plane
buf = torch.zeros(B * ntile * 6, dtype=torch.int64, device=dev)
L.lfd_debug_tile_timing(buf.data_ptr())
_lib.check(L.lfd_mft_c128_batched(descs, B, ws.data_ptr(), need, st))
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(-1, 6)
# the buffer holds the LAST stage written (stage 2 overwrites stage 1 at the same slots)
t = t[t[:, 3] > 0]
pro, loop, epi, tot = t[:, 1] - t[:, 0], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 3] - t[:, 0]
print("tiles", len(t), "stage FOLD_OUT flag", np.unique(t[:, 5]))
for name, v in (("prologue", pro), ("main loop", loop), ("epilogue", epi), ("total", tot)):
    print(f"{name:10s} mean {v.mean():9.0f}  p10 {np.percentile(v,10):9.0f}  p50 {np.percentile(v,50):9.0f}  p90 {np.percentile(v,90):9.0f} cycles")
# per-SM utilisation: sum of CTA lifetimes / (2 * span)
for sm in np.unique(t[:, 4])[:3]:
    s = t[t[:, 4] == sm]
    span = s[:, 3].max() - s[:, 0].min()
    print("sm", sm, "ctas", len(s), "span", span, "sum(total)/span", s[:, 3].sub if False else (s[:, 3] - s[:, 0]).sum() / span)
    order = np.argsort(s[:, 0])
    gaps = []
    print("  first starts", (s[order][:6, 0] - s[:, 0].min()).tolist())
