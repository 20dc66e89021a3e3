"""Where does a K2b CTA spend its cycles?  Diagnostic build only.  Development aid."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lentil_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, "lentil_b200", os.environ.get("LFD_TT_LIB", "liblentil_b200_tt.so"))
L = _lib.lib()
L.lfd_debug_c64_timing.argtypes = [C.c_void_p, C.c_int]
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
stage = os.environ.get("LFD_TT_STAGE", "both")
buf = torch.zeros(B * 128 * 16, dtype=torch.int64, device=dev)
trace = torch.zeros(16 * 128, dtype=torch.int64, device=dev)
L.lfd_debug_c64_trace.argtypes = [C.c_void_p]
L.lfd_debug_c64_trace(trace.data_ptr())
L.lfd_debug_c64_timing(buf.data_ptr(), {"both": -1, "row": 1, "col": 0}[stage])
print("stage:", stage)
_lib.check(L.lfd_mft_c64x3_batched(descs, B, ws.data_ptr(), need, st))
torch.cuda.synchronize()
tr = trace.cpu().numpy().reshape(16, 128)
t = buf.cpu().numpy().reshape(-1, 16).astype(float)
t = t[t[:, 7] > 0]
nkb = 64
print(f"{len(t)} CTAs recorded (the column stage overwrites the row stage where both use a slot)")
names = ["mma wait fullA", "mma wait fullB", "mma issue+commit", "gen wait emptyA", "gen compute", "gen st+fence+arrive", "-",
         "mma loop total"]
for i, n in enumerate(names):
    if n != "-":
        print(f"{n:22s} {(t[:, i] / t[:, 13]).mean():9.1f} cycles per k-block   (generator columns: per set, i.e. per 2 k-blocks)"
              if n.startswith("gen") else f"{n:22s} {(t[:, i] / t[:, 13]).mean():9.1f} cycles per k-block")
for i, n in [(8, "prologue (entry -> generator loop)"), (9, "generator loop"), (10, "wait for last MMAs"), (11, "epilogue proper"),
             (12, "CTA total")]:
    print(f"{n:36s} {t[:, i].mean():10.0f} cycles  (min {t[:, i].min():.0f}, max {t[:, i].max():.0f})")

if os.environ.get("LFD_TT_TRACE"):
    t0 = tr[0, 0]
    print("jb : issue start | issue end | waits done || arrival of generator warps on fullA(jb) (cycles since first issue)")
    for j in range(14, 44):
        arr = [int(tr[6 + w, j] - t0) for w in (range(4) if j % 2 == 0 else range(4, 8))]
        extra = "  w3: loop-top %d barrier-passed %d computed %d st-done %d arrive %d | w4: computed %d st-done %d arrive %d" % tuple(int(tr[e, j] - t0) for e in (5, 2, 3, 14, 8, 4, 15, 9)) if j % 2 == 0 else ""
        if j % 2 == 0:
            extra += "  w1: loop-top %d computed %d wait+fence done %d st-done %d arrive %d" % tuple(int(tr[e, j] - t0) for e in (10, 11, 12, 13, 6))
        print(f"{j:3d}: {int(tr[0, j] - t0):8d} | {int(tr[1, j] - t0):8d} | {int(tr[2, j] - t0):8d} || " + " ".join(f"{a:7d}" for a in arr) + extra)
