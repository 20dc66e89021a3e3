"""ncu target for K2b: a few batched complex64 launches (16 planes 1001^2 -> 1024^2).  Development aid."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lentil_b200 import _lib, device  # noqa: E402

L = _lib.lib()
dev = device.device()
B, m, M = int(os.environ.get("LFD_B", "16")), 1001, 1024
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
for _ in range(3):
    _lib.check(L.lfd_mft_c64x3_batched(descs, B, ws.data_ptr(), need, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok")
