"""Batched timing of small planes under the folded and the chirp-z execution (where should LFD_MFT_AUTO switch?).  Development aid."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lentil_b200 import _lib, device
L = _lib.lib(); dev = device.device()
for (m, M, B) in [(121, 128, 512), (241, 256, 512), (401, 512, 256), (501, 512, 256), (1001, 1024, 64), (2001, 2048, 16)]:
    f = torch.randn(B, m, m, dtype=torch.complex128, device=dev); o = torch.empty(B, M, M, dtype=torch.complex128, device=dev)
    descs = (_lib.MftDesc * B)()
    for b in range(B):
        d = descs[b]; d.f = f[b].data_ptr(); d.ldf = m; d.out = o[b].data_ptr(); d.ldo = M; d.m = d.n = m; d.M = d.N = M
        d.alpha_r = d.alpha_c = 1.0 / (2 * M); d.shift_r = 0.3; d.shift_c = -0.4; d.unitary = 1
    res = {}
    for name, v in (("folded", 1), ("czt", 2)):
        L.lfd_set_mft_variant(v)
        need = L.lfd_mft_workspace_bytes(descs, B); ws = torch.empty(need, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2): _lib.check(L.lfd_mft_c128_batched(descs, B, ws.data_ptr(), need, st))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): _lib.check(L.lfd_mft_c128_batched(descs, B, ws.data_ptr(), need, st))
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 3 / B * 1e3
        ref = o.clone() if name == "folded" else ref
    err = float((o - ref).abs().max() / ref.abs().max())
    print(f"{m}^2 -> {M}^2 x{B}: folded {res['folded']:.2f} us/plane, czt {res['czt']:.2f} us/plane, ratio {res['folded']/res['czt']:.2f}, max diff {err:.1e}")
L.lfd_set_mft_variant(3)
