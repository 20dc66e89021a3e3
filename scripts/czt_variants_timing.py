"""Chirp-z execution: per-plane time of the shapes that matter (bench plane, cfg3/cfg4/cfg5 planes) for one build of the
library, plus the bench pipeline (K1 fused -> K2a -> K3) in planes/s.  With --all it runs itself once per tagged build found
next to the production library (LFD_LIB=...), one subprocess each, and prints one JSON line per build.  Development aid."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one():
    import numpy as np
    import torch
    import lentil_b200 as lentil
    from lentil_b200 import _lib, device
    import bench
    L = _lib.lib()
    dev = device.device()
    res = {"lib": os.path.basename(_lib.LIB_PATH)}
    shapes = [(241, 256, 256), (501, 512, 256), (1001, 1024, 64), (2001, 2048, 16), (4081, 2048, 4)]
    for (m, M, B) in shapes:
        f = torch.randn(B, m, m, dtype=torch.complex128, device=dev)
        o = torch.empty(B, M, M, dtype=torch.complex128, device=dev)
        descs = (_lib.MftDesc * B)()
        for b in range(B):
            d = descs[b]
            d.f = f[b].data_ptr(); d.ldf = m; d.out = o[b].data_ptr(); d.ldo = M; d.m = d.n = m; d.M = d.N = M
            d.alpha_r = d.alpha_c = 1.0 / (2 * M); d.shift_r = 0.3; d.shift_c = -0.4; d.unitary = 1
        out = {}
        ref = None
        for name, v in (("folded", 1), ("czt", 2)):
            if name == "folded" and (m > 2001 or "--no-folded" in sys.argv):
                continue
            L.lfd_set_mft_variant(v)
            need = L.lfd_mft_workspace_bytes(descs, B)
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(2):
                _lib.check(L.lfd_mft_c128_batched(descs, B, ws.data_ptr(), need, st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                _lib.check(L.lfd_mft_c128_batched(descs, B, ws.data_ptr(), need, st))
            e1.record()
            torch.cuda.synchronize()
            out[name + "_us"] = round(e0.elapsed_time(e1) / 3 / B * 1e3, 2)
            if name == "folded":
                ref = o.clone()
        if ref is not None:
            out["max_diff"] = float((o - ref).abs().max() / ref.abs().max())
        res[f"{m}->{M}"] = out
        del f, o
    L.lfd_set_mft_variant(3)
    # complex64 mode: FP32 chirp-z against the tcgen05 3xTF32 form (time per plane, error of each against the FP64 result)
    for (m, M, B) in [(501, 512, 256), (1001, 1024, 64), (2001, 2048, 16)]:
        f = torch.randn(B, m, m, dtype=torch.complex64, device=dev)
        o = torch.empty(B, M, M, dtype=torch.complex64, device=dev)
        f64 = f[:2].to(torch.complex128).contiguous()
        o64 = torch.empty(2, M, M, dtype=torch.complex128, device=dev)
        descs = (_lib.MftDesc * B)()
        d64 = (_lib.MftDesc * 2)()
        for b in range(B):
            for (dd, ff, oo, nb) in ((descs, f, o, B), (d64, f64, o64, 2)):
                if b >= nb:
                    continue
                d = dd[b]
                d.f = ff[b].data_ptr(); d.ldf = m; d.out = oo[b].data_ptr(); d.ldo = M; d.m = d.n = m; d.M = d.N = M
                d.alpha_r = d.alpha_c = 1.0 / (2 * M); d.shift_r = 0.3; d.shift_c = -0.4; d.unitary = 1
        need = L.lfd_mft_workspace_bytes(d64, 2)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(L.lfd_mft_c128_batched(d64, 2, ws.data_ptr(), need, st))
        out = {}
        for name, code in (("tcgen05", 2), ("czt", 3)):
            for b in range(B):
                descs[b].execution = code
            need = L.lfd_mft_c64x3_workspace_bytes(descs, B)
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            for _ in range(2):
                _lib.check(L.lfd_mft_c64x3_batched(descs, B, ws.data_ptr(), need, st))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                _lib.check(L.lfd_mft_c64x3_batched(descs, B, ws.data_ptr(), need, st))
            e1.record()
            torch.cuda.synchronize()
            out[name + "_us"] = round(e0.elapsed_time(e1) / 3 / B * 1e3, 2)
            out[name + "_err"] = float((o[:2].to(torch.complex128) - o64).abs().max() / o64.abs().max())
        res[f"c64 {m}->{M}"] = out
        del f, o
    # bench pipeline
    w = bench.WORKLOAD
    amp, opd, wls, wts = bench.make_inputs(w["nlam"])
    pupil = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=w["dx"], focal_length=w["z"])
    pupil.freeze()

    def step():
        return lentil.propagate_dft_batch(pupil, wls, w["du"], (w["det"],) * 2, oversample=w["oversample"], weights=wts,
                                          return_device=True)
    for _ in range(3):
        psf = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        psf = step()
    e1.record()
    torch.cuda.synchronize()
    res["bench_planes_per_s"] = round(w["nlam"] * 10 / (e0.elapsed_time(e1) * 1e-3), 1)
    for name, ex in (("bench_c64_czt", "czt"), ("bench_c64_tcgen05", "folded")):
        def step32():
            return lentil.propagate_dft_batch(pupil, wls, w["du"], (w["det"],) * 2, oversample=w["oversample"], weights=wts,
                                              return_device=True, precision="c64", execution=ex)
        for _ in range(3):
            p32 = step32()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            p32 = step32()
        e1.record()
        torch.cuda.synchronize()
        res[name] = round(w["nlam"] * 10 / (e0.elapsed_time(e1) * 1e-3), 1)
        res[name + "_err"] = float((p32 - psf).abs().max() / psf.max())
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lentil_oracle as oc
    ref1 = oc.psf(amp, opd, None, wls[:1], wts[:1], (w["dx"], w["dx"]), w["z"], w["du"], (w["det"],) * 2, None, w["oversample"])
    got1 = lentil.propagate_dft_batch(pupil, wls[:1], w["du"], (w["det"],) * 2, oversample=w["oversample"], weights=wts[:1])
    res["bench_parity"] = float(np.max(np.abs(got1 - ref1)) / np.max(ref1))
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    if "--all" in sys.argv:
        libs = sorted(glob.glob(os.path.join(ROOT, "lentil_b200", "liblentil_b200*.so")))
        for k, lib in enumerate(libs):
            env = dict(os.environ, LFD_LIB=lib)
            args = [sys.executable, os.path.abspath(__file__)] + (["--no-folded"] if k else [])
            subprocess.run(args, env=env, check=False)
    else:
        one()
