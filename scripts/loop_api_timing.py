"""The drop-in per-wavelength loop (Wavefront * Pupil -> propagate_dft -> insert) on the bench workload:
how much does the reference-style call pattern cost per wavelength?  Development aid."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lentil_b200 as lentil  # noqa: E402
import bench  # noqa: E402

w = bench.WORKLOAD
amp, opd, wls, wts = bench.make_inputs(w["nlam"])
p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=w["dx"], focal_length=w["z"])
p.freeze()


def loop(dev_acc):
    img = lentil.device.zeros_f64(1024, 1024) if dev_acc else np.zeros((1024, 1024))
    for wl, wt in zip(wls, wts):
        wf = lentil.Wavefront(wl) * p
        wf = lentil.propagate_dft(wf, pixelscale=w["du"], shape=(w["det"],) * 2, oversample=w["oversample"])
        img = wf.insert(img, wt)
    return lentil.device.to_host(img) if dev_acc else img


for dev_acc in (True, False):
    loop(dev_acc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = loop(dev_acc)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"loop API, {'device' if dev_acc else 'numpy'} accumulator: {dt*1e3:.1f} ms per 100-wavelength PSF = {100/dt:.0f} planes/s")
ref = lentil.propagate_dft_batch(p, wls, w["du"], (w["det"],) * 2, oversample=w["oversample"], weights=wts)
print("loop vs batch:", float(np.max(np.abs(out - ref)) / np.max(ref)))
