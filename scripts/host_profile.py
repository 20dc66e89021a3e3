"""cProfile of the end-to-end (host buffers in / numpy out) step of bench.py.  Development aid."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import lentil_b200 as lentil  # noqa: E402
import bench  # noqa: E402

w = bench.WORKLOAD
amp, opd, wls, wts = bench.make_inputs(w["nlam"])
amp_h = torch.from_numpy(amp).pin_memory().numpy()
opd_h = torch.from_numpy(opd).pin_memory().numpy()


def step():
    p = lentil.Pupil(amplitude=amp_h, opd=opd_h, pixelscale=w["dx"], focal_length=w["z"])
    return lentil.propagate_dft_batch(p, wls, w["du"], (w["det"],) * 2, oversample=w["oversample"], weights=wts)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
print("e2e ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
