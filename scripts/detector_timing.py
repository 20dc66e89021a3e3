"""Timing of the detector-side chain on the device (development aid): pixel MTF, rebin, spline rescale and pixelate of a
1024^2 oversampled PSF (BASELINE configs[1] detector grid), against the oracle (numpy/scipy-free restatement) on the host."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lentil_b200 as lentil  # noqa: E402
from lentil_b200 import device  # noqa: E402
import lentil_oracle as oc  # noqa: E402

rng = np.random.default_rng(3)
img = rng.random((1024, 1024)) ** 6
d = device.to_dev(img)


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


for name, fn in (("pixel(os=2)", lambda: lentil.detector.pixel(d, 2)), ("rebin(2)", lambda: lentil.rebin(d, 2)),
                 ("rescale(1/2, order 3)", lambda: lentil.rescale(d, 0.5)), ("pixelate(os=2)", lambda: lentil.detector.pixelate(d, 2))):
    ms, _ = timed(fn)
    print(f"{name:24s} {ms:8.3f} ms per 1024^2 image (device resident)")
t0 = time.perf_counter()
ref = oc.pixelate(img, 2)
print(f"oracle pixelate on the host: {(time.perf_counter() - t0) * 1e3:.0f} ms; device vs oracle "
      f"{np.max(np.abs(device.to_host(lentil.detector.pixelate(d, 2)) - ref)) / ref.max():.1e}")
