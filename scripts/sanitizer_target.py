"""Small target for compute-sanitizer (memcheck / racecheck / synccheck): every chirp-z length once (FP64 and FP32 builds,
half-length and general path, staged and unstaged units, mixed batch), the fused pupil path with several wavelengths, K3.
Development aid:  compute-sanitizer --tool racecheck python scripts/sanitizer_target.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import lentil_b200 as lentil  # noqa: E402
import lentil_oracle as oc  # noqa: E402
from lentil_b200 import synth  # noqa: E402

rng = np.random.default_rng(0)
worst = 0.0
for (m, n, M, N) in [(5, 9, 30, 20), (40, 70, 60, 50), (100, 130, 120, 90), (250, 200, 260, 300), (300, 40, 200, 30), (1001, 12, 1024, 10),
                     (12, 1001, 10, 1024), (2049, 3, 2048, 3), (3, 2100, 3, 2048)]:
    f = rng.normal(size=(m, n)) + 1j * rng.normal(size=(m, n))
    alpha = (0.7 / max(m, M), 0.6 / max(n, N))
    ref = oc.dft2(f, alpha, shape=(M, N), shift=(0.4, -1.3), offset=(1, -2))
    got = lentil.fourier.dft2(f, alpha, shape=(M, N), shift=(0.4, -1.3), offset=(1, -2), execution="czt")
    worst = max(worst, float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))))
    got = lentil.fourier.dft2_c64(f.astype(np.complex64), alpha, shape=(M, N), shift=(0.4, -1.3), offset=(1, -2), execution="czt")
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-5
mask = synth.annulus((1030, 1030), 500, 0.3)          # bbox 1001: the 2048-point staged kernels, odd column offset
amp = synth.normalize_power(mask)
opd = synth.zernike_opd(mask, rng.normal(size=8) * 30e-9)
p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 1000, focal_length=20.0)
wls = np.linspace(500e-9, 900e-9, 3)
for prec in ("c128", "c64"):
    img = lentil.propagate_dft_batch(p, wls, 5e-6, (512, 512), oversample=2, weights=[0.2, 0.5, 0.3], precision=prec, execution="czt")
    ref = oc.psf(amp, opd, None, wls, [0.2, 0.5, 0.3], (1 / 1000,) * 2, 20.0, 5e-6, (512, 512), None, 2)
    e = float(np.max(np.abs(img - ref)) / np.max(ref))
    assert e < (1e-10 if prec == "c128" else 1e-5), (prec, e)
    worst = max(worst, e if prec == "c128" else 0.0)
print("sanitizer target done, worst FP64 error", worst)
