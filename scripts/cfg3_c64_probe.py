"""cfg3 (18 segments x 50 wavelengths) in both precisions with CUDA-event timing and the per-kernel split (development aid)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lentil_b200 as lentil
from lentil_b200 import synth
rng = np.random.default_rng(1)
cube = synth.hex_segments(2, 234, 6)
n = cube.shape[1]
amp = synth.normalize_power(cube.sum(axis=0).astype(float))
opd = np.zeros((n, n))
for s in range(18):
    opd += synth.zernike_opd(cube[s], rng.uniform(-1, 1, 3) * np.array([50e-9, 2e-6, 2e-6]))
p = lentil.Pupil(amplitude=amp, opd=opd, mask=cube, pixelscale=1 / 2000, focal_length=20.0)
if '--fit' in sys.argv:
    p = p.fit_tilt()
p.freeze()
wls = np.linspace(500e-9, 900e-9, 50)
for prec in ("c128", "c64", "c64", "c128"):
    fn = lambda: lentil.propagate_dft_batch(p, wls, 5e-6, (256, 256), oversample=2, weights=np.full(50, 0.02), precision=prec, return_device=True)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(prec, "events %.2f ms/PSF, wall %.2f ms/PSF" % (e0.elapsed_time(e1) / 5, (time.perf_counter() - t0) / 5 * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    lentil.propagate_dft_batch(p, wls, 5e-6, (256, 256), oversample=2, weights=np.full(50, 0.02), precision="c64", return_device=True)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
