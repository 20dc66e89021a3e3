"""BASELINE.json configs 3, 4, 5 at (near) full shape: wall time per plane and a parity spot check
against the oracle.  Development aid (run under gpurun)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lentil_b200 as lentil  # noqa: E402
import lentil_oracle as oc  # noqa: E402
from lentil_b200 import synth  # noqa: E402


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


rng = np.random.default_rng(1)
# ---- config 3: 18 hex segments, ~2100^2 pupil, per-segment piston/tip/tilt, fit_tilt, 256^2 det x os2, 50 wavelengths
cube = synth.hex_segments(2, 234, 6)
n = cube.shape[1]
amp = synth.normalize_power(cube.sum(axis=0).astype(float))
opd = np.zeros((n, n))
for s in range(18):
    opd += synth.zernike_opd(cube[s], rng.uniform(-1, 1, 3) * np.array([50e-9, 2e-6, 2e-6]))
dx, z, du = 1 / 2000, 20.0, 5e-6
p = lentil.Pupil(amplitude=amp, opd=opd, mask=cube, pixelscale=dx, focal_length=z)
t0 = time.perf_counter(); p = p.fit_tilt(); t_fit = time.perf_counter() - t0
p.freeze()
wls = np.linspace(500e-9, 900e-9, 50)
dt, img = timed(lambda: lentil.propagate_dft_batch(p, wls, du, (256, 256), oversample=2, weights=np.full(50, 0.02)))
print(f"cfg3: pupil {n}^2, 18 segments x 50 wavelengths = 900 windows: {dt*1e3:.1f} ms per PSF ({900/dt:.0f} windows/s); fit_tilt {t_fit*1e3:.0f} ms")
ptilt = [(t.x, t.y) for t in p.tilt]
ref = oc.psf(p.amplitude, p.opd, cube, wls[:1], [1.0], (dx, dx), z, du, (256, 256), None, 2, plane_tilt=ptilt)
got = lentil.propagate_dft_batch(p, wls[:1], du, (256, 256), oversample=2)
print("cfg3 parity (1 wavelength):", float(np.max(np.abs(got - ref)) / np.max(ref)))
dt32, img32 = timed(lambda: lentil.propagate_dft_batch(p, wls, du, (256, 256), oversample=2, weights=np.full(50, 0.02), precision='c64'))
print(f"cfg3 c64: {dt32*1e3:.1f} ms per PSF ({900/dt32:.0f} windows/s), PSF error vs FP64 {float(np.max(np.abs(img32 - img)) / np.max(img)):.2e}")

# ---- config 4: 512^2 pupil -> 256^2 det x os2, realisations x 32 wavelengths (x3 diversity folded into realisations)
mask = synth.circle((512, 512), 250)
amp = synth.normalize_power(mask)
R = 48
opds = np.stack([synth.zernike_opd(mask, rng.normal(size=33) * 20e-9, first=4) for _ in range(R)])
p4 = lentil.Pupil(amplitude=amp, opd=np.zeros((512, 512)), pixelscale=1 / 500, focal_length=20.0)
p4.freeze()
wl4 = np.linspace(600e-9, 700e-9, 32)
opds_dev = lentil.device.to_dev(opds)
dt, st = timed(lambda: lentil.propagate_dft_batch(p4, wl4, 5e-6, (256, 256), oversample=2, weights=np.full(32, 1 / 32),
                                                   opds=opds_dev, return_device=True))
print(f"cfg4: {R} realisations x 32 wavelengths = {R*32} planes of 501^2->512^2: {dt*1e3:.1f} ms ({R*32/dt:.0f} planes/s)")
ref = oc.psf(amp, opds[3], None, wl4[:2], [0.5, 0.5], (1 / 500, 1 / 500), 20.0, 5e-6, (256, 256), None, 2)
got = lentil.propagate_dft_batch(p4, wl4[:2], 5e-6, (256, 256), oversample=2, weights=[0.5, 0.5], opds=opds[3:4])
print("cfg4 parity:", float(np.max(np.abs(got[0] - ref)) / np.max(ref)))
dt32, st32 = timed(lambda: lentil.propagate_dft_batch(p4, wl4, 5e-6, (256, 256), oversample=2, weights=np.full(32, 1 / 32),
                                                       opds=opds_dev, return_device=True, precision='c64'))
print(f"cfg4 c64: {dt32*1e3:.1f} ms ({R*32/dt32:.0f} planes/s), PSF error vs FP64 {float((st32 - st).abs().max() / st.max()):.2e}")

# ---- config 5: 4096^2 pupil -> 1024^2 det x os2, wavelengths x 16 field points
mask = synth.annulus((4096, 4096), 2040)
amp = synth.normalize_power(mask)
opd = synth.power_law_opd(mask, 30e-9, 3)
p5 = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 4080, focal_length=20.0)
p5.freeze()
tilts = [[rx, ry] for rx in np.linspace(-20e-6, 20e-6, 4) for ry in np.linspace(-20e-6, 20e-6, 4)]
wl5 = np.linspace(500e-9, 900e-9, 4)
dt, st = timed(lambda: lentil.propagate_dft_batch(p5, wl5, 5e-6, (1024, 1024), oversample=2, weights=np.full(4, 0.25),
                                                   tilts=tilts, return_device=True), reps=1)
print(f"cfg5: 4 wavelengths x 16 field points = 64 planes of 4081^2->2048^2: {dt*1e3:.0f} ms ({64/dt:.1f} planes/s, "
      f"{64*412.3/dt/1e3:.1f} TFLOP/s algorithmic)")
ref = oc.psf(amp, opd, None, wl5[:1], [1.0], (1 / 4080, 1 / 4080), 20.0, 5e-6, (1024, 1024), None, 2, wf_tilt=tilts[5])
got = lentil.propagate_dft_batch(p5, wl5[:1], 5e-6, (1024, 1024), oversample=2, tilts=[tilts[5]])
print("cfg5 parity (1 plane):", float(np.max(np.abs(got[0] - ref)) / np.max(ref)))
dt32, st32 = timed(lambda: lentil.propagate_dft_batch(p5, wl5, 5e-6, (1024, 1024), oversample=2, weights=np.full(4, 0.25),
                                                       tilts=tilts, return_device=True, precision='c64'), reps=1)
print(f"cfg5 c64: {dt32*1e3:.0f} ms ({64/dt32:.1f} planes/s), PSF error vs FP64 {float((st32 - st).abs().max() / st.max()):.2e}")
