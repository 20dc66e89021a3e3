"""Small target for `ncu --set full`: a few batched K2a launches (8 planes of the cfg2 shape,
1001^2 -> 1024^2) plus one K1 and one K3 launch.  Development aid."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lentil_b200 as lentil  # noqa: E402
from lentil_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mask = synth.annulus((1024, 1024), 500)
amp = synth.normalize_power(mask)
opd = synth.zernike_opd(mask, np.random.default_rng(0).normal(size=15) * 30e-9)
p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 1000, focal_length=20.0)
p.freeze()
wls = np.linspace(500e-9, 900e-9, n)
for _ in range(3):
    img = lentil.propagate_dft_batch(p, wls, 5e-6, (512, 512), oversample=2, weights=np.full(n, 1 / n))
print(img.sum())
