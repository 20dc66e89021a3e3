import cProfile, pstats, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np, torch
import lentil_b200 as lentil, bench
w = bench.WORKLOAD
amp, opd, wls, wts = bench.make_inputs(w["nlam"])
p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=w["dx"], focal_length=w["z"]); p.freeze()
def loop():
    img = lentil.device.zeros_f64(1024, 1024)
    for wl, wt in zip(wls, wts):
        wf = lentil.Wavefront(wl) * p
        wf = lentil.propagate_dft(wf, pixelscale=w["du"], shape=(w["det"],) * 2, oversample=w["oversample"])
        img = wf.insert(img, wt)
    return img
loop(); torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); loop(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
