"""BASELINE configs[4] as a STRONG-scaling run: 4096^2 annular pupil with a
power-spectrum WFE -> 1024^2 detector x oversample 2, L wavelengths x 16 field points, the wavelengths dealt to
the ranks by propagate_dft_batch(distributed=True) and the (16, 2048, 2048) float64 stack all-reduced over NCCL.

    python scripts/cfg5_scaling.py --wavelengths 1000                               # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/cfg5_scaling.py --wavelengths 1000                                  # 8 GPUs

Prints one JSON line on rank 0 (time = CUDA events, max over ranks, barrier on both sides)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lentil_b200 as lentil  # noqa: E402
from lentil_b200 import device, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--wavelengths", type=int, default=1000)
    ap.add_argument("--precision", default="c128")
    ap.add_argument("--out", default=None, help="append the JSON line to this file")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    mask = synth.annulus((4096, 4096), 2040)
    amp = synth.normalize_power(mask)
    opd = lentil.power_spectrum(mask, 1 / 4080, rms=30e-9, half_power_freq=5, exp=3, seed=3)
    pupil = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=1 / 4080, focal_length=20.0)
    pupil.freeze()
    tilts = [[rx, ry] for rx in np.linspace(-20e-6, 20e-6, 4) for ry in np.linspace(-20e-6, 20e-6, 4)]
    L = args.wavelengths
    wl = np.linspace(500e-9, 900e-9, L)

    def run(wls):
        return lentil.propagate_dft_batch(pupil, wls, 5e-6, (1024, 1024), oversample=2, weights=np.full(len(wls), 1.0 / len(wls)),
                                          tilts=tilts, distributed=world > 1, return_device=True, precision=args.precision)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    run(wl[:: max(L // (2 * world), 1)][: 2 * world])           # warm-up: two wavelengths per rank
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    stack = run(wl)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=stack.device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    planes = L * len(tilts)
    if rank == 0:
        line = {"workload": "BASELINE configs[4]: 4096^2 annular pupil + power-spectrum WFE -> 1024^2 det x os2, "
                            f"{L} wavelengths x 16 field points, strong scaling", "n_gpus": world, "planes": planes,
                "seconds": ms * 1e-3, "planes_per_s": planes / (ms * 1e-3), "precision": args.precision,
                "tflops_algorithmic": planes * 412.3e9 / (ms * 1e-3) / 1e12,
                "stack_shape": list(stack.shape), "stack_sum": float(stack.sum()), "stack_max": float(stack.max())}
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as fh:
                fh.write(json.dumps(line) + "\n")
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
