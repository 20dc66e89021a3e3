"""The N > 1 path on hardware: two ranks, one GPU each, NCCL.  propagate_dft_batch(distributed=True) deals the
wavelengths to the ranks, every rank runs K1 / K2a / K3 on its share and the PSF stacks are summed with one
all-reduce (lentil_b200/propagate.py: shard_indices, reduce_stack); the result on every rank must equal the
oracle's polychromatic PSF of ALL wavelengths.  Skipped on a box with a single GPU (the CPU suite covers the same
host logic over gloo, tests/test_distributed_gloo.py; bench.py --gpus N reports `parity_distributed`)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, TOL64

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import lentil_oracle as oc
    import lentil_b200 as lentil
    from lentil_b200 import device, synth
    device.set_device(rank)

    mask = synth.annulus((160, 160), 70, 0.25)
    amp = synth.normalize_power(mask)
    opd = synth.zernike_opd(mask, np.random.default_rng(0).normal(size=8) * 30e-9)
    dx, z, du = 1 / 140, 20.0, 5e-6
    wls = np.linspace(500e-9, 900e-9, 7)
    wts = np.linspace(0.5, 1.5, 7)
    tilts = [[0.0, 0.0], [6e-6, -4e-6]]
    p = lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    errs = []
    for execution in ("auto", "folded"):
        stack = lentil.propagate_dft_batch(p, wls, du, (64, 64), oversample=2, weights=wts, tilts=tilts, distributed=True,
                                           execution=execution)
        for k, t in enumerate(tilts):
            full = oc.psf(amp, opd, None, wls, wts, (dx, dx), z, du, (64, 64), None, 2, wf_tilt=t)
            errs.append(float(np.max(np.abs(stack[k] - full)) / np.max(full)))
    # accumulate-into semantics with a pre-filled `out`: the previous contents are kept once, not once per rank
    out = torch.full((2, 128, 128), 0.25, dtype=torch.float64, device=device.device())
    lentil.propagate_dft_batch(p, wls, du, (64, 64), oversample=2, weights=wts, tilts=tilts, distributed=True, out=out,
                               return_device=True)
    errs.append(float((out - 0.25 - torch.from_numpy(stack).to(out.device)).abs().max() / stack.max()))
    ret[rank] = max(errs)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_ranks_nccl_psf_equals_oracle():
    import torch.multiprocessing as mp
    world = 2
    port = 29600 + (os.getpid() % 300)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            assert ret[r] <= TOL64, dict(ret)
