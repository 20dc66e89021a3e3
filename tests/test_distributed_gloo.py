"""The N>1 path of propagate_dft_batch on CPU: wavelength sharding + the final PSF-stack
reduction, world_size 2 over gloo.  The per-rank planes are computed by the oracle here (no GPU in
this container); on the B200 box the same two helpers wrap K1/K2a/K3 and NCCL."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lentil_oracle as oc
    from lentil_b200 import synth
    from lentil_b200.propagate import shard_indices, reduce_stack

    mask = synth.annulus((48, 48), 20)
    amp = synth.normalize_power(mask)
    opd = synth.zernike_opd(mask, np.random.default_rng(0).normal(size=6) * 30e-9)
    wls = np.linspace(500e-9, 900e-9, 7)
    wts = np.linspace(0.5, 1.5, 7)
    mine = shard_indices(len(wls), distributed=True)
    assert list(mine) == list(range(rank, len(wls), world))
    local = oc.psf(amp, opd, None, wls[mine], wts[mine], (1 / 40, 1 / 40), 10.0, 5e-6, (16, 16), None, 2)
    stack = torch.from_numpy(local.copy())
    reduce_stack(stack)
    full = oc.psf(amp, opd, None, wls, wts, (1 / 40, 1 / 40), 10.0, 5e-6, (16, 16), None, 2)
    err = float(np.max(np.abs(stack.numpy() - full)) / np.max(full))
    ret[rank] = err
    dist.destroy_process_group()


def test_wavelength_sharding_and_stack_reduce_gloo():
    world = 2
    port = 29500 + (os.getpid() % 500)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            assert ret[r] <= 1e-14, ret[r]


def test_shard_indices_partition():
    from lentil_b200.propagate import shard_indices
    for L in (1, 7, 100, 1000):
        for world in (1, 2, 4, 8):
            parts = [shard_indices(L, True, rank=r, world=world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts))
            assert np.array_equal(allidx, np.arange(L))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert np.array_equal(shard_indices(5), np.arange(5))
