"""One process, two GPUs: kernel attributes (cudaFuncSetAttribute is per device), FFT root tables, SM counts and the
workspace cache are kept per device, so switching with lentil_b200.set_device must not break the > 48 KB shared-memory
kernels (direct, folded, chirp-z with L >= 2048, tcgen05) on the second device."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import lentil_oracle as oc  # noqa: E402
from conftest import TOL64, peak_err  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_every_execution_on_two_devices_in_one_process():
    import lentil_b200 as lentil
    from lentil_b200 import device
    rng = np.random.default_rng(3)
    f = rng.normal(size=(1001, 300)) + 1j * rng.normal(size=(1001, 300))
    kw = dict(shape=(1024, 280), shift=(0.25, -1.5), offset=(3, -2))
    ref = oc.dft2(f, (4.9e-4, 1.7e-3), **kw)
    try:
        for dev in (0, 1, 0):
            device.set_device(dev)
            for execution in ("czt", "folded", "direct"):
                got = lentil.fourier.dft2(f, (4.9e-4, 1.7e-3), execution=execution, **kw)
                assert peak_err(got, ref) <= TOL64, (dev, execution)
            for execution in ("czt", "folded"):
                got = lentil.fourier.dft2_c64(f.astype(np.complex64), (4.9e-4, 1.7e-3), execution=execution, **kw)
                assert peak_err(got, ref) <= 1e-5, (dev, execution)
    finally:
        device.set_device(0)
