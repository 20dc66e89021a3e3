"""pytest plugin used by tests/test_gpu_reference_own_tests.py: before the reference's own test files are collected, rebind
the staged reference package's hot-path entry points to the CUDA library (lentil_b200.patch.enable), and at the end of the
session report how many kernels the library launched (proof that the transforms ran on the GPU).

    cd oracle/_ref && python -m pytest -p ref_patch_plugin tests/test_fourier.py ...

LFD_PATCH_LEVEL = fourier (default: only lentil.fourier.dft2 / idft2) | path (also rebin / rescale / pixel / pixelate; the
reference's propagate_dft needs lentil_b200 Wavefronts at that level, so it is restored for these files)."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for p in (_ROOT, os.path.join(_ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

_state = {}


def pytest_configure(config):
    import ref_loader
    ref = ref_loader.reference()
    assert ref is not None, "no staged reference under oracle/_ref"
    import lentil_b200
    from lentil_b200 import patch, device
    level = os.environ.get("LFD_PATCH_LEVEL", "fourier")
    patch.enable(ref, level=level)
    if level == "path":           # the reference's tests build reference Wavefronts: keep its own driver, ours below it
        saved = patch._SAVED[id(ref)]
        ref.propagate_dft = saved[(None, "propagate_dft")]
        ref.propagate.propagate_dft = saved[("propagate", "propagate_dft")]
    assert ref.fourier.dft2 is lentil_b200.fourier.dft2
    _state.update(ref=ref, n0=device.launch_count(), level=level)


def pytest_terminal_summary(terminalreporter):
    from lentil_b200 import device
    launched = device.launch_count() - _state.get("n0", 0)
    terminalreporter.write_line(f"lentil_b200 patch level={_state.get('level')}: {launched} kernels launched by liblentil_b200")
