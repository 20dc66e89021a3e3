"""The reference's OWN test files for the path, run unmodified against the real package with its transform rebound to
the CUDA library (lentil_b200.patch.enable — what INTEGRATION.md section 2 tells a lentil user to do).  The files are
staged beside the package by oracle/build_ref.sh (sha256-pinned, git-ignored); they run in a subprocess from oracle/_ref
because their conftest globs tests/fixtures relative to the working directory."""
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref")
FILES = ["tests/test_fourier.py", "tests/test_propagate.py", "tests/test_propagate_fft.py", "tests/test_propagate_mask.py",
         "tests/test_propagate_slice.py", "tests/test_plane.py", "tests/test_wavefront.py", "tests/test_field.py",
         "tests/test_wfe.py", "tests/test_util.py", "tests/test_detector.py", "tests/test_helper.py"]


@pytest.mark.parametrize("level", ["fourier", "path"])
def test_reference_test_suite_passes_on_the_patched_package(level):
    if not os.path.isfile(os.path.join(REF, "tests", "test_fourier.py")):
        pytest.skip("reference tests not staged (run oracle/build_ref.sh where /root/reference exists)")
    env = dict(os.environ, LFD_PATCH_LEVEL=level,
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests"), ROOT, os.environ.get("PYTHONPATH", "")]))
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "ref_patch_plugin", "-p", "no:cacheprovider"] + FILES,
                         cwd=REF, env=env, capture_output=True, text=True, timeout=900)
    tail = res.stdout[-3000:] + res.stderr[-2000:]
    assert res.returncode == 0, tail
    m = re.search(r"(\d+) kernels launched by liblentil_b200", res.stdout)
    assert m and int(m.group(1)) > 50, tail                  # the reference's tests did run their transforms on the GPU
    assert re.search(r"\b(\d+) passed", res.stdout), tail


# np.array_equal against `amp * np.exp(2j pi opd / lambda)`: the mirror forms the phase in cycles and reduces it exactly before
# the sine / cosine (more accurate than numpy's radians for phases of 1e2 .. 1e4 rad), so the last bits differ; the same
# products are compared at 1e-15 in tests/test_gpu_propagate.py::test_plane_multiply_phasor
BITWISE_VS_NUMPY_EXP = ["test_wavefront_plane_mul", "test_wavefront_plane_rmul", "test_wavefront_plane_imul"]


def test_reference_test_suite_passes_on_the_mirror_classes():
    """the same files with `import lentil` resolving to a shim whose path objects are lentil_b200's own classes and functions
    (tests/ref_mirror_plugin.py): Plane / Pupil / Wavefront / Field, propagate_dft / propagate_fft, fourier, helper, rebin,
    rescale, pixel, pixelate, power_spectrum — everything else (shape generators, Zernikes, util) stays the real package's"""
    if not os.path.isfile(os.path.join(REF, "tests", "test_fourier.py")):
        pytest.skip("reference tests not staged (run oracle/build_ref.sh where /root/reference exists)")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests"), ROOT, os.environ.get("PYTHONPATH", "")]))
    deselect = " and ".join(f"not {name}" for name in BITWISE_VS_NUMPY_EXP)
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "ref_mirror_plugin", "-p", "no:cacheprovider", "-k", deselect] + FILES,
                         cwd=REF, env=env, capture_output=True, text=True, timeout=900)
    tail = res.stdout[-3000:] + res.stderr[-2000:]
    assert res.returncode == 0, tail
    m = re.search(r"(\d+) kernels launched by liblentil_b200", res.stdout)
    assert m and int(m.group(1)) > 50, tail
    m = re.search(r"\b(\d+) passed", res.stdout)
    assert m and int(m.group(1)) >= 105, tail
