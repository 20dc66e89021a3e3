"""pytest plugin used by tests/test_gpu_reference_own_tests.py (mirror mode): the reference's own test files import
`lentil`; here that name resolves to a shim whose PATH objects are lentil_b200's (Plane / Pupil / Image / Tilt / Wavefront /
Field, propagate_dft / propagate_fft, the fourier, field, helper, extent, propagate, plane and wavefront modules, rebin /
rescale, detector.pixel / pixelate, power_spectrum) while everything outside the path (shape generators, Zernikes, util,
radiometry, detector noise models) stays the real package's.  So the reference's tests exercise the mirror classes themselves."""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
for p in (_ROOT, os.path.join(_ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

_state = {}

TOP = ["Plane", "Pupil", "Image", "Tilt", "DispersiveTilt", "Grism", "Wavefront", "Field", "propagate_dft", "propagate_fft",
       "scratch_shape", "rebin", "rescale", "power_spectrum", "ptype", "none", "pupil", "image", "tilt", "transform", "boundary"]
MODULES = ["fourier", "field", "helper", "extent", "propagate", "plane", "wavefront"]


def _install():
    """at plugin IMPORT time (-p ...): the reference's conftest imports its fixture modules, which `import lentil`, before
    pytest_configure runs"""
    import ref_loader
    ref = ref_loader.reference()
    assert ref is not None, "no staged reference under oracle/_ref"
    import lentil_b200 as ours
    from lentil_b200 import device
    shim = types.ModuleType("lentil")
    shim.__dict__.update({k: v for k, v in ref.__dict__.items() if not k.startswith("__")})
    shim.__path__ = ref.__path__
    used = []
    for name in TOP:
        if hasattr(ours, name):
            setattr(shim, name, getattr(ours, name))
            used.append(name)
    for name in MODULES:
        mod = getattr(ours, name, None)
        if isinstance(mod, types.ModuleType):
            setattr(shim, name, mod)
            sys.modules["lentil." + name] = mod
            used.append(name)
    det = types.ModuleType("lentil.detector")
    det.__dict__.update({k: v for k, v in ref.detector.__dict__.items() if not k.startswith("__")})
    det.pixel, det.pixelate = ours.detector.pixel, ours.detector.pixelate
    shim.detector = det
    sys.modules["lentil.detector"] = det
    sys.modules["lentil"] = shim
    _state.update(n0=device.launch_count(), used=used)


_install()


def pytest_terminal_summary(terminalreporter):
    from lentil_b200 import device
    launched = device.launch_count() - _state.get("n0", 0)
    terminalreporter.write_line(f"lentil_b200 mirror ({', '.join(_state.get('used', []))}): {launched} kernels launched by liblentil_b200")
