"""Reroute an installed andykee/lentil through this library.

lentil has no plugin registry; its operator API for the diffraction path is module attributes that are
looked up at call time: ``lentil.propagate.propagate_dft`` calls ``lentil.fourier.dft2`` through the module
(lentil/propagate.py:235), and the package namespace re-exports ``propagate_dft`` (lentil/__init__.py:35-39).

    import lentil, lentil_b200.patch
    lentil_b200.patch.enable(lentil)                 # dft2 / idft2 on the GPU, lentil's own Plane/Field/Wavefront code unchanged
    lentil_b200.patch.enable(lentil, level='path')   # also propagate_dft, rebin, rescale, pixel, pixelate (device kernels)
    lentil_b200.patch.disable(lentil)                # restore what was there

``level='fourier'`` keeps lentil's numpy semantics exactly (numpy in, numpy out per transform).  ``level='path'``
swaps the functions of the path that accept lentil's own objects: propagate_dft then needs ``lentil_b200`` Wavefronts
(build the model with ``lentil_b200.Pupil`` / ``Wavefront``), the detector helpers take numpy arrays as before.
"""
from . import detector as _detector
from . import fourier as _fourier
from . import propagate as _propagate

_SAVED = {}

_FOURIER = (("fourier", "dft2", _fourier.dft2), ("fourier", "idft2", _fourier.idft2))
_PATH = (("propagate", "propagate_dft", _propagate.propagate_dft), (None, "propagate_dft", _propagate.propagate_dft),
         ("util", "rebin", _detector.rebin), (None, "rebin", _detector.rebin),
         ("util", "rescale", _detector.rescale), (None, "rescale", _detector.rescale),
         ("detector", "pixel", _detector.pixel), ("detector", "pixelate", _detector.pixelate))


def enable(lentil_module, level='fourier'):
    """Rebind the hot-path entry points of ``lentil_module`` (the imported reference package) to this library."""
    if level not in ('fourier', 'path'):
        raise ValueError("level must be 'fourier' or 'path'")
    table = _FOURIER + (_PATH if level == 'path' else ())
    saved = _SAVED.setdefault(id(lentil_module), {})
    for sub, name, fn in table:
        owner = lentil_module if sub is None else getattr(lentil_module, sub, None)
        if owner is None or not hasattr(owner, name):
            continue                                    # older/newer lentil without that attribute: leave it alone
        saved.setdefault((sub, name), getattr(owner, name))
        setattr(owner, name, fn)
    return lentil_module


def disable(lentil_module):
    """Undo enable(): put back the attributes that were replaced."""
    for (sub, name), fn in _SAVED.pop(id(lentil_module), {}).items():
        setattr(lentil_module if sub is None else getattr(lentil_module, sub), name, fn)
    return lentil_module
