"""lentil_b200 — B200-native (sm_100a) far-field diffraction path with lentil's call surface.

    import lentil_b200 as lentil
    w = lentil.Wavefront(650e-9) * lentil.Pupil(amplitude=amp, opd=opd, pixelscale=dx, focal_length=z)
    w = lentil.propagate_dft(w, pixelscale=5e-6, shape=(512, 512), oversample=2)
    psf = w.intensity

Only the hot path of andykee/lentil is implemented: fourier.dft2/idft2,
propagate_dft, Plane/Pupil/Image/Tilt products, Field algebra, Wavefront.intensity/insert/field.
Everything array-sized runs in hand-written CUDA behind the C ABI of include/lentil_b200.h;
there is no CPU fallback.
"""
__version__ = '0.1.0'

from .ptype import ptype, none, pupil, image, tilt, transform  # noqa: F401
from . import extent, helper, field, fourier, plane, propagate, wavefront, device, detector, wfe, patch  # noqa: F401
from .field import Field  # noqa: F401
from .plane import Plane, Pupil, Image, Tilt, DispersiveTilt, Grism  # noqa: F401
from .wavefront import Wavefront  # noqa: F401
from .propagate import propagate_dft, propagate_dft_batch, propagate_fft, scratch_shape  # noqa: F401
from .helper import boundary  # noqa: F401
from .device import set_device  # noqa: F401
from .detector import rebin, rescale  # noqa: F401
from .wfe import power_spectrum  # noqa: F401
