"""Monochromatic wavefront: a few scalars plus a list of sparse Fields (same public surface as
lentil/wavefront.py:12-186 — ``Wavefront(wavelength, pixelscale, diameter, focal_length, tilt,
ptype)``, ``.data``, ``.field``, ``.intensity``, ``.insert``, ``Wavefront.empty``).

Dense views are produced by K3 (lfd_accum_field / lfd_accum_intensity).  All Fields of one
wavefront form a single coherent group: windows that overlap add as complex amplitudes before the
modulus is squared, which is what lentil.field.reduce followed by insert computes (disjoint
windows cannot interfere, so grouping them changes nothing).
"""
import importlib

import numpy as np

from . import device
from . import helper as _helper
from . import field as _field
from .field import Field

_pt = importlib.import_module(".ptype", __package__)   # the package attribute `ptype` is the factory function

_WAVEFRONT_PTYPES = (_pt.none, _pt.pupil, _pt.image)


def _as_pair(value):
    return None if value is None else _helper.pair(value)


class Wavefront:
    """One wavelength of light on its way through the planes of an optical system.

    Parameters
    ----------
    wavelength : float
        Metres.
    pixelscale : float or (2,), optional
        Sampling of the wavefront; filled in by the first plane that has one.
    diameter : float, optional
    focal_length : float, optional
        ``None`` (or 0) is a plane wave; a Pupil hands its own focal length over.
    tilt : (rx, ry), optional
        Field point: radians of tilt about the x and y axes, booked as a ``Tilt`` on the
        initial planar Field and turned into a pixel shift by ``propagate_dft``.
    ptype : ptype, optional
        Only none / pupil / image are legal for a wavefront (TypeError otherwise).
    """

    def __init__(self, wavelength, pixelscale=None, diameter=None, focal_length=None, tilt=None,
                 ptype=None):
        self._wavelength = wavelength
        self._pixelscale = _as_pair(pixelscale)
        self.diameter = diameter
        self.focal_length = focal_length or None
        self.shape = ()
        self.ptype = ptype                                  # validated by the setter
        self.data = [Field(np.array(1, dtype=complex), tilt=self._field_point(tilt))]

    @staticmethod
    def _field_point(tilt):
        if tilt is None:
            return None
        if len(tilt) != 2:
            raise ValueError('tilt must be specified as [rx, ry]')
        from .plane import Tilt
        return [Tilt(x=tilt[0], y=tilt[1])]

    @classmethod
    def empty(cls, wavelength, pixelscale=None, diameter=None, focal_length=None, tilt=None,
              shape=None, ptype=None):
        """A wavefront with the given metadata and no Fields — what planes and propagations
        fill and return."""
        new = cls(wavelength, pixelscale, diameter, focal_length, tilt, ptype)
        new.shape = shape if shape is not None else ()
        new.data = []
        return new

    # ---- read-only metadata ---------------------------------------------------------------------
    wavelength = property(lambda self: self._wavelength, doc="Wavelength in metres.")
    pixelscale = property(lambda self: self._pixelscale, doc="(2,) sampling of the wavefront, or None.")

    @property
    def ptype(self):
        return self._ptype

    @ptype.setter
    def ptype(self, value):
        tag = _pt.ptype(value)
        if tag not in _WAVEFRONT_PTYPES:
            raise TypeError(f"invalid ptype '{value}' for Wavefront")
        self._ptype = tag

    # ---- products: the plane does the work --------------------------------------------------------
    def __mul__(self, plane):
        return plane.__mul__(self)

    __rmul__ = __mul__

    # ---- dense views ----------------------------------------------------------------------------
    def _arrays(self):
        """Fields that carry an array (the default planar Field is a 0-d scalar)."""
        return [f for f in self.data if len(f.shape) == 2 and f.size > 0]

    def _accumulate(self, out, weight, intensity):
        """Add this wavefront into `out` (device tensor, modified in place)."""
        fields = self._arrays()
        if fields:
            if intensity:
                _field.accumulate_intensity(fields, out, [0] * len(fields), [weight] * len(fields))
            else:
                _field.accumulate_field(fields, out, [weight] * len(fields))
        return out

    def _dense(self, intensity):
        scalars = [f for f in self.data if len(f.shape) < 2]
        if scalars and len(scalars) == len(self.data):
            out = np.zeros(self.shape, dtype=float if intensity else complex)   # planar wavefront, no kernel
            for f in scalars:
                out = _field.insert(f, out, intensity=intensity)
            return out
        out = device.zeros_f64(*self.shape) if intensity else device.empty_c128(*self.shape).zero_()
        return device.to_host(self._accumulate(out, 1, intensity))

    @property
    def field(self):
        """Complex field on the full ``shape`` grid."""
        return self._dense(intensity=False)

    @property
    def intensity(self):
        """|field|^2 on the full ``shape`` grid (overlapping Fields add coherently first)."""
        return self._dense(intensity=True)

    def insert(self, out, weight=1):
        """``out += weight * intensity`` and return ``out``.

        ``out`` may be a numpy array (one device pass, one D2H, then the in-place add the
        reference performs) or a float64 device tensor, in which case nothing leaves HBM — the form
        to use inside a wavelength loop."""
        if not self._arrays():
            return out
        if device.is_dev(out):
            return self._accumulate(out, weight, True)
        out += device.to_host(self._accumulate(device.zeros_f64(*out.shape), weight, True))
        return out


def _overlap(field_shape, field_shift, output_shape):
    """Does a field of `field_shape`, shifted by `field_shift`, touch an output array at all?
    (lentil/wavefront.py:168-186; not used by the propagation path, kept because lentil tests it.)"""
    lo = (np.asarray(output_shape) - np.asarray(field_shape)) / 2 + np.asarray(field_shift)
    hi = lo + np.asarray(field_shape)
    return bool(np.all(lo <= np.asarray(output_shape)) and np.all(hi >= 0))
