"""Monochromatic wavefront = metadata + a list of sparse Fields (mirror of lentil/wavefront.py:12-186).

`field`, `intensity` and `insert` materialise dense arrays with K3 (lfd_accum_field /
lfd_accum_intensity): all fields of the wavefront are one coherent group, which is what
lentil.field.reduce + insert compute (overlapping windows are summed as complex amplitudes,
disjoint ones cannot interfere).
"""
import numpy as np

from . import device
from . import field as _field
import importlib
_pt = importlib.import_module(".ptype", __package__)  # the package attribute `ptype` is the factory function
from .field import Field


class Wavefront:
    """A monochromatic wavefront (lentil/wavefront.py:12-56).

    Parameters
    ----------
    wavelength : float — metres
    pixelscale : float, optional
    diameter : float, optional
    focal_length : float or None — None is a plane wave
    tilt : (2,) array_like, optional — radians about x and y, ``[rx, ry]`` (a field point)
    ptype : ptype, optional
    """

    def __init__(self, wavelength, pixelscale=None, diameter=None, focal_length=None, tilt=None,
                 ptype=None):
        from .plane import Tilt
        self.focal_length = focal_length if focal_length else None
        self.diameter = diameter
        self.shape = ()
        self._wavelength = wavelength
        self._pixelscale = None if pixelscale is None else np.broadcast_to(pixelscale, (2,))
        self.ptype = _pt.ptype(ptype)
        if tilt is not None:
            if len(tilt) != 2:
                raise ValueError('tilt must be specified as [rx, ry]')
            tilt = [Tilt(x=tilt[0], y=tilt[1])]
        self.data = [Field(data=np.array(1, dtype=complex), offset=None, tilt=tilt)]

    def __mul__(self, plane):
        return plane.__mul__(self)

    def __rmul__(self, other):
        return self.__mul__(other)

    @property
    def wavelength(self):
        return self._wavelength

    @property
    def pixelscale(self):
        return self._pixelscale

    @property
    def ptype(self):
        return self._ptype

    @ptype.setter
    def ptype(self, value):
        if _pt.ptype(value) not in (_pt.none, _pt.pupil, _pt.image):
            raise TypeError(f"invalid ptype '{value}' for Wavefront")
        self._ptype = _pt.ptype(value)

    # ---- dense views ---------------------------------------------------------------------------
    def _scalar_only(self):
        return all(len(f.shape) < 2 for f in self.data)

    @property
    def field(self):
        """Dense complex field (lentil/wavefront.py:101-112)."""
        if self._scalar_only():
            out = np.zeros(self.shape, dtype=complex)
            for f in self.data:
                out = _field.insert(f, out)
            return out
        out = device.empty_c128(*self.shape).zero_()
        _field.accumulate_field(self.data, out)
        return device.to_host(out)

    @property
    def intensity(self):
        """Dense intensity (lentil/wavefront.py:114-125)."""
        if self._scalar_only():
            out = np.zeros(self.shape, dtype=float)
            for f in self.data:
                out = _field.insert(f, out, intensity=True)
            return out
        return device.to_host(self.insert(device.zeros_f64(*self.shape)))

    @classmethod
    def empty(cls, wavelength, pixelscale=None, diameter=None, focal_length=None, tilt=None,
              shape=None, ptype=None):
        """A wavefront with no data (lentil/wavefront.py:127-142)."""
        w = cls(wavelength=wavelength, pixelscale=pixelscale, diameter=diameter,
                focal_length=focal_length, tilt=tilt, ptype=ptype)
        w.data = []
        w.shape = () if shape is None else shape
        return w

    def insert(self, out, weight=1):
        """Accumulate weight * intensity into `out` (lentil/wavefront.py:145-165).

        `out` may be a numpy array (updated in place through one device pass and a D2H) or a
        float64 device tensor, in which case nothing leaves HBM — the form to use inside a
        wavelength loop."""
        fields = [f for f in self.data if f.size > 0]
        if not fields:
            return out
        if device.is_dev(out):
            return _field.accumulate_intensity(fields, out, [0] * len(fields), [weight] * len(fields))
        tmp = device.zeros_f64(*out.shape)
        _field.accumulate_intensity(fields, tmp, [0] * len(fields), [weight] * len(fields))
        out += device.to_host(tmp)
        return out


def _overlap(field_shape, field_shift, output_shape):
    """True when a shifted field touches the output array (lentil/wavefront.py:168-186; unused
    by the propagation path, kept because the reference tests it)."""
    output_shape = np.asarray(output_shape)
    field_shape = np.asarray(field_shape)
    ul = (output_shape / 2) - (field_shape / 2) + np.asarray(field_shift)
    if ul[0] > output_shape[0] or ul[0] + field_shape[0] < 0:
        return False
    if ul[1] > output_shape[1] or ul[1] + field_shape[1] < 0:
        return False
    return True
