"""Sparse complex field: a complex128 2-D array plus an integer (row, col) offset and a list of
tilts.  Mirror of lentil/field.py (Field :7-194, boundary :196-229, insert :231-305, merge
:308-386, overlap :389-410, reduce :413-461) with the array arithmetic moved to the device:

* `Field.data` is a lazily materialised numpy view of a device buffer — fields produced by
  `Plane.__mul__` / `propagate_dft` stay in HBM until somebody looks at them;
* array products go through `lfd_field_mul`, dense accumulation through `lfd_accum_intensity`
  / `lfd_accum_field` (C ABI, include/lentil_b200.h).

The integer geometry (extents, window placement, overlap grouping) is host logic and is kept
bit-compatible with the reference.
"""
import ctypes as C
import sys
from itertools import combinations

import numpy as np

from . import _lib, device, extent as _extent


class Field:
    """Two-dimensional discretely sampled complex field (lentil/field.py:7-56).

    Parameters
    ----------
    data : array_like or device tensor
        Sampled field.  Real input is cast to complex128 (field.py:35).  A complex128 CUDA
        tensor (2-D, unit column stride) is adopted without a copy.
    pixelscale : float or None
    offset : (2,) ints or None — shift of the field centre from (0, 0) in (row, col)
    tilt : list of objects implementing ``__shift__(xs, ys, z, wavelength)``
    """
    __slots__ = ('_host', '_dev', '_shape', 'offset', 'tilt', 'pixelscale', 'extent')

    def __init__(self, data, pixelscale=None, offset=None, tilt=None):
        if device.is_dev(data):
            self._dev, self._host = data, None
            self._shape = tuple(int(v) for v in data.shape)
        else:
            self._host = np.asarray(data, dtype=complex)
            self._dev = None
            self._shape = self._host.shape
        self.pixelscale = pixelscale
        self.offset = offset if offset is not None else [0, 0]
        self.tilt = tilt if tilt else []
        self.extent = _extent.array_extent(self.shape, self.offset)

    # -- storage ----------------------------------------------------------------------------------
    @property
    def data(self):
        """complex128 ndarray.  Downloading hands ownership to the host copy (so in-place edits
        by the caller are honoured); the device copy is re-created on the next device use."""
        if self._host is None:
            self._host = device.to_host(self._dev)
            self._dev = None
        return self._host

    @data.setter
    def data(self, value):
        self._host = np.asarray(value, dtype=complex)
        self._dev = None
        self._shape = self._host.shape

    @property
    def dev(self):
        """complex128 device tensor (uploaded on first use)."""
        if self._dev is None:
            self._dev = device.to_dev(self._host, dtype=np.complex128)
        return self._dev

    @property
    def on_device(self):
        return self._dev is not None

    @property
    def shape(self):
        return self._shape

    @property
    def size(self):
        n = 1
        for v in self._shape:
            n *= v
        return n

    # -- arithmetic ---------------------------------------------------------------------------------
    def __mul__(self, other):
        """Element-wise product on the rectangle intersection (lentil/field.py:80-147).

        Scalars (size-1 fields, the default planar wavefront) broadcast and inherit the other
        operand's offset (field.py:464-484); two scalars multiply only at equal offsets
        (field.py:118-128).  An empty result has ``size == 0``."""
        tilt = self.tilt + other.tilt
        if self.size == 1 and other.size == 1:
            if np.array_equal(self.offset, other.offset):
                return Field(self.data * other.data, offset=self.offset, tilt=tilt)
            return Field([], offset=None, tilt=tilt)
        if self.size == 1 or other.size == 1:
            arr, sc = (other, self) if self.size == 1 else (self, other)
            s = complex(np.asarray(sc.data).reshape(-1)[0])
            if s == 1.0:
                return Field(arr._storage(), offset=arr.offset, tilt=tilt)
            return Field(_dev_mul(arr.dev, None, s), offset=arr.offset, tilt=tilt)
        ea = _extent.array_extent(self.shape, self.offset)
        eb = _extent.array_extent(other.shape, other.offset)
        if not _extent.intersect(ea, eb):
            return Field([], offset=None, tilt=tilt)
        sa, sb = _extent.intersection_slices(ea, eb)
        prod = _dev_mul(self.dev[sa], other.dev[sb], 1.0)
        return Field(prod, offset=_extent.intersection_shift(ea, eb), tilt=tilt)

    def _storage(self):
        return self._dev if self._dev is not None else self._host

    def shift(self, z, wavelength, pixelscale, oversample, indexing='ij'):
        """Pixel shift of the field centre caused by its tilts (lentil/field.py:149-194)."""
        if indexing not in ('xy', 'ij'):
            raise ValueError("Valid values for `indexing` are 'xy' and 'ij'")
        if pixelscale is None:
            raise ValueError('pixelscale must be defined to compute shift')
        x, y = 0, 0
        for t in self.tilt:
            x, y = t.__shift__(xs=x, ys=y, z=z, wavelength=wavelength)
        if np.ndim(pixelscale) == 0:
            pixelscale = (pixelscale, pixelscale)
        out = x / pixelscale[0] * oversample, y / pixelscale[1] * oversample
        if indexing == 'ij':
            out = -out[1], out[0]
        return out


def _dev_mul(a, b, scalar):
    """out = a * b * scalar on the device (b may be None)."""
    h, w = int(a.shape[0]), int(a.shape[1])
    out = device.empty_c128(h, w)
    s = complex(scalar)
    rc = _lib.lib().lfd_field_mul(a.data_ptr(), device.ld_of(a),
                                  b.data_ptr() if b is not None else None,
                                  device.ld_of(b) if b is not None else 0,
                                  s.real, s.imag, out.data_ptr(), w, h, w, device.stream_ptr())
    _lib.check(rc, "lfd_field_mul")
    return out


def boundary(fields):
    """Bounding extent of several fields (lentil/field.py:196-229; rmax/cmax start at 0 like
    the reference, so all-negative sets are padded out to the axis)."""
    rmin, rmax, cmin, cmax = sys.maxsize, 0, sys.maxsize, 0
    for f in fields:
        e = f.extent
        rmin, rmax = min(rmin, e[0]), max(rmax, e[1])
        cmin, cmax = min(cmin, e[2]), max(cmax, e[3])
    return rmin, rmax, cmin, cmax


def _placement(field_shape, field_offset, out_shape):
    """Upper-left corner of a field inside a dense array: out//2 - field//2 + offset
    (lentil/field.py:267-268).  Clipping happens in the kernel."""
    return (int(out_shape[0]) // 2 - int(field_shape[0]) // 2 + int(field_offset[0]),
            int(out_shape[1]) // 2 - int(field_shape[1]) // 2 + int(field_offset[1]))


def _window_array(fields, out_shape, groups, weights):
    """Pack fields into lfd_window descriptors."""
    wins = (_lib.Window * len(fields))()
    keep = []
    for k, f in enumerate(fields):
        d = f.dev
        keep.append(d)
        r0, c0 = _placement(f.shape, f.offset, out_shape)
        w = wins[k]
        w.E, w.ld = d.data_ptr(), device.ld_of(d)
        w.h, w.w = int(f.shape[0]), int(f.shape[1])
        w.r0, w.c0 = r0, c0
        w.group = int(groups[k])
        w.weight = float(weights[k])
    return wins, keep


def accumulate_intensity(fields, out_dev, groups=None, weights=None):
    """out_dev += sum_groups weight * |sum of the group's fields|^2  (K3).  `fields` holds
    non-scalar Fields; `groups` (non-decreasing ints) marks coherent sets."""
    fields = [f for f in fields if f.size > 0]
    if not fields:
        return out_dev
    n = len(fields)
    groups = [0] * n if groups is None else groups
    weights = [1.0] * n if weights is None else weights
    H, W = int(out_dev.shape[0]), int(out_dev.shape[1])
    wins, keep = _window_array(fields, (H, W), groups, weights)
    scratch = device.empty_bytes(C.sizeof(_lib.Window) * n)
    rc = _lib.lib().lfd_accum_intensity(wins, n, out_dev.data_ptr(), H, W, device.ld_of(out_dev),
                                        scratch.data_ptr(), scratch.numel(), device.stream_ptr())
    _lib.check(rc, "lfd_accum_intensity")
    return out_dev


# bench.py sets this to a list to get (start event, end event, algorithmic bytes) per K3 launch
TIMERS = None


def accumulate_windows(wins, out_dev):
    """K3 on a prepared numpy table of lfd_window rows (dtype np.dtype(_lib.Window))."""
    n = len(wins)
    if n == 0:
        return out_dev
    H, W = int(out_dev.shape[0]), int(out_dev.shape[1])
    scratch = device.empty_bytes(C.sizeof(_lib.Window) * n)
    if TIMERS is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = _lib.lib().lfd_accum_intensity(wins.ctypes.data_as(C.POINTER(_lib.Window)), n, out_dev.data_ptr(),
                                        H, W, device.ld_of(out_dev), scratch.data_ptr(), scratch.numel(),
                                        device.stream_ptr())
    _lib.check(rc, "lfd_accum_intensity")
    if TIMERS is not None:
        e1.record()
        esize = np.where(wins['c64'] == 2, 8, np.where(wins['c64'] == 1, 8, 16))     # float64 intensity / complex64 / complex128
        TIMERS.append((e0, e1, float(np.sum(wins['h'].astype(np.int64) * wins['w'] * esize) + 16.0 * H * W)))
    return out_dev


def accumulate_field(fields, out_dev, weights=None):
    """out_dev (complex) += weight * field  for each field (lentil/field.py:303-304)."""
    fields = [f for f in fields if f.size > 0]
    if not fields:
        return out_dev
    n = len(fields)
    weights = [1.0] * n if weights is None else weights
    H, W = int(out_dev.shape[0]), int(out_dev.shape[1])
    wins, keep = _window_array(fields, (H, W), list(range(n)), weights)
    scratch = device.empty_bytes(C.sizeof(_lib.Window) * n)
    rc = _lib.lib().lfd_accum_field(wins, n, out_dev.data_ptr(), H, W, device.ld_of(out_dev),
                                    scratch.data_ptr(), scratch.numel(), device.stream_ptr())
    _lib.check(rc, "lfd_accum_field")
    return out_dev


def insert(field, out, intensity=False, weight=1):
    """Insert a field into a dense array, clipped to the array (lentil/field.py:231-305).

    `out` may be a numpy array (updated in place and returned, as in the reference) or a device
    tensor (accumulated in HBM with no transfer)."""
    if field.size == 0:
        return out
    if field.size == 1 and len(field.shape) < 2:
        # default planar field: a scalar broadcast over the whole array when shapes agree
        if field.shape == out.shape:
            v = complex(np.asarray(field.data).reshape(-1)[0])
            out += (abs(v ** 2) if intensity else v) * weight
            return out
        raise ValueError("cannot insert a scalar field into an array")
    if device.is_dev(out):
        if intensity:
            return accumulate_intensity([field], out, [0], [weight])
        return accumulate_field([field], out, [weight])
    tmp = device.zeros_f64(*out.shape) if intensity else device.empty_c128(*out.shape).zero_()
    if intensity:
        accumulate_intensity([field], tmp, [0], [weight])
    else:
        accumulate_field([field], tmp, [weight])
    out += device.to_host(tmp)
    return out


def overlap(fields):
    """True when the fields form one overlapping set (lentil/field.py:389-410)."""
    if len(fields) == 2:
        return _extent.intersect(fields[0].extent, fields[1].extent)
    return len(_reduce(fields)) <= 1


def _merge_geometry(fields):
    rmin, rmax, cmin, cmax = boundary(fields)
    nrow, ncol = rmax - rmin + 1, cmax - cmin + 1
    return (nrow, ncol), (rmin + nrow // 2, cmin + ncol // 2)


def _merge(fields):
    """Coherent sum into the common bounding box (lentil/field.py:331-386)."""
    if not all(np.all(f.pixelscale == fields[0].pixelscale) for f in fields):
        raise ValueError("Can't merge: pixelscales must be equal")
    shape, offset = _merge_geometry(fields)
    out = device.empty_c128(*shape).zero_()
    # place each field relative to the merged box: same corner formula with the merged offset removed
    shifted = [Field(f.dev, offset=(f.offset[0] - offset[0], f.offset[1] - offset[1])) for f in fields]
    accumulate_field(shifted, out)
    return Field(out, pixelscale=fields[0].pixelscale, offset=offset)


def merge(a, b, enforce_overlap=True):
    """Merge two fields into one spanning both (lentil/field.py:308-327)."""
    if enforce_overlap and not overlap((a, b)):
        raise ValueError("Can't merge non-overlapping fields")
    return _merge((a, b))


def _reduce(fields):
    """Group fields into disjoint sets by transitive extent overlap (lentil/field.py:440-461).
    Same pair-scan order as the reference so that group order (and thus summation order) match."""
    groups = [{'field': [f], 'extent': f.extent} for f in fields]
    merged = True
    while merged:
        merged = False
        for a, b in combinations(range(len(groups)), 2):
            if _extent.intersect(groups[a]['extent'], groups[b]['extent']):
                groups[a]['field'].extend(groups[b]['field'])
                groups[a]['extent'] = boundary(groups[a]['field'])
                groups.pop(b)
                merged = True
                break
    return groups


def reduce(fields):
    """Disjoint set of fields with overlapping ones merged (lentil/field.py:413-437)."""
    out = []
    for grp in _reduce(fields):
        out.append(_merge(grp['field']) if len(grp['field']) > 1 else grp['field'][0])
    return out
