"""Mask bounding boxes and slice offsets (mirror of lentil/helper.py:27-123 and
lentil/util.py:190-218).  Integer host logic."""
import numpy as np


def boundary(x, threshold=0):
    """(rmin, rmax, cmin, cmax) of the entries of `x` above `threshold` (lentil/util.py:190-218)."""
    hit = np.asarray(x) > threshold
    rows = np.flatnonzero(hit.any(axis=1))
    cols = np.flatnonzero(hit.any(axis=0))
    return rows[0], rows[-1], cols[0], cols[-1]


def boundary_slice(x, threshold=0, pad=(0, 0)):
    """Bounding box of the data in `x` as a pair of slices (lentil/helper.py:27-62)."""
    pad = np.broadcast_to(np.asarray(pad), (2,))
    rmin, rmax, cmin, cmax = boundary(x, threshold)
    rmin = np.max((rmin - pad[0], 0))
    rmax = np.min((rmax + pad[0] + 1, x.shape[0]))
    cmin = np.max((cmin - pad[1], 0))
    cmax = np.min((cmax + pad[1] + 1, x.shape[1]))
    return np.s_[rmin:rmax, cmin:cmax]


def slice_offset(slice, shape):
    """Offset (r, c) of the centre of a 2-D slice from the centre of the enclosing array
    (lentil/helper.py:65-123); centres are index n//2."""
    if slice == Ellipsis:
        return (0, 0)
    if Ellipsis in slice:
        if any(isinstance(s, type(np.s_[:])) and s == np.s_[:] for s in slice):
            return (0, 0)
        raise ValueError(f"Can't compute offset from slice {slice}")
    rows, cols = slice
    off = (int(rows.start + (rows.stop - rows.start) // 2 - int(shape[0]) // 2),
           int(cols.start + (cols.stop - cols.start) // 2 - int(shape[1]) // 2))
    return (0, 0) if off == (0, 0) else off


def mesh(shape, shift=(0, 0)):
    """Centred (row, col) index grids (lentil/helper.py:7-17 without rotation)."""
    rr, cc = np.meshgrid(np.arange(shape[0]) - np.floor(shape[0] / 2.0) - shift[0],
                         np.arange(shape[1]) - np.floor(shape[1] / 2.0) - shift[1], indexing='ij')
    return rr, cc


def pair(value):
    """(2,) float/int array from a scalar or a pair — what np.broadcast_to(value, (2,)) gives, without the
    stride-tricks overhead (this sits on the per-wavelength path of the drop-in loop)."""
    if isinstance(value, np.ndarray) and value.shape == (2,):
        return value
    a = np.asarray(value)
    if a.ndim == 0:
        return np.array([a, a])
    return a if a.shape == (2,) else np.broadcast_to(a, (2,))
