"""Integer rectangle algebra on inclusive extents (rmin, rmax, cmin, cmax), origin at the
optical axis.  Host-side mirror of lentil/extent.py:5-169; results must be bit-exact with the
reference because they decide which output pixels a window covers."""
import numpy as np


def array_extent(shape, shift, parent_shape=None):
    """Extent of an array of `shape` whose centre (index n//2) sits at `shift`
    (lentil/extent.py:5-40).  int() truncates toward zero like the reference; shapes with fewer
    than two dimensions (the default planar field) count as 1x1."""
    if len(shape) < 2:
        shape = (1, 1)
    nr, nc = int(shape[0]), int(shape[1])
    rmin = int(-(nr // 2) + shift[0])
    cmin = int(-(nc // 2) + shift[1])
    rmax, cmax = rmin + nr - 1, cmin + nc - 1
    if parent_shape is not None:
        pr, pc = (int(v) // 2 for v in np.broadcast_to(parent_shape, (2,)))
        rmin, rmax, cmin, cmax = rmin + pr, rmax + pr, cmin + pc, cmax + pc
    return rmin, rmax, cmin, cmax


def array_center(extent):
    """Centre (rmin + nrow//2, cmin + ncol//2) of an extent (lentil/extent.py:43-59)."""
    rmin, rmax, cmin, cmax = extent
    return rmin + (rmax - rmin + 1) // 2, cmin + (cmax - cmin + 1) // 2


def intersect(a, b):
    """True when two extents share at least one pixel (lentil/extent.py:62-77)."""
    return bool(a[0] <= b[1] and a[1] >= b[0] and a[2] <= b[3] and a[3] >= b[2])


def intersection_extent(a, b):
    """Extent of the overlap (lentil/extent.py:80-100); degenerate when they do not intersect."""
    return max(a[0], b[0]), min(a[1], b[1]), max(a[2], b[2]), min(a[3], b[3])


def intersection_shape(a, b):
    """Shape of the overlap, () when empty (lentil/extent.py:103-124)."""
    rmin, rmax, cmin, cmax = intersection_extent(a, b)
    nr, nc = rmax - rmin + 1, cmax - cmin + 1
    return (nr, nc) if (nr > 0 and nc > 0) else ()


def intersection_slices(a, b):
    """Slices selecting the overlap inside each operand (lentil/extent.py:127-150)."""
    rmin, rmax, cmin, cmax = intersection_extent(a, b)
    return ((slice(rmin - a[0], rmax - a[0] + 1), slice(cmin - a[2], cmax - a[2] + 1)),
            (slice(rmin - b[0], rmax - b[0] + 1), slice(cmin - b[2], cmax - b[2] + 1)))


def intersection_shift(a, b):
    """Centre of the overlap = offset of the product field (lentil/extent.py:153-169)."""
    return array_center(intersection_extent(a, b))
