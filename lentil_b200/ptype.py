"""Plane-type tags (mirror of lentil/ptype.py:2-50: same names, equality and hashing)."""

PTYPES = ('none', 'pupil', 'image', 'tilt', 'transform')


class PType:
    """Tag carried by planes and wavefronts; two tags are equal when their names are."""
    __slots__ = ('_key',)

    def __init__(self, name):
        if name not in PTYPES:
            raise TypeError(f"plane type '{name}' not understood")
        self._key = name

    def __eq__(self, other):
        return self._key == other._key

    def __hash__(self):
        return hash(self._key)

    def __repr__(self):
        return f"ptype('{self._key}')"

    def __str__(self):
        return self._key


def ptype(value):
    """Coerce a string, PType or None (-> 'none') to a PType (lentil/ptype.py:5-20)."""
    if isinstance(value, PType):
        return value
    return PType('none' if value is None else value)


none = ptype('none')
pupil = ptype('pupil')
image = ptype('image')
tilt = ptype('tilt')
transform = ptype('transform')
