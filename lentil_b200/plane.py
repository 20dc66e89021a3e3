"""Planes: amplitude / OPD / mask containers whose product with a Wavefront builds the complex
phasor amp * exp(2 pi i opd / lambda) on each segment's bounding box.

Mirror of the hot part of lentil/plane.py: _PlaneBase :19-396 (properties, freeze/thaw),
Plane :414-611 (__mul__ :477-516, fit_tilt :564-611), ptype table :634-668, _plane_slice
:671-705, Pupil :708-771, Image :774-820, Tilt :884-923.  The phasor build runs on the device
(K1, lfd_pupil_prep); slicing, tilt bookkeeping and the ptype rules are host metadata.

Out of scope here: rescale/resample, DispersiveTilt/Grism, Rotate/Flip.
"""
import copy
import ctypes as C

import numpy as np

from . import _lib, device, helper
import importlib
_pt = importlib.import_module(".ptype", __package__)  # the package attribute `ptype` is the factory function
from .field import Field


class _PlaneBase:
    """Low-level plane interface (lentil/plane.py:19-396): user subclasses may override
    ``__amp__``, ``__opd__`` and ``__mask__``; they are called on the host on every access
    unless the plane is frozen."""

    def __new__(cls, *args, **kwargs):
        self = super().__new__(cls)
        self._amplitude = np.array(1)
        self._opd = np.array(0)
        self._mask = None
        self._pixelscale = None
        self._ptype = _pt.ptype(None)
        self._diameter = None
        self.tilt = []
        self._frozen = False
        self._dev_cache = None
        self.__freeze_attrs__ = ['amplitude', 'opd']
        return self

    def __repr__(self):
        return f'{self.__class__.__name__}()'

    def __amp__(self):
        return self._amplitude

    def __mask__(self):
        if self._mask is None:
            return np.copy(self.amplitude).astype(bool)
        return self._mask

    def __opd__(self):
        return self._opd

    def __mul__(self, wavefront):
        raise NotImplementedError

    def __rmul__(self, other):
        return self.__mul__(other)

    @property
    def ptype(self):
        return self._ptype

    @ptype.setter
    def ptype(self, value):
        self._ptype = _pt.ptype(value)

    @property
    def amplitude(self):
        """Electric field amplitude transmission (cached while frozen)."""
        return self._amplitude if self.frozen else self.__amp__()

    @amplitude.setter
    def amplitude(self, value):
        if self.frozen:
            raise RuntimeError('Can\'t set amplitude while Plane is frozen')
        self._amplitude = np.asarray(value)

    @property
    def opd(self):
        """Optical path difference (cached while frozen)."""
        return self._opd if self.frozen else self.__opd__()

    @opd.setter
    def opd(self, value):
        if self.frozen:
            raise RuntimeError('Can\'t set opd while Plane is frozen')
        self._opd = np.asarray(value)

    @property
    def mask(self):
        return self.__mask__()

    @property
    def global_mask(self):
        return self.mask if self.size < 2 else np.sum(self.mask, axis=0)

    @property
    def pixelscale(self):
        return self._pixelscale

    @property
    def diameter(self):
        if self._diameter is None:
            rmin, rmax, cmin, cmax = helper.boundary(self.global_mask)
            return np.max(np.max((rmax - rmin, cmax - cmin)) * np.asarray(self.pixelscale))
        return self._diameter

    @property
    def shape(self):
        m = self.mask
        return m.shape if self.size == 1 else (m.shape[1], m.shape[2])

    @property
    def size(self):
        m = self.mask
        return 1 if m.ndim in (0, 1, 2) else m.shape[0]

    @property
    def frozen(self):
        return self._frozen

    def freeze(self, inplace=True):
        """Cache the attributes named in ``__freeze_attrs__`` (lentil/plane.py:218-257).  On the
        device this is also what keeps amplitude / OPD / mask resident in HBM across a
        wavelength loop: a frozen plane uploads them once."""
        if self.frozen:
            raise RuntimeError('Plane is already frozen. Call thaw() to unfreeze.')
        target = self if inplace else copy.deepcopy(self)
        for attr in target.__freeze_attrs__:
            setattr(target, f'_{attr}', getattr(target, attr))
        target._frozen = True
        target._dev_cache = None
        if not inplace:
            return target

    def thaw(self):
        self._frozen = False
        self._dev_cache = None

    def rescale(self, scale):
        """Rescale a plane by interpolation (lentil/plane.py:271-330): amplitude by 3rd-order splines divided by `scale`
        (total power preserved), OPD by 3rd-order splines, mask by 0-order interpolation forced back to 0/1, pixelscale
        divided by `scale`.  The interpolation is the device `rescale` (rescale.cu: B-spline prefilter + separable
        evaluation, the scipy map_coordinates arithmetic of lentil/util.py:261-340)."""
        from .detector import rescale as _rescale
        plane = copy.deepcopy(self)
        frozen = plane._frozen
        plane._frozen = False
        if np.ndim(plane.amplitude) > 1:
            plane.amplitude = _rescale(np.asarray(plane.amplitude), scale=scale, shape=None, mask=None, order=3, mode='nearest',
                                       unitary=False) / scale
        if np.ndim(plane.opd) > 1:
            plane.opd = _rescale(np.asarray(plane.opd), scale=scale, shape=None, mask=None, order=3, mode='nearest', unitary=False)
        if plane._mask is not None:
            m = np.asarray(plane._mask)
            if m.ndim == 2:
                m = _rescale(m, scale=scale, shape=None, mask=None, order=0, mode='constant', unitary=False)
            else:
                m = np.asarray([_rescale(k, scale=scale, shape=None, mask=None, order=0, mode='constant', unitary=False) for k in m])
            m[np.nonzero(m)] = 1
            plane._mask = m.astype(int)
        if plane.pixelscale is not None:
            ps = np.broadcast_to(plane.pixelscale, (2,))
            plane._pixelscale = (ps[0] / scale, ps[1] / scale)
        plane._frozen = frozen
        plane._dev_cache = None
        return plane

    def resample(self, pixelscale):
        """Resample a plane to another pixelscale (lentil/plane.py:332-367)."""
        if self.pixelscale is None or not np.all(np.asarray(self.pixelscale)):
            raise ValueError("can't resample Plane with pixelscale = ()")
        ps = np.broadcast_to(self.pixelscale, (2,))
        if ps[0] != ps[1]:
            raise NotImplementedError("Can't resample non-uniformly sampled Plane")
        return self.rescale(scale=ps[0] / pixelscale)

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            setattr(new, k, None if k == '_dev_cache' else copy.deepcopy(v, memo))
        return new


def _mul_pixelscale(a, b):
    """Pixelscale of a wavefront-plane product (lentil/plane.py:614-629)."""
    a = None if a is None else np.broadcast_to(a, (2,))
    b = None if b is None else np.broadcast_to(b, (2,))
    if a is None:
        return b
    if b is None:
        return a
    if all(a == b):
        return a
    raise ValueError(f"can't multiply with inconsistent pixelscales: {a} != {b}")


# wavefront ptype -> {plane ptype: result ptype}; a missing plane ptype forbids the product
# (lentil/plane.py:634-668)
_MUL_PTYPE = {
    _pt.none: {_pt.none: _pt.none, _pt.pupil: _pt.pupil, _pt.image: _pt.image,
               _pt.tilt: _pt.none, _pt.transform: _pt.none},
    _pt.pupil: {_pt.pupil: _pt.pupil, _pt.tilt: _pt.pupil, _pt.transform: _pt.pupil},
    _pt.image: {_pt.image: _pt.image, _pt.tilt: _pt.pupil, _pt.transform: _pt.pupil},
}


def _plane_slice(mask):
    """Bounding-box slices of each segment of `mask` (lentil/plane.py:671-705)."""
    if mask is None or mask.ndim < 2:
        return [Ellipsis]
    if mask.ndim == 2:
        return [helper.boundary_slice(mask)]
    if mask.ndim == 3:
        return [helper.boundary_slice(m) for m in mask]
    raise ValueError('mask has invalid dimensions')


class Plane(_PlaneBase):
    """Finite geometric plane (lentil/plane.py:414-475).  Same constructor as the reference,
    including the ``amp`` alias."""

    def __init__(self, amplitude=None, opd=None, mask=None, pixelscale=None, diameter=None,
                 ptype=None, **kwargs):
        if 'amp' in kwargs:
            if amplitude is not None:
                raise AttributeError("Got both 'amplitude' and 'amp', which are aliases of one another")
            amplitude = kwargs['amp']
        if amplitude is not None:
            self._amplitude = np.asarray(amplitude)
        if opd is not None:
            self._opd = np.asarray(opd)
        if mask is not None:
            self._mask = np.asarray(mask)
        if pixelscale is not None:
            self._pixelscale = np.broadcast_to(pixelscale, (2,))
        if diameter is not None:
            self._diameter = diameter
        if ptype is not None:
            self._ptype = _pt.ptype(ptype)

    # ---- device-side operands -------------------------------------------------------------------
    def _operands(self):
        """Host metadata + device copies of (amplitude, opd, mask) for K1.  Uploaded once per
        multiply, or once per freeze() when the plane is frozen.  The per-segment bounding boxes
        (helper.boundary_slice in the reference, recomputed there on every multiply) come from a
        device reduction over the uploaded mask (lfd_mask_bbox)."""
        if self.frozen and self._dev_cache is not None:
            return self._dev_cache
        import torch
        amp, opd = np.asarray(self.amplitude), np.asarray(self.opd)
        # When the mask is the default one derived from the amplitude (plane.py:43-47:
        # amplitude.astype(bool)) then amp*mask == amp, so K1 needs no mask plane and only the
        # bounding box has to be found (same pixels: amp != 0).
        derived = self._mask is None and type(self).__mask__ is _PlaneBase.__mask__
        if amp.size == 1 and opd.size == 1:
            mask = (amp != 0) if derived else np.asarray(self.mask)
            nseg = 1 if mask.ndim < 3 else mask.shape[0]
            shape = tuple(mask.shape) if nseg == 1 else tuple(mask.shape[1:])
            ops = {'nseg': nseg, 'shape': shape, 'pshape': shape, 'slices': _plane_slice(mask), 'scalar': (amp, opd)}
            if self.frozen:
                self._dev_cache = ops
            return ops

        mask = None if derived else np.asarray(self.mask)
        full = False
        if derived:
            shape, nseg = tuple(amp.shape), 1
            if len(shape) != 2:
                # scalar amplitude with an array opd (a phase-only plane): the derived mask is a scalar, the
                # plane has no shape of its own and the phasor covers the whole opd array (plane.py:496-507)
                shape, full = tuple(opd.shape), True
                if len(shape) != 2:
                    raise ValueError('opd must be a scalar or a 2-D array')
        else:
            nseg = 1 if mask.ndim < 3 else mask.shape[0]
            shape = tuple(mask.shape) if nseg == 1 else tuple(mask.shape[1:])
            if len(shape) != 2:
                raise ValueError('array amplitude/opd need a 2-D (or 3-D segment) mask')
        ops = {'nseg': nseg, 'shape': shape, 'pshape': () if full else shape, 'scalar': None}
        use_mask = amp.size != 1 and not derived   # plane.py:503 — a scalar amplitude is not masked
        mk_dev, mask_binary = None, True
        if mask is not None:
            mk = mask.reshape((nseg,) + shape)
            if mk.dtype == bool:
                mk8 = mk.view(np.uint8)
            else:
                binary = mask_binary = bool(np.all((mk == 0) | (mk == 1)))
                if not binary and use_mask:
                    if nseg > 1:
                        raise NotImplementedError('non-binary segment masks are not supported')
                    amp, use_mask = amp * mk[0], False
                mk8 = (mk > 0).astype(np.uint8)
            mk_dev = device.to_dev(mk8)
        ops['amp'] = device.to_dev(amp if amp.shape == shape else np.broadcast_to(amp, shape), dtype=np.float64)
        ops['mask'] = mk_dev if use_mask else None
        # the uploaded 0/1 mask cube itself, whether or not the phasor needs it (a scalar amplitude is not masked, but
        # fit_tilt still fits over the mask: lentil/plane.py:522-562)
        ops['mask_dev'], ops['mask_binary'] = mk_dev, mask_binary
        # bounding boxes on the device
        if full:
            bb = np.array([[0, shape[0] - 1, 0, shape[1] - 1]])
        else:
            bb = torch.empty(4 * nseg, dtype=torch.int32, device=device.device())
            src, is_f64, nonzero = (ops['amp'], 1, 1) if mk_dev is None else (mk_dev, 0, 0)
            _lib.check(_lib.lib().lfd_mask_bbox(src.data_ptr(), is_f64, nonzero, shape[0], shape[1], nseg,
                                                bb.data_ptr(), device.stream_ptr()), "lfd_mask_bbox")
            # the OPD upload is queued behind the bounding-box read-back, so that it overlaps the host-side planning
            # that follows instead of delaying the (synchronous) read-back
            bb_host = bb.cpu()
            ops['opd'] = device.to_dev(opd if opd.shape == shape else np.broadcast_to(opd, shape), dtype=np.float64)
            bb = bb_host.numpy().reshape(nseg, 4)
        if 'opd' not in ops:
            ops['opd'] = device.to_dev(opd if opd.shape == shape else np.broadcast_to(opd, shape), dtype=np.float64)
        if np.any(bb[:, 1] < 0):
            raise IndexError('mask plane without any data: cannot find its boundary')   # as np.where(...)[0][[0,-1]]
        slices = [np.s_[int(b[0]):int(b[1]) + 1, int(b[2]):int(b[3]) + 1] for b in bb]
        ops['slices'] = slices
        segs = (_lib.Segment * nseg)()
        total = 0
        for k, b in enumerate(bb):
            r0, r1, c0, c1 = int(b[0]), int(b[1]) + 1, int(b[2]), int(b[3]) + 1
            segs[k].r0, segs[k].c0, segs[k].h, segs[k].w = r0, c0, r1 - r0, c1 - c0
            segs[k].mask_index = k if nseg > 1 else 0
            segs[k].out_offset = total
            total += (r1 - r0) * (c1 - c0)
        ops['segs'], ops['total'] = segs, total
        ops['offsets'] = [helper.slice_offset(s, shape) for s in slices]
        if self.frozen:
            self._dev_cache = ops
        return ops

    def _phasors(self, wavelengths, ops=None):
        """K1: phasor tiles for a list of wavelengths.  Returns a (nlam, total) complex128 device
        buffer; segment k of wavelength l is buf[l, off_k : off_k + h_k*w_k] viewed (h_k, w_k)."""
        ops = self._operands() if ops is None else ops
        lam = np.ascontiguousarray(wavelengths, dtype=np.float64).reshape(-1)
        buf = device.empty_c128(len(lam), ops['total'])
        rc = _lib.lib().lfd_pupil_prep(
            ops['amp'].data_ptr(), ops['opd'].data_ptr(),
            ops['mask'].data_ptr() if ops['mask'] is not None else None,
            ops['shape'][0], ops['shape'][1], ops['segs'], ops['nseg'],
            lam.ctypes.data_as(C.POINTER(C.c_double)), len(lam),
            buf.data_ptr(), ops['total'], device.stream_ptr())
        _lib.check(rc, "lfd_pupil_prep")
        return ops, buf

    def _phasors_into(self, buf, wavelengths, ops, opd_dev=None):
        """K1 into rows of an existing (nlam, total) buffer, optionally with another OPD map
        (a Monte-Carlo realisation) in place of the plane's own."""
        import torch
        lam = np.ascontiguousarray(wavelengths, dtype=np.float64).reshape(-1)
        opd = ops['opd'] if opd_dev is None else opd_dev
        fn = _lib.lib().lfd_pupil_prep_c64 if buf.dtype == torch.complex64 else _lib.lib().lfd_pupil_prep
        rc = fn(
            ops['amp'].data_ptr(), opd.data_ptr(),
            ops['mask'].data_ptr() if ops['mask'] is not None else None,
            ops['shape'][0], ops['shape'][1], ops['segs'], ops['nseg'],
            lam.ctypes.data_as(C.POINTER(C.c_double)), len(lam),
            buf.data_ptr(), ops['total'], device.stream_ptr())
        _lib.check(rc, "lfd_pupil_prep")
        return buf

    @staticmethod
    def _segment_view(ops, buf_row, k):
        sg = ops['segs'][k]
        return buf_row[sg.out_offset: sg.out_offset + sg.h * sg.w].view(sg.h, sg.w)

    def __mul__(self, wavefront):
        """Multiply a wavefront by this plane (lentil/plane.py:477-516): one output Field per
        (incoming field, segment), phasor = amp[s]*mask_n[s] * exp(2 pi i opd[s] / lambda) at
        offset slice_offset(s), tilt = [self.tilt[n]] when fit_tilt ran."""
        from .wavefront import Wavefront
        wf_pt, my_pt = wavefront.ptype, self.ptype
        if my_pt not in _MUL_PTYPE[wf_pt]:
            raise TypeError(f"can't multiply Wavefront with ptype '{wf_pt}' by Plane with ptype '{my_pt}'")
        ops = self._operands()
        pixelscale = _mul_pixelscale(self.pixelscale, wavefront.pixelscale)
        shape = wavefront.shape if ops['pshape'] == () else ops['pshape']
        out = Wavefront.empty(wavelength=wavefront.wavelength, pixelscale=pixelscale,
                              focal_length=wavefront.focal_length, shape=shape,
                              ptype=_MUL_PTYPE[wf_pt][my_pt])
        if ops['scalar'] is not None:
            amp, opd = ops['scalar']
            value = amp * np.exp(2 * np.pi * 1j * opd / wavefront.wavelength)
            phasors = [Field(value, pixelscale=self.pixelscale,
                             offset=helper.slice_offset(s, ops['shape']) if ops['shape'] != () else (0, 0),
                             tilt=[self.tilt[n]] if self.tilt else [])
                       for n, s in enumerate(ops['slices'])]
        else:
            _, buf = self._phasors([wavefront.wavelength], ops)
            phasors = [Field(self._segment_view(ops, buf[0], n), pixelscale=self.pixelscale,
                             offset=ops['offsets'][n], tilt=[self.tilt[n]] if self.tilt else [])
                       for n in range(ops['nseg'])]
        for field in wavefront.data:
            for phasor in phasors:
                res = field * phasor
                if res.size > 0:
                    out.data.append(res)
        return out

    @property
    def _slice(self):
        return _plane_slice(self.mask)

    # ---- tilt fitting (setup-time host code) -----------------------
    @property
    def ptt_vector(self):
        """Piston / x-tilt / y-tilt basis per segment, shape (3*size, npix)
        (lentil/plane.py:522-562); None for planes without a mask."""
        if self.shape == () or self.shape is None:
            return None
        if self.pixelscale is None:
            raise ValueError("can't create ptt_vector with pixelscale = ()")
        ps = np.broadcast_to(self.pixelscale, (2,))
        r, c = helper.mesh(self.shape)
        base = np.stack([np.ones(r.size), r.ravel() * ps[0], -c.ravel() * ps[1]])
        if self.size == 1:
            return base * self.mask.ravel()
        return np.concatenate([base * m.ravel() for m in self.mask], axis=0)

    def fit_tilt(self, inplace=False):
        """Least-squares fit and removal of per-segment tilt from the OPD; the equivalent angles
        are kept as Tilt objects in ``self.tilt`` (lentil/plane.py:564-611).

        Runs on the device (lfd_fit_tilt_moments + lfd_remove_tilt): the reference's dense
        (npix x 3) design matrix per segment has zero rows outside the mask, so its lstsq solution
        is that of the 3x3 normal equations over the masked pixels; the nine moments per segment
        are reduced on the GPU, the 3x3 systems solved here."""
        plane = self if inplace else copy.deepcopy(self)
        if plane.shape == () or plane.shape is None or plane.opd.size == 1:
            return plane
        if plane.pixelscale is None:
            raise ValueError("can't create ptt_vector with pixelscale = ()")
        ps = np.broadcast_to(plane.pixelscale, (2,))
        ops = plane._operands()
        if ops['scalar'] is not None:
            return plane
        if not ops['mask_binary']:
            raise NotImplementedError('fit_tilt with a non-binary mask is not supported')
        nseg, (n_r, n_c) = ops['nseg'], ops['shape']
        L = _lib.lib()
        # an explicit mask is the fit's support even when the phasor ignores it (scalar amplitude); a derived mask is
        # amplitude != 0, which the kernels evaluate themselves
        mask_ptr = ops['mask_dev'].data_ptr() if ops['mask_dev'] is not None else None
        moments = device.zeros_f64(nseg * 9)
        scratch = device.empty_bytes(64 * nseg)
        _lib.check(L.lfd_fit_tilt_moments(ops['opd'].data_ptr(), mask_ptr, ops['amp'].data_ptr(), n_r, n_c,
                                          float(ps[0]), float(ps[1]), ops['segs'], nseg, moments.data_ptr(),
                                          scratch.data_ptr(), scratch.numel(), device.stream_ptr()),
                   "lfd_fit_tilt_moments")
        m = device.to_host(moments).reshape(nseg, 9)
        coef = np.zeros((nseg, 3))
        for k in range(nseg):
            s1, sx, sy, sxx, sxy, syy, sz, sxz, syz = m[k]
            nmat = np.array([[s1, sx, sy], [sx, sxx, sxy], [sy, sxy, syy]])
            sol = np.linalg.solve(nmat, np.array([sz, sxz, syz]))
            sg = ops['segs'][k]
            xc = (sg.r0 + 0.5 * (sg.h - 1) - n_r // 2) * ps[0]
            yc = -(sg.c0 + 0.5 * (sg.w - 1) - n_c // 2) * ps[1]
            coef[k] = (sol[0] - sol[1] * xc - sol[2] * yc, sol[1], sol[2])
        coef_dev = device.to_dev(coef, dtype=np.float64)
        out = device.zeros_f64(n_r, n_c)
        _lib.check(L.lfd_remove_tilt(ops['opd'].data_ptr(), mask_ptr, ops['amp'].data_ptr(), n_r, n_c, nseg,
                                     float(ps[0]), float(ps[1]), coef_dev.data_ptr(), out.data_ptr(),
                                     device.stream_ptr()), "lfd_remove_tilt")
        frozen = plane._frozen
        plane._frozen = False
        plane.opd = device.to_host(out)
        plane._frozen = frozen
        plane.tilt.extend(Tilt(x=c[1], y=c[2]) for c in coef)
        plane._dev_cache = None
        return plane


class Pupil(Plane):
    """Pupil plane (lentil/plane.py:708-771): multiplying hands its focal length to the wavefront."""

    def __new__(cls, *args, **kwargs):
        self = super().__new__(cls, *args, **kwargs)
        self._focal_length = None
        self._ptype = _pt.pupil
        return self

    def __init__(self, amplitude=None, opd=None, mask=None, pixelscale=None, focal_length=None,
                 diameter=None, **kwargs):
        super().__init__(amplitude=amplitude, opd=opd, mask=mask, pixelscale=pixelscale,
                         diameter=diameter, ptype=_pt.pupil, **kwargs)
        if focal_length is not None:
            self._focal_length = focal_length

    def __mul__(self, wavefront):
        wavefront = super().__mul__(wavefront)
        wavefront.focal_length = self.focal_length
        return wavefront

    @property
    def focal_length(self):
        return self._focal_length


class Image(Plane):
    """Image plane (lentil/plane.py:774-820)."""

    def __new__(cls, *args, **kwargs):
        self = super().__new__(cls, *args, **kwargs)
        self._ptype = _pt.image
        return self

    def __init__(self, amplitude=None, opd=None, mask=None, pixelscale=None, **kwargs):
        super().__init__(amplitude=amplitude, opd=opd, mask=mask, pixelscale=pixelscale,
                         ptype=_pt.image, **kwargs)

    def __mul__(self, wavefront):
        wavefront = super().__mul__(wavefront)
        wavefront.ptype = _pt.image
        return wavefront

    def fit_tilt(self, *args, **kwargs):
        return self


class _TiltBase(Plane):
    """Planes that act through the tilt interface (lentil/plane.py:823-881): the product
    appends the plane to every field's tilt list; ``__shift__`` converts it to a shift later."""

    def __new__(cls, *args, **kwargs):
        self = super().__new__(cls, *args, **kwargs)
        self._ptype = kwargs['ptype'] if 'ptype' in kwargs else _pt.tilt
        return self

    def __mul__(self, wavefront):
        wavefront = super().__mul__(wavefront)
        for field in wavefront.data:
            field.tilt.append(self)
        return wavefront

    def __shift__(self, wavelength, x0, y0, **kwargs):
        raise NotImplementedError


class Tilt(_TiltBase):
    """Angular tilt (lentil/plane.py:884-923).  `x` is radians about the x-axis, which moves the
    image along y — hence the swap in the constructor (:898-901)."""

    def __init__(self, x, y, **kwargs):
        super().__init__(**kwargs)
        self.x = y
        self.y = x

    def __shift__(self, xs=0, ys=0, z=0, **kwargs):
        return xs - (z * self.x), ys - (z * self.y)


_GL_NODES, _GL_WEIGHTS = np.polynomial.legendre.leggauss(48)


class DispersiveTilt(_TiltBase):
    """Spectral dispersion acting as a wavelength-dependent tilt (lentil/plane.py:926-1096).

    ``dispersion`` maps distance along the spectral trace to wavelength, ``trace`` maps focal-plane
    x to y (both polynomial coefficients, highest power first, metres).  ``__shift__`` returns the
    focal-plane position of `wavelength` along the trace, anchored at the undispersed source.

    First-order polynomials are solved in closed form exactly as the reference does
    (:1030-1033, :1043-1045).  Higher orders, which the reference hands to scipy's ``leastsq`` /
    ``quad`` one scalar at a time (:1037, :1050), are solved here by a safeguarded Newton iteration
    on the polynomial and on a Gauss-Legendre arc length — vectorised, so a whole wavelength grid
    costs one call (the batch driver evaluates every plane's shift up front).
    """

    def __init__(self, trace=None, dispersion=None, **kwargs):
        super().__init__(**kwargs)
        if trace is not None:
            self.trace = np.asarray(trace)
        if dispersion is not None:
            self.dispersion = np.asarray(dispersion)

    def __shift__(self, wavelength, xs=0., ys=0., **kwargs):
        x, y = self._pos(self._dist(wavelength))
        return x + xs, y + ys

    # ---- wavelength -> distance along the trace ---------------------------------------------------
    def _dist(self, wavelength):
        disp = np.asarray(self.dispersion, dtype=float)
        if len(disp) == 2:
            return (wavelength - disp[1]) / disp[0]
        wavelength = np.asarray(wavelength, dtype=float)
        dpoly = np.polyder(disp)
        d = np.zeros_like(wavelength)
        done = np.zeros(wavelength.shape, dtype=bool)
        with np.errstate(all='ignore'):
            for _ in range(100):
                step = (np.polyval(disp, d) - wavelength) / np.polyval(dpoly, d)
                d = np.where(done, d, d - step)
                done |= np.abs(step) <= 1e-15 * np.maximum(np.abs(d), 1e-300)
                if np.all(done):
                    break
        if not np.all(done):
            # no real root for these wavelengths: the reference's least-squares solve settles on
            # the stationary point of the residual, i.e. a root of the derivative (:1037)
            ddpoly = np.polyder(dpoly)
            e = np.zeros_like(wavelength)
            for _ in range(100):
                step = np.polyval(dpoly, e) / np.polyval(ddpoly, e)
                e = e - step
                if np.all(np.abs(step) <= 1e-15 * np.maximum(np.abs(e), 1e-300)):
                    break
            d = np.where(done, d, e)
        return d

    # ---- distance along the trace -> (x, y) -------------------------------------------------------
    def _arc_len(self, x):
        """Arc length of the trace from 0 to x (Gauss-Legendre on sqrt(1 + y'(t)^2))."""
        slope = np.polyder(np.asarray(self.trace, dtype=float))
        x = np.asarray(x, dtype=float)
        t = 0.5 * x[..., None] * (_GL_NODES + 1.0)
        return 0.5 * x * np.sum(_GL_WEIGHTS * np.sqrt(1.0 + np.polyval(slope, t) ** 2), axis=-1)

    def _pos(self, dist):
        trace = np.asarray(self.trace, dtype=float)
        if len(trace) == 2:
            x = dist / np.sqrt(1 + trace[0] ** 2)
        else:
            slope = np.polyder(trace)
            dist = np.asarray(dist, dtype=float)
            x = dist / np.sqrt(1.0 + np.polyval(slope, 0.0) ** 2)
            for _ in range(100):
                step = (self._arc_len(x) - dist) / np.sqrt(1.0 + np.polyval(slope, x) ** 2)
                x = x - step
                if np.all(np.abs(step) <= 1e-15 * np.maximum(np.abs(x), 1e-300)):
                    break
        return x, np.polyval(trace, x)


class Grism(DispersiveTilt):
    """Deprecated alias of :class:`DispersiveTilt` (lentil/plane.py:1099-1167 warns the same way)."""

    def __init__(self, trace, dispersion, **kwargs):
        import warnings
        warnings.warn("Grism is deprecated; DispersiveTilt replaces it.", DeprecationWarning,
                      stacklevel=2)
        super().__init__(trace=trace, dispersion=dispersion, **kwargs)
