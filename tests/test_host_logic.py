"""Host-side mirror of the reference interface (integer geometry, metadata, error behaviour)
and the C-ABI export table.  CPU only: nothing here launches a kernel."""
import ctypes
import os
import re

import numpy as np
import pytest

import lentil_b200 as lentil
import lentil_oracle as oc
from lentil_b200 import _lib, extent, helper
from lentil_b200.propagate import plan_window, _mask_shape, _mask_shift
from conftest import ROOT


# ---- C ABI ------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "lentil_b200.h")).read()
    declared = set(re.findall(r"\b(lfd_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in include/lentil_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_struct_layouts_match():
    L = _lib.lib()
    assert L.lfd_abi_version() == 2
    assert L.lfd_struct_size(0) == ctypes.sizeof(_lib.MftDesc)
    assert L.lfd_struct_size(1) == ctypes.sizeof(_lib.Segment)
    assert L.lfd_struct_size(2) == ctypes.sizeof(_lib.Window)


def test_argument_errors_do_not_need_a_gpu():
    L = _lib.lib()
    d = (_lib.MftDesc * 1)()
    assert L.lfd_mft_c128_batched(d, 1, None, 0, None) != 0
    assert b"workspace" in L.lfd_last_error()
    assert L.lfd_mft_c128_batched(d, 0, None, 0, None) == 0      # empty batch is a no-op


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        lentil.fourier.dft2(np.ones((4, 4)), 0.25)


# ---- extent / helper: bit-exact integer logic -----------------------------------------------------
def test_extent_matches_golden(golden):
    for row in golden("extent")["rows"]:
        sa, sb, ha, hb = tuple(row[0:2]), tuple(row[2:4]), tuple(row[4:6]), tuple(row[6:8])
        ea, eb = tuple(row[8:12]), tuple(row[12:16])
        assert extent.array_extent(sa, ha) == ea and extent.array_extent(sb, hb) == eb
        assert extent.intersect(ea, eb) == bool(row[16])
        ishape = extent.intersection_shape(ea, eb)
        assert (tuple(ishape) if ishape else (0, 0)) == tuple(row[17:19])
        assert extent.intersection_shift(ea, eb) == tuple(row[19:21])
        assert extent.intersection_slices(ea, eb) == oc.intersection_slices(ea, eb)
        assert extent.array_center(ea) == oc.array_center(ea)


def test_extent_float_shift_truncates_toward_zero():
    # lentil/extent.py:28-29 uses int(); callers pass np.fix'd shifts
    for s in (-2.7, -0.2, 0.9, 3.999):
        assert extent.array_extent((5, 4), (s, -s)) == oc.array_extent((5, 4), (s, -s))
    assert extent.array_extent((), (3, -2)) == oc.array_extent((), (3, -2)) == (3, 3, -2, -2)


def test_helper_matches_golden(golden):
    d = golden("helper")
    for i in range(int(d["nh"])):
        a = d[f"h{i}_a"]
        slc = helper.boundary_slice(a)
        assert [slc[0].start, slc[0].stop, slc[1].start, slc[1].stop] == list(d[f"h{i}_slice"])
        assert tuple(helper.slice_offset(slc, a.shape)) == tuple(d[f"h{i}_offset"])
    assert helper.slice_offset(Ellipsis, (4, 4)) == (0, 0)


def test_boundary_slice_random():
    # reference tests/test_helper.py:14-26
    rng = np.random.default_rng(5)
    for _ in range(20):
        a = np.zeros((10, 10))
        r, c = rng.integers(0, 10, 2)
        a[r:r + 3, c:c + 3] = 1
        assert helper.boundary_slice(a) == (slice(r, min(r + 3, 10)), slice(c, min(c + 3, 10)))


def test_plan_window_matches_oracle_on_random_shifts():
    rng = np.random.default_rng(11)
    n_hit = 0
    for _ in range(500):
        shape_out = tuple(int(v) for v in rng.integers(8, 200, 2))
        prop = tuple(int(v) for v in np.minimum(rng.integers(4, 200, 2), shape_out))
        shift = rng.uniform(-150, 150, 2)
        if rng.random() < 0.3:
            mshape = tuple(int(v) for v in rng.integers(1, min(shape_out) + 1, 2))
            mshift = tuple(int(v) for v in rng.integers(-20, 20, 2))
            oe = extent.array_extent(mshape, mshift)
        else:
            oe = extent.array_extent(shape_out, (0, 0))
        a, b = plan_window(shift, prop, oe), oc.plan_window(shift, prop, oe)
        if b is None:
            assert a is None
            continue
        n_hit += 1
        assert tuple(a[0]) == tuple(b[0]) and tuple(a[1]) == tuple(b[1])
        assert np.array_equal(a[2], b[2])
    assert n_hit > 100


def test_plan_window_worked_example():
    # SURVEY.md appendix A.3
    out_extent = extent.array_extent((512, 512), (0, 0))
    ishape, ishift, dshift = plan_window((26.4, 13.6), (512, 512), out_extent)
    assert tuple(ishape) == (486, 499) and tuple(ishift) == (13, 6)
    assert np.allclose(dshift, (13.4, 7.6))


def test_mask_window_matches_oracle():
    rng = np.random.default_rng(2)
    for _ in range(20):
        m = np.zeros((40, 50))
        r, c = rng.integers(0, 30, 2)
        h, w = rng.integers(1, 10, 2)
        m[r:r + h, c:c + w] = rng.random((min(h, 40 - r), min(w, 50 - c))) + 0.1
        shape, shift = oc.mask_window(m)
        assert tuple(_mask_shape(m)) == tuple(shape) and tuple(_mask_shift(m)) == tuple(shift)


# ---- metadata objects ----------------------------------------------------------------------------
def test_ptype():
    assert lentil.ptype(None) == lentil.none and lentil.ptype('pupil') == lentil.pupil
    assert hash(lentil.ptype('image')) == hash(lentil.image) and str(lentil.tilt) == 'tilt'
    with pytest.raises(TypeError):
        lentil.ptype('bogus')


def test_default_wavefront():
    # reference tests/test_wavefront.py:6-14
    w = lentil.Wavefront(wavelength=500e-9)
    assert np.array_equal(w.field, 1 + 0j)
    assert np.array_equal(w.intensity, 1)
    assert lentil.Plane() * w


def test_propagate_requires_pupil_or_image():
    # reference tests/test_wavefront.py:16-19
    w = lentil.Wavefront(wavelength=500e-9)
    with pytest.raises(TypeError):
        lentil.propagate_dft(w, shape=(64, 64), pixelscale=5e-6)


def test_wavefront_ptype_guard_and_tilt_arg():
    with pytest.raises(TypeError):
        lentil.Wavefront(500e-9, ptype='tilt')
    with pytest.raises(ValueError):
        lentil.Wavefront(500e-9, tilt=[1, 2, 3])
    w = lentil.Wavefront(500e-9, tilt=[3e-6, -2e-6])
    t = w.data[0].tilt[0]
    assert (t.x, t.y) == (-2e-6, 3e-6)          # constructor swap, lentil/plane.py:898-901
    assert t.__shift__(xs=1.0, ys=2.0, z=10.0) == (1.0 + 2e-5, 2.0 - 3e-5)


def test_field_shift_matches_oracle():
    rng = np.random.default_rng(4)
    for _ in range(20):
        tl = rng.normal(size=(3, 2)) * 1e-5
        f = lentil.Field(1, tilt=[lentil.Tilt(x=a, y=b) for a, b in tl])
        g = oc.make_field(1, None, [oc.tilt_entry(a, b) for a, b in tl])
        a = f.shift(z=17.0, wavelength=6e-7, pixelscale=(5e-6, 4e-6), oversample=3)
        assert a == oc.field_shift(g, 17.0, (5e-6, 4e-6), 3)


@pytest.mark.parametrize('field_shape, field_shift, output_shape', [
    ((5, 5), (-25, 0), (10, 10)), ((5, 5), (25, 0), (10, 10)),
    ((5, 5), (0, -25), (10, 10)), ((5, 5), (0, 25), (10, 10))])
def test_overlap(field_shape, field_shift, output_shape):
    # reference tests/test_wavefront.py:22-29
    assert lentil.wavefront._overlap(field_shape, field_shift, output_shape) is False


def test_scalar_field_products():
    # the scalar rows of reference tests/test_field.py:7-27 (array rows run on the GPU)
    F = lentil.Field
    c = F(1, pixelscale=1, offset=[0, 0]) * F(1, pixelscale=1, offset=[0, 0])
    assert np.array_equal(c.data, 1 + 0j) and np.array_equal(c.offset, [0, 0])
    assert (F(1, pixelscale=1, offset=[0, 0]) * F(1, pixelscale=1, offset=[1, 0])).size == 0
    assert (F(1, pixelscale=1, offset=[0, 0]) * F(1, pixelscale=1, offset=[-10, 0])).size == 0


def test_mul_ptype_rules_and_pixelscale():
    w = lentil.Wavefront(500e-9, ptype='image')
    with pytest.raises(TypeError):
        lentil.Pupil() * w                       # image wavefront x pupil plane is not allowed
    with pytest.raises(ValueError):
        lentil.Plane(pixelscale=1.0) * lentil.Wavefront(500e-9, pixelscale=2.0)
    out = lentil.Tilt(1e-6, 2e-6) * lentil.Wavefront(500e-9)
    assert len(out.data) == 1 and len(out.data[0].tilt) == 1


def test_freeze_semantics():
    p = lentil.Plane(amplitude=np.ones((4, 4)), opd=np.zeros((4, 4)))
    p.freeze()
    with pytest.raises(RuntimeError):
        p.opd = np.ones((4, 4))
    with pytest.raises(RuntimeError):
        p.freeze()
    p.thaw()
    p.opd = np.ones((4, 4))
    assert p.size == 1 and p.shape == (4, 4)


# ---- DispersiveTilt (lentil/plane.py:926-1167) --------------------------------------------------------
def test_dispersive_tilt_center_and_shift():
    # reference tests/test_plane.py:163-181
    dt = lentil.DispersiveTilt(dispersion=[1, 650e-9], trace=[2, 0])
    x, y = dt.__shift__(wavelength=650e-9, xs=1.25, ys=-3.5)
    assert x == 1.25 and y == -3.5
    dt = lentil.DispersiveTilt(trace=[1, 1], dispersion=[1, 650e-9])
    x, y = dt.__shift__(wavelength=900e-9)
    assert x == (900e-9 - dt.dispersion[1]) / np.sqrt(2) and y == 1 + x


def test_dispersive_tilt_class_attribute_overload():
    # reference tests/test_plane.py:213-219
    class Disperser(lentil.DispersiveTilt):
        dispersion = [1, 0]
        trace = [1, 0]
    assert np.allclose(Disperser().__shift__(1), np.sqrt(2) / 2)
    assert Disperser().ptype == lentil.tilt


def test_dispersive_tilt_matches_reference_vectors(golden):
    d = golden("dispersive_tilt")
    wl = d["wl"]
    for i in range(int(d["n"])):
        dt = lentil.DispersiveTilt(trace=d[f"c{i}_trace"], dispersion=d[f"c{i}_disp"])
        ref = d[f"c{i}_xy"]
        x, y = dt.__shift__(wavelength=wl, xs=1e-3, ys=-2e-3)           # whole grid in one call
        scale = np.max(np.abs(ref))
        assert np.max(np.abs(x - ref[:, 0])) <= 2e-7 * scale and np.max(np.abs(y - ref[:, 1])) <= 2e-7 * scale
        for k in (0, len(wl) - 1):                                       # scalar calls agree with the grid
            xs, ys = dt.__shift__(wavelength=float(wl[k]), xs=1e-3, ys=-2e-3)
            assert np.isclose(xs, x[k], rtol=1e-13, atol=0) and np.isclose(ys, y[k], rtol=1e-13, atol=0)


def test_grism_alias_warns():
    with pytest.warns(DeprecationWarning):
        g = lentil.Grism(trace=[1, 0], dispersion=[1, 0])
    assert np.allclose(g.__shift__(1), np.sqrt(2) / 2)


def test_dispersive_tilt_in_propagation_shift_chain():
    # a DispersiveTilt in a Field's tilt list moves the window by the dispersed position
    dt = lentil.DispersiveTilt(trace=[0.5, 0.0], dispersion=[2e-3, 650e-9])
    w = dt * lentil.Wavefront(700e-9)
    x, y = dt.__shift__(wavelength=700e-9)
    r, c = w.data[0].shift(z=10.0, wavelength=700e-9, pixelscale=(5e-6, 5e-6), oversample=2)
    assert np.isclose(r, -y / 5e-6 * 2) and np.isclose(c, x / 5e-6 * 2)


def test_fft_geometry_equals_padded_shifted_fft():
    # propagate_fft never builds the padded array: the dft2 offset/shift it hands to K2a must make
    # the matrix transform of the surviving rows equal ifftshift(fft2(fftshift(pad(x)))) for even
    # and odd padded sizes, pads and crops (lentil/propagate.py:83-84,135-136, util.py:31-90)
    from lentil_b200.propagate import _padded_fft_geometry
    rng = np.random.default_rng(0)
    for have, npix in [((20, 20), (32, 32)), ((21, 20), (33, 31)), ((20, 21), (31, 32)),
                       ((40, 41), (32, 31)), ((33, 33), (33, 33)), ((10, 11), (11, 10))]:
        x = rng.normal(size=have) + 1j * rng.normal(size=have)
        ref = np.fft.ifftshift(np.fft.fft2(np.fft.fftshift(oc.pad(x, npix)), norm='ortho'))
        r0, nr, off_r, odd_r = _padded_fft_geometry(have[0], npix[0])
        c0, nc, off_c, odd_c = _padded_fft_geometry(have[1], npix[1])
        F = oc.dft2(x[r0:r0 + nr, c0:c0 + nc], (1 / npix[0], 1 / npix[1]), shape=npix,
                    shift=(odd_r, odd_c), offset=(off_r, off_c))
        assert np.max(np.abs(F - ref)) <= 1e-13 * np.max(np.abs(ref))


def test_scratch_shape_and_fft_shape(golden):
    d = golden("propagate_fft")
    assert tuple(lentil.scratch_shape([500e-9, 700e-9], 1 / 36, 5e-6, 10.0, 2)) == tuple(d["scratch"])
    from lentil_b200.propagate import _fft_shape
    npix, pw = _fft_shape(np.array([1 / 36, 1 / 36]), np.array([5e-6, 5e-6]), 10.0, 650e-9, 2)
    onpix, opw = oc.fft_shape(np.array([1 / 36, 1 / 36]), np.array([5e-6, 5e-6]), 10.0, 650e-9, 2)
    assert tuple(npix) == tuple(onpix) and pw == opw


def test_propagate_fft_refuses_tilt_before_touching_the_device():
    w = lentil.Wavefront(650e-9, tilt=[1e-6, 0])
    with pytest.raises(NotImplementedError):
        lentil.propagate_fft(w, pixelscale=5e-6, shape=(8, 8))


def test_spline_tap_tables_match_the_oracle():
    # host side of rescale (lentil/util.py:329-343): per-axis tap indices and B-spline weights for every order and
    # extension mode, including coordinates outside the array
    from lentil_b200.detector import _spline_taps
    coords = np.array([-30.2, -13.5, -12.2, -3.2, -1.0, -0.4, 0, 0.49, 0.5, 1.0, 2.3, 7.5, 8.0, 8.3, 9.7, 13.1, 19.9, 50.0])
    for n in (1, 2, 9, 40):
        for order in range(6):
            for mode in ("nearest", "constant", "reflect", "mirror", "wrap"):
                npad, idx, w = _spline_taps(coords, n, order, mode)
                rpad, ridx, rw = oc.spline_taps(coords, n, order, mode)
                assert npad == rpad
                live = np.abs(rw) > 0                      # taps with zero weight may point anywhere
                assert np.array_equal(idx[live], ridx[live]), (n, order, mode)
                assert np.max(np.abs(w - rw)) <= (1e-15 if order <= 3 else 5e-14), (n, order, mode)   # orders 4, 5: sum formula with cancellation
                assert idx.min() >= 0 and idx.max() < n + 2 * npad


def test_rescale_argument_errors():
    with pytest.raises(RuntimeError):
        lentil.rescale(np.ones((4, 4)), 0.5, order=7)


def test_patch_rebinds_and_restores_the_reference_entry_points():
    # lentil resolves dft2 / propagate_dft through module attributes at call time
    import types
    from lentil_b200 import patch, fourier, propagate, detector
    ref = types.SimpleNamespace(
        fourier=types.SimpleNamespace(dft2="ref_dft2", idft2="ref_idft2"),
        propagate=types.SimpleNamespace(propagate_dft="ref_prop"), propagate_dft="ref_prop",
        util=types.SimpleNamespace(rebin="ref_rebin", rescale="ref_rescale"), rebin="ref_rebin", rescale="ref_rescale",
        detector=types.SimpleNamespace(pixel="ref_pixel"))              # no pixelate: must be skipped, not created
    patch.enable(ref)
    assert ref.fourier.dft2 is fourier.dft2 and ref.fourier.idft2 is fourier.idft2 and ref.propagate_dft == "ref_prop"
    patch.enable(ref, level='path')
    assert ref.propagate.propagate_dft is propagate.propagate_dft and ref.propagate_dft is propagate.propagate_dft
    assert ref.util.rescale is detector.rescale and ref.rebin is detector.rebin and ref.detector.pixel is detector.pixel
    assert not hasattr(ref.detector, "pixelate")
    patch.disable(ref)
    assert ref.fourier.dft2 == "ref_dft2" and ref.propagate.propagate_dft == "ref_prop" and ref.util.rescale == "ref_rescale"
    assert ref.detector.pixel == "ref_pixel" and ref.rebin == "ref_rebin"
    with pytest.raises(ValueError):
        patch.enable(ref, level='everything')


def test_k2a_execution_choice_and_workspace_are_host_decisions():
    # lfd_set_mft_variant / lfd_mft_execution / lfd_mft_workspace_bytes never touch the device: which execution of the
    # transform a batch gets (direct / folded DMMA / chirp-z) follows from the plane shapes alone
    L = _lib.lib()
    saved = L.lfd_get_mft_variant()

    def descs(*shapes):
        d = (_lib.MftDesc * len(shapes))()
        for k, (m, n, M, N) in enumerate(shapes):
            d[k].m, d[k].n, d[k].M, d[k].N = m, n, M, N
        return d, len(shapes)

    try:
        bench_plane = descs((1001, 1001, 1024, 1024))
        cfg5_plane = descs((4081, 4081, 2048, 2048))          # 6128 -> 8192-point transform: the longest that fits shared memory
        huge = descs((6000, 16, 4096, 16))                    # 10095 -> 16384 points: folded DMMA form
        mixed = descs((241, 241, 256, 256), (6000, 16, 4096, 16))
        tiny = descs((7, 5, 4, 9))
        assert L.lfd_get_mft_variant() == 3                    # LFD_MFT_AUTO is the default
        for variant, expect in ((0, (0, 0, 0, 0, 0)), (1, (1, 1, 1, 1, 1)), (2, (2, 2, 1, 1, 2)), (3, (2, 2, 1, 1, 2))):
            assert L.lfd_set_mft_variant(variant) == 0
            got = tuple(L.lfd_mft_execution(*b) for b in (bench_plane, cfg5_plane, huge, mixed, tiny))
            assert got == expect, (variant, got)
            for b in (bench_plane, cfg5_plane, huge, mixed, tiny):
                assert L.lfd_mft_workspace_bytes(*b) > 0
        # the chirp-z workspace holds the transposed intermediate (N x m complex128) plus three tables per axis
        L.lfd_set_mft_variant(2)
        need = L.lfd_mft_workspace_bytes(*bench_plane)
        assert 1024 * 1008 * 16 <= need <= 1024 * 1008 * 16 + (1 << 20)
        assert L.lfd_set_mft_variant(7) != 0 and b"unknown MFT variant" in L.lfd_last_error()
    finally:
        L.lfd_set_mft_variant(saved)


def test_bind_cpu_affinity_is_harmless_without_nvml():
    # on a box without a GPU / NVML the helper changes nothing and says so
    import os
    from lentil_b200 import device
    before = os.sched_getaffinity(0)
    got = device.bind_cpu_affinity(0)
    assert got is None or isinstance(got, set)
    if got is None:
        assert os.sched_getaffinity(0) == before
    else:
        os.sched_setaffinity(0, before)


def test_execution_names_map_to_the_descriptor_field():
    from lentil_b200 import fourier, _lib
    assert [fourier.execution_code(k) for k in (None, 'direct', 'folded', 'czt', 'auto')] == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        fourier.execution_code('fft')
    # a per-call choice overrides the process default in the host-side resolver (no device needed)
    L = _lib.lib()
    d = (_lib.MftDesc * 1)()
    d[0].m = d[0].n = 1001
    d[0].M = d[0].N = 1024
    assert L.lfd_mft_execution(d, 1) == 2                       # default AUTO -> chirp-z
    d[0].execution = 2
    assert L.lfd_mft_execution(d, 1) == 1                       # forced folded (1 + LFD_MFT_FOLDED)
    assert L.lfd_mft_c64_execution(d, 1) == 1                   # complex64: tcgen05 form
    d[0].execution = 4
    assert L.lfd_mft_execution(d, 1) == 2 and L.lfd_mft_c64_execution(d, 1) == 2
    d[0].m = 9000                                               # 10023 -> beyond the 8192-point kernel
    assert L.lfd_mft_execution(d, 1) == 1 and L.lfd_mft_c64_execution(d, 1) == 1
