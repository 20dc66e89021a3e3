"""The oracle restatement against the golden vectors the REAL reference produced
(oracle/make_golden.py).  CPU only."""
import numpy as np

import lentil_oracle as oc
from conftest import unpack_fields


def test_dft2_golden(golden):
    d = golden("dft2")
    for i in range(int(d["ncase"])):
        f, alpha = d[f"c{i}_f"], tuple(d[f"c{i}_alpha"])
        shape, shift, offset = tuple(d[f"c{i}_shape"]), tuple(d[f"c{i}_shift"]), tuple(d[f"c{i}_offset"])
        unitary = bool(d[f"c{i}_unitary"])
        F = oc.dft2(f, alpha, shape=shape, shift=shift, offset=offset, unitary=unitary)
        assert np.array_equal(F, d[f"c{i}_F"])
        iF = oc.idft2(f, alpha, shape=shape, shift=shift, unitary=unitary)
        assert np.array_equal(iF, d[f"c{i}_iF"])


def test_dft2_is_shifted_fft():
    # reference tests/test_fourier.py:7-44,104-111
    rng = np.random.default_rng(0)
    for m, n in [(10, 10), (11, 11), (10, 11)]:
        f = rng.random((m, n)) + 1j * rng.random((m, n))
        F = oc.dft2(f, [1 / m, 1 / n], unitary=False)
        assert np.allclose(F, np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(f))))
        g = oc.idft2(f, [1 / m, 1 / n], unitary=False)
        assert np.allclose(g, np.fft.fftshift(np.fft.ifft2(np.fft.ifftshift(f))))


def test_extent_golden(golden):
    rows = golden("extent")["rows"]
    for row in rows:
        sa, sb, ha, hb = tuple(row[0:2]), tuple(row[2:4]), tuple(row[4:6]), tuple(row[6:8])
        ea, eb = tuple(row[8:12]), tuple(row[12:16])
        assert oc.array_extent(sa, ha) == ea and oc.array_extent(sb, hb) == eb
        assert oc.intersect(ea, eb) == bool(row[16])
        ishape = oc.intersection_shape(ea, eb)
        assert (tuple(ishape) if ishape else (0, 0)) == tuple(row[17:19])
        assert oc.intersection_shift(ea, eb) == tuple(row[19:21])


def test_helper_golden(golden):
    d = golden("helper")
    for i in range(int(d["nh"])):
        a = d[f"h{i}_a"]
        slc = oc.boundary_slice(a)
        assert [slc[0].start, slc[0].stop, slc[1].start, slc[1].stop] == list(d[f"h{i}_slice"])
        assert tuple(oc.slice_offset(slc, a.shape)) == tuple(d[f"h{i}_offset"])


def test_field_golden(golden):
    d = golden("field")
    for i in range(int(d["n"])):
        c = oc.field_mul(oc.make_field(d[f"k{i}_a"], d[f"k{i}_ao"]), oc.make_field(d[f"k{i}_b"], d[f"k{i}_bo"]))
        if bool(d[f"k{i}_empty"]):
            assert c is None
        else:
            assert np.array_equal(c["data"], d[f"k{i}_c"]) and tuple(c["offset"]) == tuple(d[f"k{i}_co"])
    flds = [oc.make_field(d[f"r{i}_data"], d[f"r{i}_offset"]) for i in range(int(d["r_n"]))]
    out = oc.wavefront_insert(flds, np.zeros(d["r_intensity"].shape), 0.7)
    assert np.array_equal(out, d["r_intensity"])


def _check(fields, gold):
    assert len(fields) == len(gold)
    for f, (data, off) in zip(fields, gold):
        assert tuple(f["offset"]) == off
        assert np.array_equal(f["data"], data)


def test_propagate_golden(golden):
    d = golden("propagate")
    # A: monolithic + field-point tilt, three wavelengths accumulated
    dx = float(d["A_dx"])
    w0 = [oc.make_field(np.array(1, dtype=complex), None, [oc.tilt_entry(*d["A_tilt"])])]
    ph = oc.plane_multiply(w0, d["A_amp"], d["A_opd"], None, d["A_wls"][0])
    _check(ph, unpack_fields(d, "A_phasor"))
    pr, _ = oc.propagate_dft(ph, d["A_wls"][0], (dx, dx), float(d["A_z"]), float(d["A_du"]),
                             tuple(d["A_shape"]), None, int(d["A_oversample"]))
    _check(pr, unpack_fields(d, "A_prop"))
    img = oc.psf(d["A_amp"], d["A_opd"], None, d["A_wls"], d["A_wts"], (dx, dx), float(d["A_z"]),
                 float(d["A_du"]), tuple(d["A_shape"]), None, int(d["A_oversample"]), wf_tilt=d["A_tilt"])
    assert np.array_equal(img, d["A_img"])
    # B: segments + fit_tilt + prop_shape
    dx = float(d["B_dx"])
    ptilt = [tuple(t) for t in d["B_ptilt"]]
    ph = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], d["B_amp"], d["B_opd"],
                           d["B_mask"].astype(bool), float(d["B_wl"]), ptilt)
    _check(ph, unpack_fields(d, "B_phasor"))
    pr, so = oc.propagate_dft(ph, float(d["B_wl"]), (dx, dx), float(d["B_z"]), float(d["B_du"]),
                              tuple(d["B_shape"]), tuple(d["B_prop_shape"]), int(d["B_oversample"]))
    _check(pr, unpack_fields(d, "B_prop"))
    assert np.array_equal(oc.wavefront_intensity(pr, so), d["B_intensity"])
    assert np.array_equal(oc.wavefront_field(pr, so), d["B_field"])
    # C: detector mask
    dx = float(d["C_dx"])
    ph = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], d["C_amp"], d["C_opd"], None, float(d["C_wl"]))
    pr, so = oc.propagate_dft(ph, float(d["C_wl"]), (dx, dx), float(d["C_z"]), float(d["C_du"]), 32, None, 2,
                              d["C_omask"])
    _check(pr, unpack_fields(d, "C_prop"))
    assert np.array_equal(oc.wavefront_intensity(pr, so), d["C_intensity"])
    # D: PSF pushed off the detector -> no output fields
    ph = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], d["D_amp"], d["D_opd"], None, 650e-9,
                           [tuple(t) for t in d["D_ptilt"]])
    pr, _ = oc.propagate_dft(ph, 650e-9, (1 / 56, 1 / 56), 10.0, 5e-6, (16, 16), None, 2)
    assert pr == []


def test_detector_golden(golden):
    d = golden("detector")
    assert np.array_equal(oc.rebin(d["img"], 3), d["rebin3"])
    assert np.array_equal(oc.rebin(d["cube"], 2), d["rebin_cube2"])
    assert np.array_equal(oc.pixel(d["img"], 2), d["pixel2"])
    assert np.array_equal(oc.pixel(d["img"][:45, :45], 3), d["pixel3"])


def test_propagate_fft_golden(golden):
    # oracle restatement of lentil/propagate.py:9-88 against what the reference produced
    d = golden("propagate_fft")
    for i in range(int(d["n"])):
        amp, opd, wl = d[f"c{i}_amp"], d[f"c{i}_opd"], float(d[f"c{i}_wl"])
        dx, z, du, os_ = float(d[f"c{i}_dx"]), float(d[f"c{i}_z"]), float(d[f"c{i}_du"]), int(d[f"c{i}_os"])
        shape = None if d[f"c{i}_shape"][0] < 0 else tuple(int(v) for v in d[f"c{i}_shape"])
        fields = oc.plane_multiply([oc.make_field(np.array(1, dtype=complex))], amp, opd, None, wl)
        F, shape_out, pw = oc.propagate_fft(fields, amp.shape, wl, (dx, dx), z, du, shape, os_)
        assert tuple(shape_out) == tuple(d[f"c{i}_shape_out"]) and pw == float(d[f"c{i}_prop_wl"])
        if f"c{i}_F" in d:
            assert np.array_equal(F["data"], d[f"c{i}_F"])
        assert np.array_equal(oc.wavefront_intensity([F], shape_out), d[f"c{i}_intensity"])


def test_propagate_fft_refuses_tilt_and_oversized_shape():
    f = oc.make_field(np.ones((8, 8)), None, [oc.tilt_entry(1e-6, 0)])
    try:
        oc.propagate_fft([f], (8, 8), 600e-9, (1 / 8, 1 / 8), 10.0, 5e-6, (4, 4), 2)
        assert False
    except NotImplementedError:
        pass
    g = oc.make_field(np.ones((8, 8)))
    try:
        oc.propagate_fft([g], (8, 8), 600e-9, (1 / 8, 1 / 8), 10.0, 5e-6, (4000, 4000), 2)
        assert False
    except ValueError:
        pass


def test_power_spectrum_golden(golden):
    # oracle restatement of lentil/wfe.py:8-70 against the reference's output (same numpy generator, same FFT calls)
    d = golden("power_spectrum")
    for i in range(int(d["n"])):
        opd = oc.power_spectrum(d[f"c{i}_mask"], 1 / (2 * int(d[f"c{i}_radius"])), 30e-9, 5, 3, seed=int(d[f"c{i}_seed"]))
        assert np.array_equal(opd, d[f"c{i}_opd"])


def rescale_golden_cases(d):
    """(name, input key, args, kwargs) for every vector of tests/golden/rescale.npz."""
    s = d["scales"]
    cases = [(f"default_{k}", "img", (float(s[k]),), {}) for k in range(4)]
    cases += [(f"order{o}_third", "img", (1 / 3,), dict(order=o)) for o in range(6)]
    cases += [(f"mode_{m}", "img", (0.37,), dict(mode=m)) for m in ("constant", "reflect", "wrap")]
    cases += [("complex_half", "z", (0.5,), {})]
    return cases


def test_rescale_golden(golden):
    # lentil/util.py:261-347 through scipy.ndimage.map_coordinates: numerical pin (1e-13 of the peak)
    d = golden("rescale")
    for name, key, args, kw in rescale_golden_cases(d):
        got = oc.rescale(d[key], *args, **kw)
        assert got.shape == d[name].shape, name
        assert np.max(np.abs(got - d[name])) <= 1e-13 * np.max(np.abs(d[name])), name
    got = oc.rescale(d["img"], 0.5, shape=40, mask=np.ones_like(d["img"]), unitary=False)
    assert np.max(np.abs(got - d["shape40_ones_nonunitary"])) <= 1e-13 * np.max(d["shape40_ones_nonunitary"])
    for os_ in (3, 4):
        ref = d[f"pixelate{os_}"]
        got = oc.pixelate(d["psf"], os_)
        assert got.shape == ref.shape == (96 // os_, 96 // os_)
        assert np.max(np.abs(got - ref)) <= 1e-13 * ref.max()
