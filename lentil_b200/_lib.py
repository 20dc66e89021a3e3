"""ctypes binding of liblentil_b200.so (the C ABI declared in include/lentil_b200.h).

The library is the product: there is no CPU fallback.  Importing this module never touches the
GPU; the first call that needs the library loads it and raises if it is missing.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LFD_LIB: development aid — load a tagged build (python -m lentil_b200.csrc.build --tag ...) instead of the production library
LIB_PATH = os.environ.get("LFD_LIB") or os.path.join(_HERE, "liblentil_b200.so")

_lib = None


class LfdError(RuntimeError):
    pass


class MftDesc(C.Structure):
    """struct lfd_mft_desc"""
    _fields_ = [
        ("f", C.c_void_p), ("ldf", C.c_int64),
        ("out", C.c_void_p), ("ldo", C.c_int64),
        ("m", C.c_int32), ("n", C.c_int32), ("M", C.c_int32), ("N", C.c_int32),
        ("alpha_r", C.c_double), ("alpha_c", C.c_double),
        ("shift_r", C.c_double), ("shift_c", C.c_double),
        ("off_r", C.c_double), ("off_c", C.c_double),
        ("unitary", C.c_int32), ("inverse", C.c_int32),
        ("execution", C.c_int32), ("reserved_", C.c_int32),
    ]


class PupilSrc(C.Structure):
    """struct lfd_pupil_src"""
    _fields_ = [
        ("amp", C.c_void_p), ("opd", C.c_void_p), ("mask", C.c_void_p),
        ("n_r", C.c_int32), ("n_c", C.c_int32), ("r0", C.c_int32), ("c0", C.c_int32),
        ("wavelength", C.c_double),
    ]


class Segment(C.Structure):
    """struct lfd_segment"""
    _fields_ = [
        ("r0", C.c_int32), ("c0", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("mask_index", C.c_int32), ("pad_", C.c_int32), ("out_offset", C.c_int64),
    ]


class Window(C.Structure):
    """struct lfd_window"""
    _fields_ = [
        ("E", C.c_void_p), ("ld", C.c_int64),
        ("h", C.c_int32), ("w", C.c_int32), ("r0", C.c_int32), ("c0", C.c_int32),
        ("group", C.c_int32), ("c64", C.c_int32), ("weight", C.c_double),
    ]


# name -> (restype, argtypes); the list doubles as the export manifest the CPU tests check
SIGNATURES = {
    "lfd_abi_version": (C.c_int, []),
    "lfd_last_error": (C.c_char_p, []),
    "lfd_launch_count": (C.c_uint64, []),
    "lfd_struct_size": (C.c_size_t, [C.c_int]),
    "lfd_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "lfd_set_mft_variant": (C.c_int, [C.c_int]),
    "lfd_get_mft_variant": (C.c_int, []),
    "lfd_mft_execution": (C.c_int, [C.POINTER(MftDesc), C.c_int]),
    "lfd_mft_workspace_bytes": (C.c_size_t, [C.POINTER(MftDesc), C.c_int]),
    "lfd_mft_c128_batched": (C.c_int, [C.POINTER(MftDesc), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lfd_mft_c128": (C.c_int, [C.POINTER(MftDesc), C.c_void_p, C.c_size_t, C.c_void_p]),
    "lfd_mft_c128_from_pupil": (C.c_int, [C.POINTER(MftDesc), C.POINTER(PupilSrc), C.c_int, C.c_int, C.c_void_p,
                                          C.c_size_t, C.c_void_p]),
    "lfd_mft_c64_execution": (C.c_int, [C.POINTER(MftDesc), C.c_int]),
    "lfd_mft_c64x3_workspace_bytes": (C.c_size_t, [C.POINTER(MftDesc), C.c_int]),
    "lfd_mft_c64x3_batched": (C.c_int, [C.POINTER(MftDesc), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lfd_mft_c64x3_from_pupil": (C.c_int, [C.POINTER(MftDesc), C.POINTER(PupilSrc), C.c_int, C.c_int, C.c_void_p,
                                          C.c_size_t, C.c_void_p]),
    "lfd_pupil_prep": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                 C.POINTER(Segment), C.c_int32, C.POINTER(C.c_double), C.c_int32,
                                 C.c_void_p, C.c_int64, C.c_void_p]),
    "lfd_pupil_prep_c64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.POINTER(Segment), C.c_int32, C.POINTER(C.c_double), C.c_int32,
                                     C.c_void_p, C.c_int64, C.c_void_p]),
    "lfd_accum_intensity": (C.c_int, [C.POINTER(Window), C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lfd_accum_field": (C.c_int, [C.POINTER(Window), C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                  C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lfd_field_mul": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_double,
                                C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p]),
    "lfd_fit_tilt_moments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double,
                                       C.POINTER(Segment), C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lfd_remove_tilt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                  C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lfd_mask_bbox": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "lfd_rebin": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "lfd_scale_separable": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lfd_abs_c128": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "lfd_opd_synth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                C.c_void_p]),
    "lfd_spline_prefilter": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    "lfd_spline_eval": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "lfd_sum_f64": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lfd_rescale_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_void_p,
                                     C.c_void_p]),
    "lfd_ctx_create": (C.c_void_p, [C.c_int]),
    "lfd_ctx_destroy": (None, [C.c_void_p]),
    "lfd_ctx_dft2_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_double, C.c_double, C.c_int32, C.c_int32,
                                    C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_int32, C.c_int32, C.c_void_p, C.c_int64]),
    "lfd_probe_fp64": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
}


def lib():
    """Load (once) and return the shared library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LfdError(
                f"{LIB_PATH} is missing: build it with `python -m lentil_b200.csrc.build` "
                "(lentil_b200 has no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.lfd_abi_version() != 2:
            raise LfdError("liblentil_b200.so ABI version mismatch")
        for which, struct in enumerate((MftDesc, Segment, Window)):
            if handle.lfd_struct_size(which) != C.sizeof(struct):
                raise LfdError(f"ABI struct layout mismatch for {struct.__name__}")
        v = os.environ.get("LFD_MFT_VARIANT")          # direct | folded | czt | auto: process-wide execution of K2a (default: the library's)
        if v:
            if handle.lfd_set_mft_variant({"direct": 0, "folded": 1, "czt": 2, "auto": 3}[v.lower()]) != 0:
                raise LfdError(f"LFD_MFT_VARIANT={v!r} rejected")
        _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().lfd_last_error().decode("utf-8", "replace")
        raise LfdError(f"{what or 'lentil_b200'} failed (rc={rc}): {msg}")
