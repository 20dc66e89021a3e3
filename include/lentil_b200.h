/*
 * lentil_b200 — C-ABI of the B200-native far-field diffraction path.
 *
 * Drop-in boundary for the hot path of andykee/lentil (v0.8.8).  lentil has no FFI of its
 * own: its "operator API" for this path is a handful of Python callables.  Every entry point
 * below names the reference interface (file:line under the lentil source tree) it replaces;
 * the Python shim in lentil_b200/ binds them with ctypes.
 *
 * Conventions
 *   - plain C, no CUDA/torch types: device pointers travel as void*, streams as void*
 *     (a cudaStream_t; NULL = the legacy default stream).
 *   - complex128 = interleaved (re, im) doubles, exactly numpy's layout.  Leading dimensions
 *     (ld*) are in ELEMENTS (complex elements for complex arrays).
 *   - every function returns 0 on success, non-zero on failure; lfd_last_error() then holds a
 *     message (thread-local).  No exceptions cross the ABI.  Nothing allocates behind the
 *     caller's back except the lfd_ctx_* host-buffer convenience layer.
 *   - "dev" in a name/param = device memory; "host" = host memory.
 */
#ifndef LENTIL_B200_H
#define LENTIL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LFD_ABI_VERSION 2

/* ---- errors / introspection ------------------------------------------------------------ */
int         lfd_abi_version(void);
const char *lfd_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py gpu_launches) */
uint64_t    lfd_launch_count(void);
/* sizeof() of the ABI structs as compiled: 0 = lfd_mft_desc, 1 = lfd_segment, 2 = lfd_window
 * (bindings assert their own layout against these) */
size_t      lfd_struct_size(int which);
/* device properties the host planner needs: [0]=SM count, [1]=cc major, [2]=cc minor */
int         lfd_device_info(int device, int *out3);

/* ---- K2a: matrix Fourier transform, complex128 ------------------------------------------
 * One plane of  F = E1 . f . E2 . sqrt|ar*ac|   with
 *   E1[k,i] = exp(-2 pi i ar (R_i + off_r)(U_k - shift_r)),  R_i = i - floor(m/2), U_k = k - floor(M/2)
 *   E2[j,l] = exp(-2 pi i ac (S_j + off_c)(V_l - shift_c))
 * replaces lentil/fourier.py:5-103 (dft2) with E1/E2 of :106-121 generated on the fly, and
 * lentil/fourier.py:124-198 (idft2 = conj(dft2(conj F))/F.size) when inverse != 0.
 */
typedef struct lfd_mft_desc {
    const void *f;        /* dev, complex128 m x n, row-major                        */
    int64_t     ldf;      /* elements between rows of f (>= n)                        */
    void       *out;      /* dev, complex128 M x N, row-major; may NOT alias f        */
    int64_t     ldo;      /* elements between rows of out (>= N)                      */
    int32_t     m, n;     /* input shape                                              */
    int32_t     M, N;     /* output shape                                             */
    double      alpha_r, alpha_c;
    double      shift_r, shift_c;   /* output-plane shift of the DC pixel (pixels)     */
    double      off_r, off_c;       /* input-plane offset (pixels)                     */
    int32_t     unitary;            /* fourier.py:100-101                              */
    int32_t     inverse;            /* 0: dft2, 1: idft2 semantics                     */
    int32_t     execution;          /* 0: the process default (lfd_set_mft_variant); 1 + LFD_MFT_x: run this batch with
                                       execution x whatever the default is.  Read from the FIRST descriptor of a batch
                                       (one launch = one execution); per-call, so concurrent threads / streams can
                                       choose independently                                                   */
    int32_t     reserved_;          /* must be 0                                       */
} lfd_mft_desc;

/* Three executions of the same transform (results agree to rounding):
 *   LFD_MFT_DIRECT : complex twiddle x complex data, 4 real DMMAs per complex 8x8x4 block
 *   LFD_MFT_FOLDED : even/odd folding of both axes -> real twiddles, 4x fewer DMMAs (default)
 *   LFD_MFT_CZT    : chirp-z (Bluestein) execution on the FP64 pipe: per row FFT_L -> x FFT(chirp) -> IFFT_L in shared
 *                    memory, ~8x fewer flops than the folded form; planes whose FFT length would exceed 8192
 *                    (n_in + n_out - 1 > 8192 on an axis) are run by the folded execution instead
 * lfd_set_mft_variant sets the process-wide DEFAULT (affects lfd_mft_workspace_bytes and the lfd_mft_* launches that
 * follow, for descriptors whose `execution` field is 0); lfd_mft_desc.execution overrides it per call. */
#define LFD_MFT_DIRECT 0
#define LFD_MFT_FOLDED 1
#define LFD_MFT_CZT    2
#define LFD_MFT_AUTO   3   /* default: chirp-z whenever every plane of the batch fits it (FFT length <= 8192 on both axes;
                            * measured faster than the folded form at every size from 128^2 up), else folded */
int lfd_set_mft_variant(int variant);
int lfd_get_mft_variant(void);
/* which execution (LFD_MFT_DIRECT / FOLDED / CZT) a batch runs under the current setting */
int lfd_mft_execution(const lfd_mft_desc *descs_host, int count);

/* bytes of device workspace lfd_mft_c128_batched needs for `count` planes whose largest
 * intermediate is max over planes of (n * M) complex elements */
size_t lfd_mft_workspace_bytes(const lfd_mft_desc *descs_host, int count);

/* run `count` independent planes (descriptors in HOST memory, data in device memory) */
int lfd_mft_c128_batched(const lfd_mft_desc *descs_host, int count,
                         void *workspace_dev, size_t workspace_bytes, void *stream);

/* single plane convenience (same as count == 1) */
int lfd_mft_c128(const lfd_mft_desc *desc_host, void *workspace_dev, size_t workspace_bytes,
                 void *stream);

/* Fused K1 + K2a (chirp-z and folded executions; a batch whose execution resolves to DIRECT runs the folded
 * one): the input of plane i is not read from desc.f but formed on the fly as amp * mask *
 * exp(+2 pi i opd / wavelength) over the m x n window at (r0, c0) of the pupil arrays
 * (lentil/plane.py:502-507), so phasors never exist in HBM.  With intensity_out != 0 the result written
 * to desc.out is |F|^2 as float64 (ldo in doubles) instead of the complex field — valid when the plane
 * is the only Field of its wavefront (no coherent merge, lentil/field.py:413-461).
 * Workspace: lfd_mft_workspace_bytes for the same descriptors (with DIRECT: size it with execution = 1 +
 * LFD_MFT_FOLDED). */
typedef struct lfd_pupil_src {
    const double  *amp, *opd;     /* dev, n_r x n_c float64                                   */
    const uint8_t *mask;          /* dev, this segment's n_r x n_c mask plane, or NULL         */
    int32_t        n_r, n_c, r0, c0;
    double         wavelength;
} lfd_pupil_src;
int lfd_mft_c128_from_pupil(const lfd_mft_desc *descs_host, const lfd_pupil_src *src_host, int count,
                            int intensity_out, void *workspace_dev, size_t workspace_bytes, void *stream);

/* ---- K2b: the same transform with complex64 in/out, 3xTF32 split precision on tcgen05 -------
 * Same descriptor, but `f` and `out` are complex64 (interleaved floats).  Every real product is
 * evaluated as hi*hi + hi*lo + lo*hi in TF32 with fp32 accumulation in TMEM; peak-normalised error
 * ~1e-6 (gate 1e-5).  lentil itself has no complex64 mode (Field casts to complex128,
 * lentil/field.py:35): this is the optional fast path BASELINE.json's north star names.
 *
 * Two executions, chosen like K2a's (lfd_mft_desc.execution of the first plane, else the process default):
 *   LFD_MFT_DIRECT / LFD_MFT_FOLDED : the folded transform as 3xTF32 tcgen05 MMAs (mft_c64.cu)
 *   LFD_MFT_CZT / LFD_MFT_AUTO      : the FP32 build of the chirp-z row transform (mft_czt.cu; chirps and the transformed
 *                                     chirp filter are built in float64 and rounded once) whenever every plane fits it,
 *                                     else the tcgen05 form.  Peak-normalised error ~3e-7 at any size.
 * lfd_mft_c64_execution reports which one a batch runs (LFD_MFT_FOLDED = tcgen05, LFD_MFT_CZT = chirp-z). */
int    lfd_mft_c64_execution(const lfd_mft_desc *descs_host, int count);
size_t lfd_mft_c64x3_workspace_bytes(const lfd_mft_desc *descs_host, int count);
int lfd_mft_c64x3_batched(const lfd_mft_desc *descs_host, int count,
                          void *workspace_dev, size_t workspace_bytes, void *stream);
/* K1 fused into K2b, the complex64 twin of lfd_mft_c128_from_pupil: desc.f / ldf are ignored, the fold kernel forms
 * amp * mask * exp(+2 pi i opd / lambda) itself (phase reduced in float64, phasor rounded to complex64;
 * lentil/plane.py:502-507).  With intensity_out != 0 desc.out receives |F|^2 as float64 (ldo in doubles). */
int lfd_mft_c64x3_from_pupil(const lfd_mft_desc *descs_host, const lfd_pupil_src *src_host, int count,
                             int intensity_out, void *workspace_dev, size_t workspace_bytes, void *stream);

/* ---- K1: pupil prep ------------------------------------------------------------------
 * For each wavelength w and segment s:  out_{w,s}[r,c] = amp[r,c] * mask_s[r,c] *
 * exp(+2 pi i opd[r,c] / lambda_w)  over the segment's bounding box.
 * replaces the phasor build in lentil/plane.py:494-510 (Plane.__mul__) for array operands.
 */
typedef struct lfd_segment {
    int32_t r0, c0;      /* upper-left corner of the bbox in the pupil array            */
    int32_t h, w;        /* bbox shape                                                  */
    int32_t mask_index;  /* which mask plane (0 for a 2-D mask)                         */
    int32_t pad_;
    int64_t out_offset;  /* element offset of this segment's (h x w, ld = w) tile inside
                            one wavelength's output block                               */
} lfd_segment;

int lfd_pupil_prep(const double *amp_dev, const double *opd_dev, /* n_r x n_c, row-major */
                   const uint8_t *mask_dev,                       /* nmask x n_r x n_c, 0/1; NULL = all ones */
                   int32_t n_r, int32_t n_c,
                   const lfd_segment *segs_host, int32_t nseg,
                   const double *wavelengths_host, int32_t nlam,
                   void *out_dev,               /* complex128, nlam blocks              */
                   int64_t out_lam_stride,      /* elements between wavelength blocks   */
                   void *stream);

/* same, writing complex64 phasors (input of lfd_mft_c64x3_batched) */
int lfd_pupil_prep_c64(const double *amp_dev, const double *opd_dev, const uint8_t *mask_dev,
                       int32_t n_r, int32_t n_c, const lfd_segment *segs_host, int32_t nseg,
                       const double *wavelengths_host, int32_t nlam,
                       void *out_dev, int64_t out_lam_stride, void *stream);

/* ---- K3: coherent merge + |E|^2 accumulate ---------------------------------------------
 * I[r,c] += sum over groups g of  weight_g * | sum over windows v in g covering (r,c) of E_v |^2
 * replaces lentil/field.py:231-305 (insert), :308-346 (merge), :413-461 (reduce) and
 * lentil/wavefront.py:114-165 (intensity / insert).  Owner-computes: one thread per output
 * pixel, groups visited in order, so results are run-to-run bit-identical.
 */
typedef struct lfd_window {
    const void *E;       /* dev complex128 h x w                                          */
    int64_t     ld;      /* elements between rows                                         */
    int32_t     h, w;
    int32_t     r0, c0;  /* position of E[0,0] in the output image (may be negative/clipped) */
    int32_t     group;   /* windows with equal group are summed coherently; groups must be
                            contiguous and non-decreasing in the array                    */
    int32_t     c64;     /* 0: E is complex128; 1: E is complex64 (output of the K2b path);
                            2: E is a float64 INTENSITY window (lfd_mft_c128_from_pupil, intensity_out) */
    double      weight;  /* weight of the group (taken from its first window)             */
} lfd_window;

int lfd_accum_intensity(const lfd_window *wins_host, int32_t nwin,
                        double *I_dev, int32_t H, int32_t W, int64_t ldI,
                        void *scratch_dev, size_t scratch_bytes, /* >= nwin*sizeof(lfd_window) */
                        void *stream);

/* complex (field) insert:  out[r,c] += weight * E_v  — replaces field.py:303-304 / wavefront.py:101-112 */
int lfd_accum_field(const lfd_window *wins_host, int32_t nwin,
                    void *out_dev, int32_t H, int32_t W, int64_t ldo,
                    void *scratch_dev, size_t scratch_bytes, void *stream);

/* element-wise product of two complex128 windows times a complex scalar:
 *   out[r,c] = a[r,c] * b[r,c] * (s_re + i s_im)   (b == NULL: out = a * scalar)
 * the caller points a/b at the upper-left corner of the rectangle intersection.
 * replaces the array branch of Field.__mul__, lentil/field.py:130-147. */
int lfd_field_mul(const void *a_dev, int64_t lda, const void *b_dev, int64_t ldb,
                  double s_re, double s_im, void *out_dev, int64_t ldo,
                  int32_t h, int32_t w, void *stream);

/* ---- fit_tilt -------------------------------------------------
 * Per-segment least-squares fit of [1, r*dx0, -c*dx1] to the OPD over the segment mask and removal
 * of the two tilt terms: replaces Plane.fit_tilt / ptt_vector, lentil/plane.py:522-611.
 * lfd_fit_tilt_moments reduces, per segment, the 9 moments S1 Sx Sy Sxx Sxy Syy Sz Sxz Syz with x, y
 * measured from the centre of the segment's bounding box (x_c = (r0 + (h-1)/2 - n_r/2)*dx0,
 * y_c = -(c0 + (w-1)/2 - n_c/2)*dx1); the caller solves the 3x3 systems.  mask_dev: uint8 cube
 * (nmask x n_r x n_c) or NULL, in which case the mask is amp_for_mask != 0 (lentil/plane.py:43-47). */
int lfd_fit_tilt_moments(const double *opd_dev, const uint8_t *mask_dev, const double *amp_for_mask_dev,
                         int32_t n_r, int32_t n_c, double dx0, double dx1,
                         const lfd_segment *segs_host, int32_t nseg, double *moments_dev /* nseg x 9 */,
                         void *scratch_dev, size_t scratch_bytes /* >= nseg * 40 */, void *stream);
/* out = opd with each segment's tilt removed: (opd - t1*x*m - t2*y*m) summed over segments
 * (single segment: pixels outside the mask keep their OPD).  coef_dev: nseg x 3 (piston, t1, t2). */
int lfd_remove_tilt(const double *opd_dev, const uint8_t *mask_dev, const double *amp_for_mask_dev,
                    int32_t n_r, int32_t n_c, int32_t nseg, double dx0, double dx1,
                    const double *coef_dev, double *out_dev, void *stream);

/* Bounding box (rmin, rmax, cmin, cmax; inclusive) of the support of each of `nplanes` mask planes
 * (n_r x n_c, uint8 or float64): entries > 0, or != 0 when `nonzero` (a mask derived from an
 * amplitude, lentil/plane.py:43-47).  An empty plane yields (n_r, -1, n_c, -1).
 * replaces lentil/util.py:190-218 (boundary) under helper.boundary_slice / plane._plane_slice. */
int lfd_mask_bbox(const void *x_dev, int32_t is_f64, int32_t nonzero, int32_t n_r, int32_t n_c,
                  int32_t nplanes, int32_t *out_dev /* 4 x nplanes */, void *stream);

/* ---- detector-side sampling of the oversampled PSF -----------------
 * lfd_rebin: out[r,c] = sum of the factor x factor block of img — lentil/util.py:221-258 (rebin).
 * lfd_scale_separable: F[r,c] *= my[r]*mx[c] (complex128, in place) — the pixel-MTF multiply of
 *   lentil/detector.py:213-220 (pixel); lfd_abs_c128: out = |F| (detector.py:220). */
int lfd_rebin(const double *img_dev, int64_t ld, int32_t h, int32_t w, int32_t factor, double *out_dev, void *stream);
int lfd_scale_separable(void *F_dev, int64_t ld, int32_t h, int32_t w, const double *my_dev, const double *mx_dev, void *stream);
int lfd_abs_c128(const void *F_dev, int64_t ld, int32_t h, int32_t w, double *out_dev, void *stream);

/* OPD synthesis: out[r] = base + sum_k coeffs[r][k] * basis[k], the
 * np.einsum('ijk,i->jk', basis, coeff) of docs/user/wavefront_error.rst:118-135 for R coefficient
 * vectors at once.  basis: K x npix, coeffs: R x K (device), base: npix or NULL, out: R x npix. K <= 64 per
 * call; accumulate != 0 adds to what `out` already holds (to chain more than 64 terms). */
int lfd_opd_synth(const double *basis_dev, const double *coeffs_dev, const double *base_dev, int64_t npix,
                  int32_t K, int32_t R, int32_t accumulate, double *out_dev, void *stream);

/* ---- spline rescale to native sampling ---------------------------
 * lentil/util.py:261-347 (rescale) under lentil/detector.py:223-249 (pixelate); the reference calls
 * scipy.ndimage.map_coordinates there (util.py:334-343).
 * lfd_spline_prefilter: coef ((h+2*npad) x (w+2*npad), dense) = B-spline coefficients of the image padded
 *   by npad edge samples; reflect != 0 selects the half-sample-symmetric boundary (map_coordinates modes
 *   'nearest' and 'reflect'), 0 the whole-sample one ('mirror', 'constant', legacy 'wrap'). Orders 0 and 1
 *   only pad. `scratch` has the size of `coef`.
 * lfd_spline_eval: out[i,j] = sum_a sum_b coef[iy[i,a], ix[j,b]] * wy[i,a] * wx[j,b]; iy/wy are ny x ntaps,
 *   ix/wx nx x ntaps (device).  nonzero != 0 reads the source as (v != 0) — the default mask of util.py:315-319.
 * lfd_sum_f64: out[0] = sum of n doubles, fixed two-stage order; `partials` holds >= 256 doubles.
 * lfd_rescale_finish: out = (re [+ i im]) * sum(img)/sum(out) * mask, mask < eps -> 0 (util.py:335-347).
 *   sums = {Re sum(img), Im sum(img), Re sum(out), Im sum(out)} on the device or NULL (unitary=False);
 *   im == NULL: real output (n doubles), else complex128 output (n elements); mask may be NULL. */
int lfd_spline_prefilter(const double *img_dev, int64_t ld, int32_t h, int32_t w, int32_t npad, int32_t order,
                         int32_t reflect, double *coef_dev, double *scratch_dev, void *stream);
int lfd_spline_eval(const double *coef_dev, int64_t ld, int32_t nonzero, const int32_t *iy_dev, const double *wy_dev,
                    int32_t ny, const int32_t *ix_dev, const double *wx_dev, int32_t nx, int32_t ntaps,
                    double *out_dev, void *stream);
int lfd_sum_f64(const double *x_dev, int64_t n, double *partials_dev, double *out_dev, void *stream);
int lfd_rescale_finish(const double *re_dev, const double *im_dev, const double *mask_dev, const double *sums_dev,
                       double eps, int64_t n, double *out_dev, void *stream);

/* ---- host-buffer convenience layer (what bench.py's e2e leg and the numpy shim call) ------
 * A context owns a device workspace, pinned staging buffers and one stream on `device`.
 */
typedef struct lfd_ctx lfd_ctx;
lfd_ctx *lfd_ctx_create(int device);
void     lfd_ctx_destroy(lfd_ctx *ctx);

/* dft2 / idft2 with numpy (host) buffers in and out: H2D, K2a, D2H, synchronises.
 * `f_host` m x n complex128 (ldf elements), `out_host` M x N complex128 (ldo elements);
 * out_host may alias f_host (tests/test_fourier.py:97-101). */
int lfd_ctx_dft2_host(lfd_ctx *ctx, const void *f_host, int64_t ldf, int32_t m, int32_t n,
                      double alpha_r, double alpha_c, int32_t M, int32_t N,
                      double shift_r, double shift_c, double off_r, double off_c,
                      int32_t unitary, int32_t inverse, void *out_host, int64_t ldo);

/* ---- diagnostics --------------------------------------------------------------------- */
/* DMMA.8x8x4 / DFMA issue-rate probe on the current device: out[0] = DMMA TFLOP/s,
 * out[1] = DFMA TFLOP/s, out[2] = SM clock (MHz) seen by the probe. */
int lfd_probe_fp64(double *out3, int iters);

#ifdef __cplusplus
}
#endif
#endif /* LENTIL_B200_H */
