// Detector-side sampling of an oversampled PSF — the step that follows
// the PSF accumulation in every lentil example (docs/examples/simple.rst:62-64).
//
//  * lfd_rebin            : integer-factor binning, lentil/util.py:221-258 (rebin: reshape + sum)
//  * lfd_scale_separable  : F[r,c] *= my[r] * mx[c] for a complex field — the pixel-MTF multiply of
//                           lentil/detector.py:213-220 (pixel): fft2(img) * outer(sinc, sinc); the two
//                           transforms around it are K2a launches (dft2 with alpha = 1/n is the FFT).
//  * lfd_abs_c128         : |z| of a complex field into float64 (detector.py:220, np.abs(ifft2(...)))
// All three are HBM-bound element-wise kernels.
#include "lfd_common.cuh"

namespace lfd {

__global__ void __launch_bounds__(256)
rebin_kernel(const double *__restrict__ img, long long ld, int factor, double *__restrict__ out, int oh, int ow) {
    const long long n = (long long)oh * ow;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / ow), c = (int)(e % ow);
        const double *p = img + (long long)r * factor * ld + (long long)c * factor;
        double s = 0.0;
        // same association as numpy's .sum(-1).sum(1): inner (column) sums first, then over the rows
        for (int i = 0; i < factor; ++i) {
            double row = 0.0;
            for (int j = 0; j < factor; ++j) row += p[(long long)i * ld + j];
            s += row;
        }
        out[e] = s;
    }
}

__global__ void __launch_bounds__(256)
scale_separable_kernel(double2 *__restrict__ F, long long ld, int h, int w, const double *__restrict__ my,
                       const double *__restrict__ mx) {
    const long long n = (long long)h * w;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / w), c = (int)(e % w);
        const double k = my[r] * mx[c];
        double2 v = F[(long long)r * ld + c];
        F[(long long)r * ld + c] = make_double2(v.x * k, v.y * k);
    }
}

__global__ void __launch_bounds__(256)
abs_c128_kernel(const double2 *__restrict__ F, long long ld, int h, int w, double *__restrict__ out) {
    const long long n = (long long)h * w;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(e / w), c = (int)(e % w);
        const double2 v = F[(long long)r * ld + c];
        out[e] = hypot(v.x, v.y);
    }
}

// OPD synthesis for Monte-Carlo wavefront error:
//   out[r][pix] = base[pix] + sum_k coeffs[r][k] * basis[k][pix]
// the np.einsum('ijk,i->jk', basis, coeff) of docs/user/wavefront_error.rst:118-135 for a batch of
// coefficient vectors.  One thread per pixel keeps its K basis values in registers and streams
// the R realisations: basis read once, 8 B/pixel/realisation written.
constexpr int SYN_MAX_K = 64;
__global__ void __launch_bounds__(256)
opd_synth_kernel(const double *__restrict__ basis, const double *__restrict__ coeffs, const double *__restrict__ base,
                 long long npix, int K, int R, int accumulate, double *__restrict__ out) {
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
         pix += (long long)gridDim.x * blockDim.x) {
        double b[SYN_MAX_K];
#pragma unroll 8
        for (int k = 0; k < K; ++k) b[k] = basis[(long long)k * npix + pix];
        const double b0 = base ? base[pix] : 0.0;
        for (int r = 0; r < R; ++r) {
            double v = accumulate ? out[(long long)r * npix + pix] : b0;
            for (int k = 0; k < K; ++k) v = fma(coeffs[(long long)r * K + k], b[k], v);
            out[(long long)r * npix + pix] = v;
        }
    }
}

static inline unsigned grid_for(long long n) {
    long long b = (n + 255) / 256;
    return (unsigned)(b > sm_or_default() * 16 ? sm_or_default() * 16 : (b < 1 ? 1 : b));
}

}  // namespace lfd

using namespace lfd;

extern "C" int lfd_rebin(const double *img, int64_t ld, int32_t h, int32_t w, int32_t factor, double *out, void *stream) {
    LFD_REQUIRE(img && out && factor > 0 && h >= factor && w >= factor && ld >= w, "lfd_rebin: bad arguments");
    const int oh = h / factor, ow = w / factor;
    rebin_kernel<<<grid_for((long long)oh * ow), 256, 0, (cudaStream_t)stream>>>(img, ld, factor, out, oh, ow);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int lfd_scale_separable(void *F, int64_t ld, int32_t h, int32_t w, const double *my, const double *mx, void *stream) {
    LFD_REQUIRE(F && my && mx && h > 0 && w > 0 && ld >= w, "lfd_scale_separable: bad arguments");
    scale_separable_kernel<<<grid_for((long long)h * w), 256, 0, (cudaStream_t)stream>>>((double2 *)F, ld, h, w, my, mx);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int lfd_abs_c128(const void *F, int64_t ld, int32_t h, int32_t w, double *out, void *stream) {
    LFD_REQUIRE(F && out && h > 0 && w > 0 && ld >= w, "lfd_abs_c128: bad arguments");
    abs_c128_kernel<<<grid_for((long long)h * w), 256, 0, (cudaStream_t)stream>>>((const double2 *)F, ld, h, w, out);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int lfd_opd_synth(const double *basis, const double *coeffs, const double *base, int64_t npix, int32_t K,
                             int32_t R, int32_t accumulate, double *out, void *stream) {
    LFD_REQUIRE(basis && coeffs && out && npix > 0 && K > 0 && R > 0, "lfd_opd_synth: bad arguments");
    LFD_REQUIRE(K <= SYN_MAX_K, "lfd_opd_synth: at most %d basis terms per call (got %d)", SYN_MAX_K, K);
    opd_synth_kernel<<<grid_for(npix), 256, 0, (cudaStream_t)stream>>>(basis, coeffs, base, npix, K, R, accumulate, out);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}
