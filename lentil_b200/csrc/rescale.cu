// Spline rescale of the oversampled PSF to native sampling:
// lentil/util.py:261-347 (rescale) under lentil/detector.py:223-249 (pixelate).  The reference delegates
// the interpolation to scipy.ndimage.map_coordinates; what runs here is that algorithm for the meshgrid
// coordinates of util.py:329-332:
//
//  * lfd_spline_prefilter : edge-pad (mode 'nearest': 12 samples) and turn the image into B-spline
//                           coefficients, Unser's causal/anti-causal recursion per pole with the exact
//                           boundary initialisation (whole-sample "mirror" or half-sample "reflect"),
//                           axis 0 first, then axis 1 (through two tile transposes so that both passes
//                           run one thread per line with coalesced accesses).
//  * lfd_spline_eval      : out[i,j] = sum_a sum_b C[iy[i,a], ix[j,b]] * wy[i,a] * wx[j,b]; the tap tables
//                           (extension mode, B-spline weights) are per-row / per-column vectors the host
//                           computes (O(h + w) work).  With nonzero != 0 the source is read as (v != 0),
//                           the default mask of util.py:315-319.
//  * lfd_sum_f64          : deterministic two-stage sum (np.sum(img), np.sum(out) of util.py:344-345)
//  * lfd_rescale_finish   : out *= sum(img)/sum(out); out *= mask (mask < eps -> 0), util.py:335-347;
//                           real or complex (re/im planes in, complex128 out).
// All HBM/latency-bound; the image is a few MB (L2-resident).
#include "lfd_common.cuh"

namespace lfd {

constexpr int SPL_BATCH = 8;

__global__ void __launch_bounds__(256)
pad_edge_kernel(const double *__restrict__ img, long long ld, int h, int w, int npad, double *__restrict__ out) {
    const int H = h + 2 * npad, W = w + 2 * npad;
    const long long n = (long long)H * W;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / W) - npad, c = (int)(e % W) - npad;
        r = r < 0 ? 0 : (r >= h ? h - 1 : r);
        c = c < 0 ? 0 : (c >= w ? w - 1 : c);
        out[e] = img[(long long)r * ld + c];
    }
}

__global__ void __launch_bounds__(1024)
transpose_kernel(const double *__restrict__ in, int h, int w, double *__restrict__ out) {
    __shared__ double tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    int r = r0 + threadIdx.y, c = c0 + threadIdx.x;
    if (r < h && c < w) tile[threadIdx.y][threadIdx.x] = in[(long long)r * w + c];
    __syncthreads();
    r = c0 + threadIdx.y;   // row of the transposed array
    c = r0 + threadIdx.x;
    if (r < w && c < h) out[(long long)r * h + c] = tile[threadIdx.x][threadIdx.y];
}

// One thread per column of the dense n x w array `c`; all poles in sequence, in place.
// reflect != 0: half-sample-symmetric boundary, else whole-sample-symmetric.
__global__ void __launch_bounds__(128)
spline_filter_cols_kernel(double *__restrict__ c, int n, int w, double z0, double z1, int npoles, int reflect) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w || n < 2) return;
    double *col = c + j;
    const long long ld = w;
    double gain = (1.0 - z0) * (1.0 - 1.0 / z0);
    if (npoles > 1) gain *= (1.0 - z1) * (1.0 - 1.0 / z1);
    for (int p = 0; p < npoles; ++p) {
        const double z = p ? z1 : z0;
        const double g = p ? 1.0 : gain;      // the gain is applied to the line before the first pole
        // ---- causal initialisation: sum_i z^i (s[i] + z^L s[n-1-i]) over the symmetric extension.
        // Terms beyond |z|^i < 1e-40 cannot change a double sum of comparable samples: stop there.
        const int L = reflect ? n : n - 1;
        const double zL = pow(z, (double)L);
        int horizon = (int)ceil(-92.1 / log(fabs(z))) + 1;
        if (horizon > L) horizon = L;
        const double s0 = g * col[0];
        double acc = s0 + zL * (g * col[(long long)(n - 1) * ld]);
        double zi = z;
        for (int i = 1; i < horizon; ++i) {
            acc += zi * (g * col[(long long)i * ld] + zL * (g * col[(long long)(n - 1 - i) * ld]));
            zi *= z;
        }
        double prev = reflect ? acc * (z / (1.0 - zL * zL)) + s0 : acc / (1.0 - zL * zL);
        col[0] = prev;
        // ---- causal recursion c[i] += z c[i-1], loads of a batch issued together
        double beforelast = prev;
        for (int i = 1; i < n; i += SPL_BATCH) {
            double v[SPL_BATCH];
#pragma unroll
            for (int k = 0; k < SPL_BATCH; ++k) v[k] = (i + k < n) ? col[(long long)(i + k) * ld] : 0.0;
#pragma unroll
            for (int k = 0; k < SPL_BATCH; ++k) {
                if (i + k < n) {
                    beforelast = prev;
                    prev = g * v[k] + z * prev;
                    col[(long long)(i + k) * ld] = prev;
                }
            }
        }
        // ---- anti-causal initialisation and recursion c[i] = z (c[i+1] - c[i])
        double next = reflect ? prev * (z / (z - 1.0)) : (z * beforelast + prev) * z / (z * z - 1.0);
        col[(long long)(n - 1) * ld] = next;
        for (int i = n - 2; i >= 0; i -= SPL_BATCH) {
            double v[SPL_BATCH];
#pragma unroll
            for (int k = 0; k < SPL_BATCH; ++k) v[k] = (i - k >= 0) ? col[(long long)(i - k) * ld] : 0.0;
#pragma unroll
            for (int k = 0; k < SPL_BATCH; ++k) {
                if (i - k >= 0) {
                    next = z * (next - v[k]);
                    col[(long long)(i - k) * ld] = next;
                }
            }
        }
    }
}

constexpr int SPL_MAX_TAPS = 6;

__global__ void __launch_bounds__(256)
spline_eval_kernel(const double *__restrict__ C, long long ld, int nonzero, const int *__restrict__ iy,
                   const double *__restrict__ wy, int ny, const int *__restrict__ ix, const double *__restrict__ wx,
                   int nx, int ntaps, double *__restrict__ out) {
    const long long n = (long long)ny * nx;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e / nx), j = (int)(e % nx);
        int cx[SPL_MAX_TAPS];
        double vx[SPL_MAX_TAPS];
        for (int b = 0; b < ntaps; ++b) { cx[b] = ix[j * ntaps + b]; vx[b] = wx[j * ntaps + b]; }
        double t = 0.0;
        for (int a = 0; a < ntaps; ++a) {
            const double *row = C + (long long)iy[i * ntaps + a] * ld;
            const double va = wy[i * ntaps + a];
            for (int b = 0; b < ntaps; ++b) {
                double v = row[cx[b]];
                if (nonzero) v = (v != 0.0) ? 1.0 : 0.0;
                // same association as the C loop of map_coordinates: (coefficient * wy) * wx, summed in tap order
                t = __dadd_rn(t, __dmul_rn(__dmul_rn(v, va), vx[b]));
            }
        }
        out[e] = t;
    }
}

constexpr int SUM_BLOCKS = 256;

__device__ __forceinline__ double block_sum_256(double v) {
    __shared__ double sh[8];
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    v = (threadIdx.x < 8) ? sh[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) for (int o = 4; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    return v;
}

__global__ void __launch_bounds__(256)
sum_stage1_kernel(const double *__restrict__ x, long long n, double *__restrict__ partials) {
    double v = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) v += x[e];
    v = block_sum_256(v);
    if (threadIdx.x == 0) partials[blockIdx.x] = v;
}

__global__ void __launch_bounds__(256)
sum_stage2_kernel(const double *__restrict__ partials, int count, double *__restrict__ out) {
    double v = (threadIdx.x < count) ? partials[threadIdx.x] : 0.0;
    v = block_sum_256(v);
    if (threadIdx.x == 0) out[0] = v;
}

__global__ void __launch_bounds__(256)
rescale_finish_kernel(const double *__restrict__ re, const double *__restrict__ im, const double *__restrict__ mask,
                      const double *__restrict__ sums, double eps, long long n, double *__restrict__ out) {
    // factor = sum(img) / sum(out), complex when an imaginary plane is present
    double fr = 1.0, fi = 0.0;
    if (sums) {
        if (im) {
            const double a = sums[0], b = sums[1], c = sums[2], d = sums[3];
            const double den = c * c + d * d;
            fr = (a * c + b * d) / den;
            fi = (b * c - a * d) / den;
        } else {
            fr = sums[0] / sums[2];
        }
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        double m = 1.0;
        if (mask) { m = mask[e]; if (m < eps) m = 0.0; }
        if (im) {
            const double x = re[e], y = im[e];
            out[2 * e] = (x * fr - y * fi) * m;
            out[2 * e + 1] = (x * fi + y * fr) * m;
        } else {
            out[e] = (re[e] * fr) * m;
        }
    }
}

static inline unsigned spl_grid(long long n) {
    long long b = (n + 255) / 256;
    return (unsigned)(b > sm_or_default() * 16 ? sm_or_default() * 16 : (b < 1 ? 1 : b));
}

static int spline_poles(int order, double *z) {
    switch (order) {
    case 2: z[0] = sqrt(8.0) - 3.0; return 1;
    case 3: z[0] = sqrt(3.0) - 2.0; return 1;
    case 4:
        z[0] = sqrt(664.0 - sqrt(438976.0)) + sqrt(304.0) - 19.0;
        z[1] = sqrt(664.0 + sqrt(438976.0)) - sqrt(304.0) - 19.0;
        return 2;
    case 5:
        z[0] = sqrt(67.5 - sqrt(4436.25)) + sqrt(26.25) - 6.5;
        z[1] = sqrt(67.5 + sqrt(4436.25)) - sqrt(26.25) - 6.5;
        return 2;
    default: return 0;
    }
}

}  // namespace lfd

using namespace lfd;

extern "C" int lfd_spline_prefilter(const double *img, int64_t ld, int32_t h, int32_t w, int32_t npad, int32_t order,
                                    int32_t reflect, double *coef, double *scratch, void *stream) {
    LFD_REQUIRE(img && coef && scratch && h > 0 && w > 0 && ld >= w && npad >= 0, "lfd_spline_prefilter: bad arguments");
    LFD_REQUIRE(order >= 0 && order <= 5, "lfd_spline_prefilter: spline order %d not supported (0..5)", order);
    cudaStream_t s = (cudaStream_t)stream;
    const int H = h + 2 * npad, W = w + 2 * npad;
    pad_edge_kernel<<<spl_grid((long long)H * W), 256, 0, s>>>(img, ld, h, w, npad, coef);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    double z[2] = {0.0, 0.0};
    const int npoles = spline_poles(order, z);
    if (npoles == 0) return 0;                     // orders 0 and 1 interpolate the samples themselves
    const dim3 tb(32, 32);
    spline_filter_cols_kernel<<<(W + 127) / 128, 128, 0, s>>>(coef, H, W, z[0], z[1], npoles, reflect);
    transpose_kernel<<<dim3((W + 31) / 32, (H + 31) / 32), tb, 0, s>>>(coef, H, W, scratch);
    spline_filter_cols_kernel<<<(H + 127) / 128, 128, 0, s>>>(scratch, W, H, z[0], z[1], npoles, reflect);
    transpose_kernel<<<dim3((H + 31) / 32, (W + 31) / 32), tb, 0, s>>>(scratch, W, H, coef);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch(4);
    return 0;
}

extern "C" int lfd_spline_eval(const double *coef, int64_t ld, int32_t nonzero, const int32_t *iy, const double *wy,
                               int32_t ny, const int32_t *ix, const double *wx, int32_t nx, int32_t ntaps, double *out,
                               void *stream) {
    LFD_REQUIRE(coef && iy && wy && ix && wx && out && ny > 0 && nx > 0, "lfd_spline_eval: bad arguments");
    LFD_REQUIRE(ntaps >= 1 && ntaps <= SPL_MAX_TAPS, "lfd_spline_eval: %d taps per axis (1..%d)", ntaps, SPL_MAX_TAPS);
    spline_eval_kernel<<<spl_grid((long long)ny * nx), 256, 0, (cudaStream_t)stream>>>(coef, ld, nonzero, iy, wy, ny, ix, wx,
                                                                                     nx, ntaps, out);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int lfd_sum_f64(const double *x, int64_t n, double *partials, double *out, void *stream) {
    LFD_REQUIRE(x && partials && out && n > 0, "lfd_sum_f64: bad arguments");
    sum_stage1_kernel<<<SUM_BLOCKS, 256, 0, (cudaStream_t)stream>>>(x, n, partials);
    sum_stage2_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partials, SUM_BLOCKS, out);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch(2);
    return 0;
}

extern "C" int lfd_rescale_finish(const double *re, const double *im, const double *mask, const double *sums, double eps,
                                  int64_t n, double *out, void *stream) {
    LFD_REQUIRE(re && out && n > 0, "lfd_rescale_finish: bad arguments");
    rescale_finish_kernel<<<spl_grid(n), 256, 0, (cudaStream_t)stream>>>(re, im, mask, sums, eps, n, out);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}
