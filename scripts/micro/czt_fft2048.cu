// Feasibility probe for a chirp-z (Bluestein) execution of the matrix Fourier transform in FP64 ("what comes
// next"): the inner kernel of such a path is, per row, a 2048-point complex128 FFT in shared memory, a point-wise product
// with the transformed chirp, and the inverse FFT:  y = IFFT(FFT(pre * x, zero-padded) * H) * post.  One plane of the bench
// workload (1001^2 -> 1024^2) is 1001 + 1024 such rows (m + M - 1 = 2024 <= 2048), i.e. ~0.5 GFLOP instead of the 4.15 GFLOP
// the folded DMMA kernel executes.  This file measures what the row kernel achieves and checks it against a direct
// convolution.  Development aid:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o czt_fft2048 czt_fft2048.cu && ./czt_fft2048
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int N = 2048, T = 256;
constexpr int NP = N + N / 8;                                   // one element of padding per 8: stride-8 stores are conflict-free
__device__ __forceinline__ int P(int i) { return i + (i >> 3); }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i (S = +1, forward) or +i (S = -1, inverse)
template <int S> __device__ __forceinline__ double2 mul_mi(double2 a) { return S > 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x); }

template <int S> __device__ __forceinline__ void dft4(double2 &x0, double2 &x1, double2 &x2, double2 &x3) {
    const double2 s0 = cadd(x0, x2), s1 = csub(x0, x2), s2 = cadd(x1, x3), s3 = mul_mi<S>(csub(x1, x3));
    x0 = cadd(s0, s2); x2 = csub(s0, s2); x1 = cadd(s1, s3); x3 = csub(s1, s3);
}
// 8-point DFT, natural order in and out
template <int S> __device__ __forceinline__ void dft8(double2 (&v)[8]) {
    const double h = 0.70710678118654752440;
    double2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    double2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
    // b_r *= w8^r, w8 = exp(-+ i pi / 4)
    b1 = S > 0 ? make_double2(h * (b1.x + b1.y), h * (b1.y - b1.x)) : make_double2(h * (b1.x - b1.y), h * (b1.y + b1.x));
    b2 = mul_mi<S>(b2);
    b3 = S > 0 ? make_double2(h * (b3.y - b3.x), -h * (b3.x + b3.y)) : make_double2(-h * (b3.x + b3.y), h * (b3.x - b3.y));
    dft4<S>(a0, a1, a2, a3);
    dft4<S>(b0, b1, b2, b3);
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// one Stockham pass of radix R for butterfly j: reads in[j + r N/R], writes out[(j - k) R + k + r Ns], k = j mod Ns
template <int R, int S> __device__ __forceinline__ void pass(const double2 *in, double2 *out, int j, int Ns, const double2 *__restrict__ tw) {
    const int k = j & (Ns - 1);
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in[P(j + r * (N / R))];
    if (Ns > 1) {
        double2 w1 = tw[k * (N / (Ns * R))];                     // exp(-2 pi i k / (Ns R)); powers by multiplication
        if (S < 0) w1.y = -w1.y;
        double2 w = w1;
#pragma unroll
        for (int r = 1; r < R; ++r) { v[r] = cmul(v[r], w); if (r + 1 < R) w = cmul(w, w1); }
    }
    if constexpr (R == 8) dft8<S>(v);
    else dft4<S>(v[0], v[1], v[2], v[3]);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out[P(j0 + r * Ns)] = v[r];
}

template <int S> __device__ __forceinline__ void fft2048(double2 *&a, double2 *&b, int t, const double2 *__restrict__ tw) {
    pass<8, S>(a, b, t, 1, tw); __syncthreads();
    pass<8, S>(b, a, t, 8, tw); __syncthreads();
    pass<8, S>(a, b, t, 64, tw); __syncthreads();
    pass<4, S>(b, a, t, 512, tw); pass<4, S>(b, a, t + T, 512, tw); __syncthreads();
    // result in a
}

// y[row, 0..M) = post * IFFT(FFT(pre * x[row, 0..m), zero-padded to N) * H)
__global__ void __launch_bounds__(T, 2)
czt_rows_kernel(const double2 *__restrict__ x, int ldx, int m, const double2 *__restrict__ pre, const double2 *__restrict__ H,
                const double2 *__restrict__ post, const double2 *__restrict__ tw, double2 *__restrict__ y, int ldy, int M, int nrows) {
    extern __shared__ double2 sm[];
    double2 *a = sm, *b = sm + NP;
    const int t = threadIdx.x;
    for (int row = blockIdx.x; row < nrows; row += gridDim.x) {
#pragma unroll
        for (int i = t; i < N; i += T) a[P(i)] = i < m ? cmul(x[(long long)row * ldx + i], pre[i]) : make_double2(0.0, 0.0);
        __syncthreads();
        fft2048<1>(a, b, t, tw);
#pragma unroll
        for (int i = t; i < N; i += T) a[P(i)] = cmul(a[P(i)], H[i]);
        __syncthreads();
        fft2048<-1>(a, b, t, tw);
#pragma unroll
        for (int i = t; i < M; i += T) y[(long long)row * ldy + i] = cmul(a[P(i)], post[i]);
        __syncthreads();
    }
}

static void host_dft(const std::vector<double> &re, const std::vector<double> &im, std::vector<double> &ore, std::vector<double> &oim, int sgn) {
    const int n = (int)re.size();
    ore.assign(n, 0); oim.assign(n, 0);
    for (int k = 0; k < n; ++k) {
        long double sr = 0, si = 0;
        for (int j = 0; j < n; ++j) {
            const long long idx = ((long long)k * j) % n;
            const long double ang = -2.0L * M_PIl * sgn * idx / n;
            const long double c = cosl(ang), s = sinl(ang);
            sr += re[j] * c - im[j] * s; si += re[j] * s + im[j] * c;
        }
        ore[k] = (double)sr; oim[k] = (double)si;
    }
}

int main() {
    const int m = 1001, M = 1024, rows = 148 * 2 * 16;          // 4736 rows ~ 2.3 planes of the bench workload
    std::vector<double2> hx((size_t)rows * m), hpre(m), hH(N), hpost(M), htw(N);
    srand(1);
    for (auto &v : hx) v = make_double2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    const double alpha = 1.0 / 2048;
    for (int i = 0; i < m; ++i) { double r = i - m / 2; double ph = -M_PI * alpha * r * r; hpre[i] = make_double2(cos(ph), sin(ph)); }
    for (int i = 0; i < M; ++i) { double u = i - M / 2; double ph = -M_PI * alpha * u * u; hpost[i] = make_double2(cos(ph) / N, sin(ph) / N); }
    for (int i = 0; i < N; ++i) { long double ang = -2.0L * M_PIl * i / N; htw[i] = make_double2((double)cosl(ang), (double)sinl(ang)); }
    // chirp filter h[d] = exp(+i pi alpha d^2) for d = U - R in [-(M/2) - (m - 1 - m/2), M - 1 - M/2 + m/2], stored circularly
    std::vector<double> hr(N, 0), hi(N, 0), Hr, Hi;
    const int dmin = -(M / 2) - (m - 1 - m / 2), dmax = (M - 1 - M / 2) + m / 2;
    for (int d = dmin; d <= dmax; ++d) {
        // output index u - 0 pairs with input index r: d = (u - M/2) - (r - m/2); circular position (u - r) mod N
        const int pos = ((d + M / 2 - m / 2) % N + N) % N;
        const double ph = M_PI * alpha * (double)d * d;
        hr[pos] = cos(ph); hi[pos] = sin(ph);
    }
    host_dft(hr, hi, Hr, Hi, 1);
    for (int i = 0; i < N; ++i) hH[i] = make_double2(Hr[i], Hi[i]);

    double2 *dx, *dpre, *dH, *dpost, *dtw, *dy;
    cudaMalloc(&dx, hx.size() * 16); cudaMalloc(&dpre, m * 16); cudaMalloc(&dH, N * 16); cudaMalloc(&dpost, M * 16);
    cudaMalloc(&dtw, N * 16); cudaMalloc(&dy, (size_t)rows * M * 16);
    cudaMemcpy(dx, hx.data(), hx.size() * 16, cudaMemcpyHostToDevice); cudaMemcpy(dpre, hpre.data(), m * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(dH, hH.data(), N * 16, cudaMemcpyHostToDevice); cudaMemcpy(dpost, hpost.data(), M * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(dtw, htw.data(), N * 16, cudaMemcpyHostToDevice);
    const int smem = 2 * NP * 16;
    cudaFuncSetAttribute(czt_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, czt_rows_kernel, T, smem);
    const int grid = nsm * occ;
    for (int i = 0; i < 3; ++i) czt_rows_kernel<<<grid, T, smem>>>(dx, m, m, dpre, dH, dpost, dtw, dy, M, M, rows);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) czt_rows_kernel<<<grid, T, smem>>>(dx, m, m, dpre, dH, dpost, dtw, dy, M, M, rows);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    cudaError_t err = cudaGetLastError();
    printf("SMs %d, CTAs/SM %d, smem %d B/CTA, status %s\n", nsm, occ, smem, cudaGetErrorString(err));
    const double us_row = ms * 1e3 / rows, us_plane = us_row * (m + M);
    printf("%d rows of 2048-point FFT -> xH -> IFFT: %.3f ms = %.4f us per row; one 1001^2 -> 1024^2 plane = %d rows = %.1f us = %.0f planes/s\n",
           rows, ms, us_row, m + M, us_plane, 1e6 / us_plane);
    printf("flops (5 N log2 N per FFT, x2, + pointwise): %.2f TFLOP/s FP64\n", rows * (2 * 5.0 * N * 11 + 6.0 * (N + m + M)) / (ms * 1e-3) / 1e12);

    // check row 0 and the last row against the direct chirp-z sum  y[u] = post[u] sum_r pre[r] x[r] exp(i pi alpha (U - R)^2) = sum_r x[r] exp(-2 pi i alpha R U)
    std::vector<double2> hy((size_t)rows * M);
    cudaMemcpy(hy.data(), dy, hy.size() * 16, cudaMemcpyDeviceToHost);
    double worst = 0, peak = 0;
    for (int row : {0, rows - 1})
        for (int u = 0; u < M; u += 7) {
            long double sr = 0, si = 0;
            for (int r = 0; r < m; ++r) {
                const long double ph = -2.0L * M_PIl * alpha * (long double)(r - m / 2) * (long double)(u - M / 2);
                const long double c = cosl(ph), s = sinl(ph);
                const double2 v = hx[(size_t)row * m + r];
                sr += v.x * c - v.y * s; si += v.x * s + v.y * c;
            }
            const double2 g = hy[(size_t)row * M + u];
            worst = fmax(worst, fmax(fabs(g.x - (double)sr), fabs(g.y - (double)si)));
            peak = fmax(peak, fmax(fabs((double)sr), fabs((double)si)));
        }
    printf("max |czt - direct| = %.3e, peak %.3e -> relative %.2e\n", worst, peak, worst / peak);
    return 0;
}
