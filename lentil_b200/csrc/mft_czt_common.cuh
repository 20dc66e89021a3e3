// Shared by the two translation units of the chirp-z execution (mft_czt.cu: FP64 build, tables, host side; mft_czt_f32.cu: FP32
// build): pass plan, tuning macros, CTA shapes, the type-erased plane descriptor and the stage-launch helper.
#pragma once
#include "lfd_common.cuh"
#include <type_traits>

namespace lfd {
namespace czt {

constexpr int MAX_LOG2L = 13, MIN_LOG2L = 6;
// pass plan of a length: nreg radix-16 passes (the first of them has no twiddles), then the turn of radix 2 / 4 / 8 / 16
__host__ __device__ constexpr int nreg(int lg) { return (lg - 1) / 4; }
__host__ __device__ constexpr int turn_radix(int lg) { return 1 << (lg - 4 * nreg(lg)); }
// Pass twiddles, one contiguous run per (FFT length, pass): pass p >= 1 has sub-transform length Ns = 16^p and reads
// g_tw[lg][tw_offset(p) + k] = exp(-2 pi i k / (Ns R)), k = 0 .. Ns - 1 (R = 16, or the turn's radix for p = nreg).
// Consecutive butterflies read consecutive entries.
__host__ __device__ constexpr int tw_offset(int p) { return ((1 << (4 * p)) - 16) / 15; }
constexpr int TW_PER_LEN = 4400;                 // 16 + 256 + 4096 = 4368 entries for the longest transform
// rows per CTA and buffers per row, by length (measured, r02: profiles/r02_czt_variants.md): 128-thread CTAs up to
// L = 1024 (ROWS * L = 2048), two rows (256 threads) at L = 2048, one row beyond.  LFD_CZT_ELEMS overrides ROWS * L.
#ifndef LFD_CZT_ELEMS
#define LFD_CZT_ELEMS 0
#endif
#ifndef LFD_CZT_NBUF
#define LFD_CZT_NBUF 1
#endif
// tables fetched before the barrier that precedes their pass (registers are free there, but only so many of them)
#ifndef LFD_CZT_HPRE
#define LFD_CZT_HPRE 8        // how many of a thread's 16 values of H are fetched before the turn's barrier (0, 4, 8 or 16)
#endif
#ifndef LFD_CZT_PRE_MAXLG
#define LFD_CZT_PRE_MAXLG 11  // ... all three prefetches only for transforms up to this length: beyond it (radix-16 turn, 512-thread
#endif                        // CTAs) they cost registers or L2 bandwidth and measured 3 % slower (r02, profiles/r02_czt_variants.md)
#ifndef LFD_CZT_L2PRE
#define LFD_CZT_L2PRE 1       // prefetch.global.L2 of the next unit's input rows
#endif
#ifndef LFD_CZT_CONTIG
#define LFD_CZT_CONTIG 0      // 1: one contiguous run of work units per CTA instead of the round-robin deal (see czt_stage_kernel)
#endif
#ifndef LFD_CZT_F64_THREADS
#define LFD_CZT_F64_THREADS 512
#endif
#ifndef LFD_CZT_F32_THREADS
#define LFD_CZT_F32_THREADS 768
#endif
#ifndef LFD_CZT_STAGING
#define LFD_CZT_STAGING 1     // next unit's input rows copied into shared memory by the bulk-copy engine while this unit computes
#endif
#ifndef LFD_CZT_STAGING_MINLG
#define LFD_CZT_STAGING_MINLG 11   // ... for transforms of at least this length (measured r02: +1 % at 2048, +5 % at 4096, -15 % at 1024)
#endif
#ifndef LFD_CZT_PPRE
#define LFD_CZT_PPRE 4        // post-chirp factors prefetched before the last pass (0 .. 8)
#endif
__host__ __device__ constexpr int rows_for(int lg) {
    return LFD_CZT_ELEMS ? ((1 << lg) >= LFD_CZT_ELEMS ? 1 : LFD_CZT_ELEMS >> lg) : (lg <= 10 ? 2048 >> lg : (lg == 11 ? 2 : 1));
}
__host__ __device__ constexpr int cta_threads(int lg) { return ((1 << lg) / 16) * rows_for(lg); }
__host__ __device__ constexpr int nbuf_for(int lg) { return ((size_t)LFD_CZT_NBUF * rows_for(lg) * ((1 << lg) + (1 << lg) / 16) * 16 > 200 * 1024) ? 1 : LFD_CZT_NBUF; }

// complex arrays are complex128 in the FP64 build and complex64 in the FP32 build (type-erased here)
struct Plane {
    const void *f; long long ldf;
    void *Gt;                     // stage A result, transposed: N x mpad
    void *out; long long ldo;
    int m, n, M, N, mpad, logLA, logLB, intensity;
    void *preA, *postA, *HA;      // axis 1 (n -> N): pre[n], post[N], H[LA]
    void *preB, *postB, *HB;      // axis 0 (m -> M)
    double alpha_r, alpha_c, x0r, y0r, x0c, y0c, sgn, scale;
    // fused pupil prep: when amp != NULL, f(i, c) = amp * mask * exp(+2 pi i opd / lambda) at pupil pixel (pr0 + i, pc0 + c)
    const double *amp, *opd;
    const unsigned char *mask;
    long long pld;
    int pr0, pc0;
    double wavelength, inv_wavelength;
};

// fill one length's pass twiddles (double precision values; OUT = double2 or float2)
template <class OUT>
__device__ __forceinline__ void fill_roots(OUT (*table)[TW_PER_LEN]) {
    const int lg = MIN_LOG2L + blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int np = nreg(lg);
    for (int p = 1; p <= np; ++p) {
        const int Ns = 1 << (4 * p), off = tw_offset(p);
        if (e >= off && e < off + Ns) {
            const int k = e - off;
            const double den = p < np ? 16.0 * Ns : (double)(1 << lg);
            double s, c;
            sincospi(-2.0 * (double)k / den, &s, &c);      // k / den is exact (power-of-two denominator)
            table[blockIdx.y][e].x = (decltype(table[0][0].x))c;
            table[blockIdx.y][e].y = (decltype(table[0][0].y))s;
        }
    }
}

// launch one stage kernel on as many CTAs as fit the device (or as there are work units)
template <class K>
static int launch_stage(K kernel, int threads, int smem, int total, int dev, int nsm, const Plane *dd, const int *starts, int count,
                        cudaStream_t stream) {
    int occ = 1;
    if (ensure_dynamic_smem(dev, (const void *)kernel, smem)) return 1;
    LFD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
    LFD_REQUIRE(occ > 0, "lfd_mft (chirp-z): a stage kernel does not fit an SM (%d threads, %d bytes of shared memory)", threads, smem);
    const int grid = total < nsm * occ ? total : nsm * occ;
    kernel<<<grid, threads, smem, stream>>>(dd, starts, count);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

// bytes of dynamic shared memory of a stage kernel: the FFT buffers and, from LFD_CZT_STAGING_MINLG up, the input staging buffer
inline int stage_smem_bytes(int lg, size_t elem_bytes) {
    const int Lr = 1 << lg;
    return (int)(nbuf_for(lg) * rows_for(lg) * (Lr + Lr / 16) * elem_bytes) +
           ((LFD_CZT_STAGING && lg >= LFD_CZT_STAGING_MINLG) ? rows_for(lg) * (Lr / 2 + 4) * 16 : 0);
}

// the FP32 build lives in mft_czt_f32.cu
int czt_f32_ensure_roots(int dev, cudaStream_t stream);
int czt_f32_launch_stage(int lg, bool stage_a, int total, int dev, int nsm, const Plane *dd, const int *starts, int count, cudaStream_t stream);

}  // namespace czt
}  // namespace lfd
