// K2a — matrix Fourier transform, complex128, FP64 tensor cores (DMMA.8x8x4) on sm_100a.
//
// Replaces lentil/fourier.py:95-101 (F = E1.f.E2 * sqrt|ar ac|) with the DFT matrices of
// lentil/fourier.py:106-121 generated on the fly IN REGISTERS, directly in the DMMA A-fragment
// layout, so E1/E2 never exist in shared memory or HBM.
//
// One kernel serves both GEMM stages.  A "stage" computes, for a data matrix D (K x C, complex,
// row-major) and the twiddle matrix Tw (R x K),
//
//        O[c][r] = scale * sum_k Tw[r][k] * D[k][c],     Tw[r][k] = exp(sgn 2 pi i alpha x_k y_r)
//        x_k = x0 + k  (input coordinate),  y_r = y0 + r  (output coordinate)
//
// i.e. it writes the product TRANSPOSED.  Stage 1 (D = f, Tw = E1) leaves T^t (n x M) in the
// workspace; stage 2 (D = T^t, Tw = E2^t) transposes back and lands F (M x N) in `out`.  The
// data operand is therefore always "K x C with C contiguous", staged with cp.async, and the
// twiddle operand is always the register-resident A fragment.
//
// Twiddles: a lane owns a fixed (row, k mod 4) slot of every 8x4 A fragment, so along K the
// twiddle it needs advances by a constant rotation exp(sgn 2 pi i alpha 4 y_r).  One complex
// multiply per DMMA k-step keeps it current; every RESEED_K elements it is re-seeded from an
// exactly range-reduced sincospi so rounding cannot accumulate (error << 1e-13, gate is 1e-10).
#include "lfd_common.cuh"

namespace lfd {

constexpr int BR = 128;      // twiddle rows (output coordinate) per CTA
constexpr int BC = 64;       // data columns per CTA
constexpr int BK = 16;       // K elements per smem stage
constexpr int STAGES = 4;
constexpr int LDS = BC + 2;  // complex elements per smem row; (LDS mod 8) == 2 => LDS.128 conflict-free
constexpr int NTHREADS = 256;
constexpr int WARPS_C = 2;   // 4 x 2 warps, each a 32 x 32 complex tile
constexpr int RESEED_TILES = 16;  // re-seed twiddles every 16 smem tiles (256 K elements)
constexpr size_t SMEM_BYTES = (size_t)STAGES * BK * LDS * sizeof(double2);

struct StageDesc {
    const double2 *D;
    long long ldd;
    double2 *O;
    long long ldo;
    int K, C, R;
    int tiles_r, tiles_c, tile_base;
    double alpha, x0, y0, scale, sgn;
};

__device__ __forceinline__ void load_tile(double2 *sd, const double2 *__restrict__ D, long long ldd,
                                          int K, int C, int k_base, int c_base, int tid) {
#pragma unroll
    for (int i = 0; i < BK * BC / NTHREADS; ++i) {
        int idx = tid + i * NTHREADS;
        int kk = idx / BC, cc = idx % BC;
        int gk = k_base + kk, gc = c_base + cc;
        bool ok = (gk < K) && (gc < C);
        const double2 *src = ok ? (D + (long long)gk * ldd + gc) : D;
        cp_async16(sd + kk * LDS + cc, src, ok);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1)
mft_stage_kernel(const StageDesc *__restrict__ descs, int count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *sD = reinterpret_cast<double2 *>(smem_raw);

    // ---- which plane / tile ----
    int tile = blockIdx.x;
    int lo = 0, hi = count - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (descs[mid].tile_base <= tile) lo = mid; else hi = mid - 1;
    }
    const StageDesc d = descs[lo];
    tile -= d.tile_base;
    const int tr = tile % d.tiles_r, tc = tile / d.tiles_r;
    const int r_base = tr * BR, c_base = tc * BC;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = warp / WARPS_C, wc = warp % WARPS_C;

    const int KT = (d.K + BK - 1) / BK;

    // ---- prologue: fill STAGES-1 stages ----
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(sD + s * BK * LDS, d.D, d.ldd, d.K, d.C, s * BK, c_base, tid);
        cp_async_commit();
    }

    double accR[4][4][2], accI[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            accR[mb][nb][0] = accR[mb][nb][1] = 0.0;
            accI[mb][nb][0] = accI[mb][nb][1] = 0.0;
        }

    // twiddle state for this lane's A-fragment slots, and the per-row rotation for k += 4
    double twr[4], twi[4], rotr[4], roti[4], yrow[4];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        yrow[mb] = d.y0 + (double)(r_base + wr * 32 + mb * 8 + g);
        cis_cycles(d.alpha, 4.0, yrow[mb], d.sgn, rotr[mb], roti[mb]);
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_tile(sD + (nk % STAGES) * BK * LDS, d.D, d.ldd, d.K, d.C, nk * BK, c_base, tid);
            cp_async_commit();
        }
        if ((kt % RESEED_TILES) == 0) {
            double xk = d.x0 + (double)(kt * BK + t);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) cis_cycles(d.alpha, xk, yrow[mb], d.sgn, twr[mb], twi[mb]);
        }
        const double2 *sd = sD + (kt % STAGES) * BK * LDS + wc * 32 + g;

#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double br[4], bi[4];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                double2 v = sd[(ks * 4 + t) * LDS + nb * 8];
                br[nb] = v.x;
                bi[nb] = v.y;
            }
            double ntwi[4];
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) ntwi[mb] = neg_f64(twi[mb]);

            // (ar + i ai)(br + i bi): four real DMMAs per 8x8 block, ordered term-major so that
            // 16 independent accumulators sit between dependent issues
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(accR[mb][nb][0], accR[mb][nb][1], twr[mb], br[nb]);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(accI[mb][nb][0], accI[mb][nb][1], twr[mb], bi[nb]);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(accR[mb][nb][0], accR[mb][nb][1], ntwi[mb], bi[nb]);
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) dmma884(accI[mb][nb][0], accI[mb][nb][1], twi[mb], br[nb]);

            // advance the twiddles by 4 along K
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                double nr = twr[mb] * rotr[mb] - twi[mb] * roti[mb];
                double ni = twr[mb] * roti[mb] + twi[mb] * rotr[mb];
                twr[mb] = nr;
                twi[mb] = ni;
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: transposed store, 8 consecutive rows (128 B) per quarter-warp ----
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        int r = r_base + wr * 32 + mb * 8 + g;
        if (r >= d.R) continue;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                int c = c_base + wc * 32 + nb * 8 + 2 * t + i;
                if (c < d.C)
                    d.O[(long long)c * d.ldo + r] =
                        make_double2(accR[mb][nb][i] * d.scale, accI[mb][nb][i] * d.scale);
            }
    }
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// mft_folded.cu
size_t folded_workspace_bytes(const lfd_mft_desc *descs, int count);
int launch_mft_folded(const lfd_mft_desc *descs, int count, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream, const lfd_pupil_src *src, int intensity_out);

bool czt_supported(const lfd_mft_desc *descs, int count);
bool czt_preferred(const lfd_mft_desc *descs, int count);
size_t czt_workspace_bytes(const lfd_mft_desc *descs, int count, bool c64);
int launch_mft_czt(const lfd_mft_desc *descs, int count, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                   const lfd_pupil_src *src, int intensity_out, bool c64);

static std::atomic<int> g_variant{LFD_MFT_AUTO};

// the execution a batch runs under the current process-wide setting
int mft_resolve_execution(const lfd_mft_desc *descs, int count) {
    // per-call choice (lfd_mft_desc.execution of the first plane) before the process default
    const int forced = (descs && count > 0) ? descs[0].execution : 0;
    const int v = (forced >= 1 && forced <= 1 + LFD_MFT_AUTO) ? forced - 1 : g_variant.load();
    if (v == LFD_MFT_CZT) return czt_supported(descs, count) ? LFD_MFT_CZT : LFD_MFT_FOLDED;
    if (v == LFD_MFT_AUTO) return czt_preferred(descs, count) ? LFD_MFT_CZT : LFD_MFT_FOLDED;
    return v;
}

}  // namespace lfd

using namespace lfd;

extern "C" int lfd_set_mft_variant(int variant) {
    LFD_REQUIRE(variant >= LFD_MFT_DIRECT && variant <= LFD_MFT_AUTO, "unknown MFT variant %d", variant);
    g_variant.store(variant);
    return 0;
}
extern "C" int lfd_get_mft_variant(void) { return g_variant.load(); }
extern "C" int lfd_mft_execution(const lfd_mft_desc *descs, int count) {
    return (descs && count > 0) ? mft_resolve_execution(descs, count) : g_variant.load();
}

extern "C" size_t lfd_mft_workspace_bytes(const lfd_mft_desc *descs, int count) {
    const int variant = mft_resolve_execution(descs, count);
    if (variant == LFD_MFT_CZT) return czt_workspace_bytes(descs, count, false);
    if (variant == LFD_MFT_FOLDED) return folded_workspace_bytes(descs, count);
    size_t bytes = align_up((size_t)2 * count * sizeof(StageDesc), 256);
    for (int i = 0; i < count; ++i)
        bytes += align_up((size_t)descs[i].n * descs[i].M * sizeof(double2), 256);
    return bytes;
}

extern "C" int lfd_mft_c128_batched(const lfd_mft_desc *descs, int count, void *workspace,
                                    size_t workspace_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (count == 0) return 0;
    LFD_REQUIRE(descs != nullptr && count > 0, "lfd_mft_c128_batched: bad descriptor array");
    LFD_REQUIRE(workspace != nullptr, "lfd_mft_c128_batched: workspace is NULL");
    const int variant = mft_resolve_execution(descs, count);
    if (variant == LFD_MFT_CZT)
        return launch_mft_czt(descs, count, workspace, workspace_bytes, stream, nullptr, 0, false);
    if (variant != LFD_MFT_DIRECT)
        return launch_mft_folded(descs, count, workspace, workspace_bytes, stream, nullptr, 0);
    size_t need = lfd_mft_workspace_bytes(descs, count);
    LFD_REQUIRE(workspace_bytes >= need, "lfd_mft_c128_batched: workspace too small (%zu < %zu)",
                workspace_bytes, need);

    int dev = 0;
    LFD_CUDA_OK(cudaGetDevice(&dev));
    if (ensure_dynamic_smem(dev, (const void *)mft_stage_kernel, (int)SMEM_BYTES)) return 1;

    StageDesc *h = (StageDesc *)malloc((size_t)2 * count * sizeof(StageDesc));
    LFD_REQUIRE(h != nullptr, "out of host memory");
    char *ws = (char *)workspace;
    size_t off = align_up((size_t)2 * count * sizeof(StageDesc), 256);
    int tiles1 = 0, tiles2 = 0;
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        if (!(p.m > 0 && p.n > 0 && p.M > 0 && p.N > 0 && p.ldf >= p.n && p.ldo >= p.N && p.f && p.out)) {
            free(h);
            LFD_REQUIRE(false, "lfd_mft_c128_batched: plane %d has invalid shape/ld/pointers", i);
        }
        double2 *Tt = (double2 *)(ws + off);
        off += align_up((size_t)p.n * p.M * sizeof(double2), 256);
        double sgn = p.inverse ? 1.0 : -1.0;
        double scale = p.unitary ? sqrt(fabs(p.alpha_r * p.alpha_c)) : 1.0;
        if (p.inverse) scale /= ((double)p.m * (double)p.n);

        StageDesc &s1 = h[i];
        s1.D = (const double2 *)p.f; s1.ldd = p.ldf;
        s1.O = Tt; s1.ldo = p.M;
        s1.K = p.m; s1.C = p.n; s1.R = p.M;
        s1.alpha = p.alpha_r;
        s1.x0 = -floor(p.m / 2.0) + p.off_r;
        s1.y0 = -floor(p.M / 2.0) - p.shift_r;
        s1.scale = 1.0; s1.sgn = sgn;
        s1.tiles_r = (s1.R + BR - 1) / BR; s1.tiles_c = (s1.C + BC - 1) / BC;
        s1.tile_base = tiles1; tiles1 += s1.tiles_r * s1.tiles_c;

        StageDesc &s2 = h[count + i];
        s2.D = Tt; s2.ldd = p.M;
        s2.O = (double2 *)p.out; s2.ldo = p.ldo;
        s2.K = p.n; s2.C = p.M; s2.R = p.N;
        s2.alpha = p.alpha_c;
        s2.x0 = -floor(p.n / 2.0) + p.off_c;
        s2.y0 = -floor(p.N / 2.0) - p.shift_c;
        s2.scale = scale; s2.sgn = sgn;
        s2.tiles_r = (s2.R + BR - 1) / BR; s2.tiles_c = (s2.C + BC - 1) / BC;
        s2.tile_base = tiles2; tiles2 += s2.tiles_r * s2.tiles_c;
    }
    cudaError_t e = cudaMemcpyAsync(workspace, h, (size_t)2 * count * sizeof(StageDesc),
                                    cudaMemcpyHostToDevice, stream);
    free(h);  // pageable source: the copy has been staged by the time the call returns
    LFD_CUDA_OK(e);

    const StageDesc *dd = (const StageDesc *)workspace;
    mft_stage_kernel<<<tiles1, NTHREADS, SMEM_BYTES, stream>>>(dd, count);
    LFD_CUDA_OK(cudaGetLastError());
    mft_stage_kernel<<<tiles2, NTHREADS, SMEM_BYTES, stream>>>(dd + count, count);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch(2);
    return 0;
}

extern "C" int lfd_mft_c128(const lfd_mft_desc *desc, void *workspace, size_t workspace_bytes,
                            void *stream) {
    return lfd_mft_c128_batched(desc, 1, workspace, workspace_bytes, stream);
}
