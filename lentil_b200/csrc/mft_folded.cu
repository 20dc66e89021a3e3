// K2a, folded variant — the same matrix Fourier transform with 4x fewer FP64 MMAs.
//
// Each 1-D stage  T[U] = sum_R exp(sgn 2 pi i a (R+o)(U-s)) f[R]  (lentil/fourier.py:106-110) is
// rewritten on grids that are symmetric about zero.  With R' = R + cR and U' = U + cU, where
// cR (cU) is 1/2 for an even input (output) length and 0 for an odd one, R' and U' run over
// +-(r + cR), +-(u + cU) and
//
//     (R+o)(U-s) = R'U' - s'R' + o'(U'-s'),        o' = o - cR,  s' = s + cU
//
//     exp(sgn 2 pi i a (R+o)(U-s)) = [cos th + i sgn sin th] * pre(R') * post(U'),   th = 2 pi a R'U'
//     pre(R')  = exp(-sgn 2 pi i a s' R')      (a phase ramp on the input rows:  fold kernel)
//     post(U') = exp( sgn 2 pi i a o'(U'-s'))  (a phase ramp on the output rows: MFT epilogue)
//
// cos th is even and sin th odd in both R' and U', so with g = pre * f folded into
//     ge[r] = g[+R'_r] + g[-R'_r],   go[r] = g[+R'_r] - g[-R'_r]            (r = 0 .. ceil(K/2)-1)
//     A[u] = sum_r cos(th_ru) ge[r],  B[u] = sum_r sin(th_ru) go[r]          (u = 0 .. ceil(M/2)-1)
//     T[+U'_u] = post(+U'_u) (A[u] + i sgn B[u]),   T[-U'_u] = post(-U'_u) (A[u] - i sgn B[u])
//
// the twiddle operand is REAL and half as tall, the K extent is half as long: two real x complex
// GEMMs of (M/2 x K/2) instead of one complex x complex GEMM of (M x K) — M*K*n real MACs per
// stage instead of 4*M*K*n.  Exact in exact arithmetic, for any alpha / shift / offset / parity;
// rounding differs from the direct form by a few ulp (parity gate 1e-10, observed ~1e-14).
//
// Kernel structure is that of mft_c128.cu: cos/sin A-fragments live in registers and advance by
// a per-row rotation each DMMA k-step, ge/go tiles are staged with cp.async, the product is
// stored transposed so that the next stage again sees "K x C, C contiguous".
#include "lfd_common.cuh"

namespace lfd {

constexpr int FBR = 128;     // folded output rows per CTA
constexpr int FBC = 32;      // complex data columns per CTA
constexpr int FBK = 16;      // folded K rows per smem stage
constexpr int FSTAGES = 4;
constexpr int FLDS = FBC + 2;  // complex elements per smem row: LDS.128 conflict-free (see mft_c128.cu)
constexpr int FTHREADS = 256;
constexpr int FWARPS_C = 2;  // 4 x 2 warps, each 32 folded rows x 16 complex columns
constexpr int FRESEED_TILES = 16;
constexpr size_t FSMEM_BYTES = (size_t)FSTAGES * 2 * FBK * FLDS * sizeof(double2);

struct FoldDesc {
    const double2 *D;   // K x C
    long long ldd;
    double2 *G;         // ge plane (Kf x C, ld = C) followed by go plane
    int K, C, Kf, hm, cR2;
    int row_base;
    double alpha, sprime, sgn;
};

struct FStageDesc {
    const double2 *G;   // ge at G, go at G + Kf*C
    double2 *O;
    long long ldo;
    int Kf, C, Rf, M, hM, cR2, cU2;
    int tiles_r, tiles_c, tile_base;
    double alpha, oprime, sprime, scale, sgn;
};

// ---- fold: g = pre * f, ge/go = g[+R'] +- g[-R'] -------------------------------------------------
__global__ void __launch_bounds__(256)
fold_kernel(const FoldDesc *__restrict__ descs, int count) {
    int row = blockIdx.x;
    int lo = 0, hi = count - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (descs[mid].row_base <= row) lo = mid; else hi = mid - 1;
    }
    const FoldDesc d = descs[lo];
    const int r = row - d.row_base;
    const double Rp = (double)r + 0.5 * d.cR2;
    __shared__ double pre[2];
    if (threadIdx.x == 0) {
        double c, s;
        cis_cycles(d.alpha, d.sprime, Rp, -d.sgn, c, s);
        pre[0] = c;
        pre[1] = s;
    }
    __syncthreads();
    const double pc = pre[0], ps = pre[1];
    const bool center = (d.cR2 == 0) && (r == 0);
    const int ip = d.hm + r, im = d.hm - r - d.cR2;
    const double2 *__restrict__ rowp = d.D + (long long)ip * d.ldd;
    const double2 *__restrict__ rowm = d.D + (long long)im * d.ldd;
    double2 *__restrict__ ge = d.G + (long long)r * d.C;
    double2 *__restrict__ go = d.G + ((long long)d.Kf + r) * d.C;
    const bool has_p = ip < d.K;
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
        double2 a = has_p ? rowp[c] : make_double2(0.0, 0.0);
        double2 gp = make_double2(a.x * pc - a.y * ps, a.x * ps + a.y * pc);
        if (center) {
            ge[c] = gp;
            go[c] = make_double2(0.0, 0.0);
        } else {
            double2 b = rowm[c];
            double2 gm = make_double2(b.x * pc + b.y * ps, b.y * pc - b.x * ps);   // conj(pre) * b
            ge[c] = make_double2(gp.x + gm.x, gp.y + gm.y);
            go[c] = make_double2(gp.x - gm.x, gp.y - gm.y);
        }
    }
}

// ---- folded MFT stage ---------------------------------------------------------------------------
__device__ __forceinline__ void f_load_tile(double2 *sd, const FStageDesc &d, int k_base, int c_base,
                                            int tid) {
#pragma unroll
    for (int i = 0; i < 2 * FBK * FBC / FTHREADS; ++i) {
        int idx = tid + i * FTHREADS;
        int p = idx / (FBK * FBC), rem = idx % (FBK * FBC);
        int kk = rem / FBC, cc = rem % FBC;
        int gk = k_base + kk, gc = c_base + cc;
        bool ok = (gk < d.Kf) && (gc < d.C);
        const double2 *src = ok ? (d.G + ((long long)p * d.Kf + gk) * d.C + gc) : d.G;
        cp_async16(sd + (p * FBK + kk) * FLDS + cc, src, ok);
    }
}

__global__ void __launch_bounds__(FTHREADS, 1)
mft_folded_kernel(const FStageDesc *__restrict__ descs, int count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *sD = reinterpret_cast<double2 *>(smem_raw);
    constexpr int STAGE_ELEMS = 2 * FBK * FLDS;

    int tile = blockIdx.x;
    int lo = 0, hi = count - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (descs[mid].tile_base <= tile) lo = mid; else hi = mid - 1;
    }
    const FStageDesc d = descs[lo];
    tile -= d.tile_base;
    const int tr = tile % d.tiles_r, tc = tile / d.tiles_r;
    const int r_base = tr * FBR, c_base = tc * FBC;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = warp / FWARPS_C, wc = warp % FWARPS_C;
    const int KT = (d.Kf + FBK - 1) / FBK;

#pragma unroll
    for (int s = 0; s < FSTAGES - 1; ++s) {
        if (s < KT) f_load_tile(sD + s * STAGE_ELEMS, d, s * FBK, c_base, tid);
        cp_async_commit();
    }

    // accA[mb][q][part]: A = cos-GEMM on ge; accB: B = sin-GEMM on go.  part 0 = Re columns, 1 = Im.
    double accA[4][2][2][2], accB[4][2][2][2];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                accA[mb][q][p][0] = accA[mb][q][p][1] = 0.0;
                accB[mb][q][p][0] = accB[mb][q][p][1] = 0.0;
            }

    double tc_[4], ts_[4], rc_[4], rs_[4], up[4];
    const double cR = 0.5 * d.cR2, cU = 0.5 * d.cU2;
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        up[mb] = (double)(r_base + wr * 32 + mb * 8 + g) + cU;          // U' of this lane's row
        cis_cycles(d.alpha, 4.0, up[mb], 1.0, rc_[mb], rs_[mb]);
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<FSTAGES - 2>();
        __syncthreads();
        {
            int nk = kt + FSTAGES - 1;
            if (nk < KT) f_load_tile(sD + (nk % FSTAGES) * STAGE_ELEMS, d, nk * FBK, c_base, tid);
            cp_async_commit();
        }
        if ((kt % FRESEED_TILES) == 0) {
            double rp = (double)(kt * FBK + t) + cR;                     // R' of this lane's K slot
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) cis_cycles(d.alpha, rp, up[mb], 1.0, tc_[mb], ts_[mb]);
        }
        const double2 *se = sD + (kt % FSTAGES) * STAGE_ELEMS + wc * 16 + g;
        const double2 *so = se + FBK * FLDS;

#pragma unroll
        for (int ks = 0; ks < FBK / 4; ++ks) {
            double2 ve[2], vo[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                ve[q] = se[(ks * 4 + t) * FLDS + q * 8];
                vo[q] = so[(ks * 4 + t) * FLDS + q * 8];
            }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    dmma884(accA[mb][q][0][0], accA[mb][q][0][1], tc_[mb], ve[q].x);
                    dmma884(accA[mb][q][1][0], accA[mb][q][1][1], tc_[mb], ve[q].y);
                }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb)
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    dmma884(accB[mb][q][0][0], accB[mb][q][0][1], ts_[mb], vo[q].x);
                    dmma884(accB[mb][q][1][0], accB[mb][q][1][1], ts_[mb], vo[q].y);
                }
#pragma unroll
            for (int mb = 0; mb < 4; ++mb) {
                double nc = tc_[mb] * rc_[mb] - ts_[mb] * rs_[mb];
                double ns = tc_[mb] * rs_[mb] + ts_[mb] * rc_[mb];
                tc_[mb] = nc;
                ts_[mb] = ns;
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: unfold to the +U' and -U' output rows, post phase, scale, transposed store ----
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
        const int u = r_base + wr * 32 + mb * 8 + g;
        if (u >= d.Rf) continue;
        const int kp = d.hM + u, km = d.hM - u - d.cU2;
        const bool has_p = kp < d.M;
        const bool has_m = (km >= 0) && !(d.cU2 == 0 && u == 0);
        double ppc, pps, pmc, pms;
        cis_cycles(d.alpha, d.oprime, up[mb] - d.sprime, d.sgn, ppc, pps);
        cis_cycles(d.alpha, d.oprime, -up[mb] - d.sprime, d.sgn, pmc, pms);
        ppc *= d.scale; pps *= d.scale; pmc *= d.scale; pms *= d.scale;
#pragma unroll
        for (int q = 0; q < 2; ++q)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c = c_base + wc * 16 + q * 8 + 2 * t + i;
                if (c >= d.C) continue;
                const double Ar = accA[mb][q][0][i], Ai = accA[mb][q][1][i];
                const double Br = d.sgn * accB[mb][q][0][i], Bi = d.sgn * accB[mb][q][1][i];
                double2 *col = d.O + (long long)c * d.ldo;
                if (has_p) {   // A + i sgn B
                    double xr = Ar - Bi, xi = Ai + Br;
                    col[kp] = make_double2(xr * ppc - xi * pps, xr * pps + xi * ppc);
                }
                if (has_m) {   // A - i sgn B
                    double xr = Ar + Bi, xi = Ai - Br;
                    col[km] = make_double2(xr * pmc - xi * pms, xr * pms + xi * pmc);
                }
            }
    }
}

static inline size_t f_align(size_t v, size_t a) { return (v + a - 1) / a * a; }

size_t folded_workspace_bytes(const lfd_mft_desc *descs, int count) {
    size_t bytes = f_align((size_t)2 * count * (sizeof(FoldDesc) + sizeof(FStageDesc)), 256);
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        size_t g1 = (size_t)2 * ((p.m + 1) / 2) * p.n, g2 = (size_t)2 * ((p.n + 1) / 2) * p.M;
        bytes += f_align((g1 > g2 ? g1 : g2) * sizeof(double2), 256);
        bytes += f_align((size_t)p.n * p.M * sizeof(double2), 256);
    }
    return bytes;
}

int launch_mft_folded(const lfd_mft_desc *descs, int count, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
    size_t need = folded_workspace_bytes(descs, count);
    LFD_REQUIRE(workspace_bytes >= need, "lfd_mft_c128_batched: workspace too small (%zu < %zu)",
                workspace_bytes, need);
    static bool attr_set = false;
    if (!attr_set) {
        LFD_CUDA_OK(cudaFuncSetAttribute(mft_folded_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FSMEM_BYTES));
        attr_set = true;
    }
    const size_t nfold = (size_t)2 * count, nstage = (size_t)2 * count;
    const size_t hdr_bytes = nfold * sizeof(FoldDesc) + nstage * sizeof(FStageDesc);
    char *h = (char *)malloc(hdr_bytes);
    LFD_REQUIRE(h != nullptr, "out of host memory");
    FoldDesc *hf = (FoldDesc *)h;
    FStageDesc *hs = (FStageDesc *)(h + nfold * sizeof(FoldDesc));
    char *ws = (char *)workspace;
    size_t off = f_align(hdr_bytes, 256);
    int rows1 = 0, rows2 = 0, tiles1 = 0, tiles2 = 0;
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        if (!(p.m > 0 && p.n > 0 && p.M > 0 && p.N > 0 && p.ldf >= p.n && p.ldo >= p.N && p.f && p.out)) {
            free(h);
            LFD_REQUIRE(false, "lfd_mft_c128_batched: plane %d has invalid shape/ld/pointers", i);
        }
        size_t g1 = (size_t)2 * ((p.m + 1) / 2) * p.n, g2 = (size_t)2 * ((p.n + 1) / 2) * p.M;
        double2 *G = (double2 *)(ws + off);
        off += f_align((g1 > g2 ? g1 : g2) * sizeof(double2), 256);
        double2 *Tt = (double2 *)(ws + off);
        off += f_align((size_t)p.n * p.M * sizeof(double2), 256);
        const double sgn = p.inverse ? 1.0 : -1.0;
        double scale = p.unitary ? sqrt(fabs(p.alpha_r * p.alpha_c)) : 1.0;
        if (p.inverse) scale /= ((double)p.m * (double)p.n);

        for (int st = 0; st < 2; ++st) {
            // stage 0: rows (K = m, C = n, out rows M); stage 1: columns (K = n, C = M, out rows N)
            const int K = st == 0 ? p.m : p.n, C = st == 0 ? p.n : p.M, Mo = st == 0 ? p.M : p.N;
            const double alpha = st == 0 ? p.alpha_r : p.alpha_c;
            const double o = st == 0 ? p.off_r : p.off_c, s = st == 0 ? p.shift_r : p.shift_c;
            const int cR2 = (K % 2 == 0) ? 1 : 0, cU2 = (Mo % 2 == 0) ? 1 : 0;
            FoldDesc &fd = hf[st * count + i];
            fd.D = st == 0 ? (const double2 *)p.f : Tt;
            fd.ldd = st == 0 ? p.ldf : p.M;
            fd.G = G;
            fd.K = K; fd.C = C; fd.Kf = (K + 1) / 2; fd.hm = K / 2; fd.cR2 = cR2;
            fd.alpha = alpha; fd.sprime = s + 0.5 * cU2; fd.sgn = sgn;
            fd.row_base = st == 0 ? rows1 : rows2;
            (st == 0 ? rows1 : rows2) += fd.Kf;

            FStageDesc &sd = hs[st * count + i];
            sd.G = G;
            sd.O = st == 0 ? Tt : (double2 *)p.out;
            sd.ldo = st == 0 ? p.M : p.ldo;
            sd.Kf = fd.Kf; sd.C = C; sd.Rf = (Mo + 1) / 2; sd.M = Mo; sd.hM = Mo / 2;
            sd.cR2 = cR2; sd.cU2 = cU2;
            sd.alpha = alpha; sd.oprime = o - 0.5 * cR2; sd.sprime = s + 0.5 * cU2;
            sd.scale = st == 0 ? 1.0 : scale; sd.sgn = sgn;
            sd.tiles_r = (sd.Rf + FBR - 1) / FBR; sd.tiles_c = (C + FBC - 1) / FBC;
            sd.tile_base = st == 0 ? tiles1 : tiles2;
            (st == 0 ? tiles1 : tiles2) += sd.tiles_r * sd.tiles_c;
        }
    }
    cudaError_t e = cudaMemcpyAsync(workspace, h, hdr_bytes, cudaMemcpyHostToDevice, stream);
    free(h);
    LFD_CUDA_OK(e);
    const FoldDesc *df = (const FoldDesc *)workspace;
    const FStageDesc *ds = (const FStageDesc *)((char *)workspace + nfold * sizeof(FoldDesc));
    fold_kernel<<<rows1, 256, 0, stream>>>(df, count);
    LFD_CUDA_OK(cudaGetLastError());
    mft_folded_kernel<<<tiles1, FTHREADS, FSMEM_BYTES, stream>>>(ds, count);
    LFD_CUDA_OK(cudaGetLastError());
    fold_kernel<<<rows2, 256, 0, stream>>>(df + count, count);
    LFD_CUDA_OK(cudaGetLastError());
    mft_folded_kernel<<<tiles2, FTHREADS, FSMEM_BYTES, stream>>>(ds + count, count);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch(4);
    return 0;
}

}  // namespace lfd
