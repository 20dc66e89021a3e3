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
//     pre(R')  = exp(-sgn 2 pi i a s' R')      (a phase ramp on the input rows)
//     post(U') = exp( sgn 2 pi i a o'(U'-s'))  (a phase ramp on the output rows)
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
// Launch sequence for a batch:   fold_kernel (f -> ge1/go1)
//                                mft_folded_kernel<true>   rows:    ge1/go1 -> ge2/go2
//                                mft_folded_kernel<false>  columns: ge2/go2 -> F
// The row stage writes its result ALREADY FOLDED for the column stage: a CTA owns 16 folded column
// indices = 16 columns and their 16 mirror images, laid out in shared memory so that a lane's
// accumulators hold T[., j+] and T[., j-] side by side; its epilogue forms pre*T+ +- conj(pre)*T-
// and stores them transposed.  The intermediate T never exists in HBM.
//
// Kernel structure otherwise follows mft_c128.cu: cos/sin A-fragments live in registers and
// advance by a per-row rotation each DMMA k-step.
//
// Operand layout in HBM ("blocked"): ge/go are not stored as matrices but pre-tiled for the consumer,
//     [column tile][K tile][plane ge|go][16 K rows][32 columns]   (16 KB per (column tile, K tile))
// with the column index XOR-ed by 2*(row & 3), so that ONE cp.async.bulk per K tile lands a block in
// shared memory exactly as the DMMA B-fragment loads want it (LDS.128, conflict-free, no padding).
// Writers (fold_kernel, the row stage's epilogue) pay the permutation; the MMA warps issue no copies.
#include "lfd_common.cuh"

namespace lfd {
int mft_resolve_execution(const lfd_mft_desc *descs, int count);
int launch_mft_czt(const lfd_mft_desc *descs, int count, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                   const lfd_pupil_src *src, int intensity_out, bool c64);
}
namespace lfd {

constexpr int FBR = 64;      // folded output rows per CTA
constexpr int FBC = 32;      // complex data columns per CTA
constexpr int FBK = 16;      // folded K rows per smem stage
constexpr int FSTAGES = 4;
constexpr int FTHREADS = 128;  // 4 warps stacked along the rows, each 16 folded rows x 32 complex columns
                               // (a twiddle element feeds 8 DMMAs); two CTAs per SM so that one CTA's
                               // prologue / epilogue hides under the other's MMAs
constexpr int FRESEED_TILES = 64;  // re-seed the twiddle recurrence every 1024 folded K (2048 input rows)
constexpr int FOLD_MAX_TILES = 1024;  // folded-column tiles per plane the support map can hold (inputs up to 32768 columns)
constexpr int FBLOCK_ELEMS = 2 * FBK * FBC;                      // complex elements per (column tile, K tile) block
constexpr unsigned FBLOCK_BYTES = FBLOCK_ELEMS * sizeof(double2);   // 16 KB
constexpr size_t FSMEM_BYTES = (size_t)FSTAGES * FBLOCK_BYTES;

// element (K row r, plane p, slot sl) of column tile tl in a blocked operand with Kt K tiles
__host__ __device__ __forceinline__ size_t blk_index(int Kt, int tl, int r, int p, int sl) {
    return ((((size_t)tl * Kt + (r >> 4)) * 2 + p) * FBK + (r & 15)) * FBC + (sl ^ ((r & 3) << 1));
}

struct FoldDesc {
    const double2 *D;   // K x C
    long long ldd;
    double2 *G;         // blocked ge/go operand of the row stage (see blk_index)
    int K, C, Kf, hm, cR2, Kt;   // Kt = K tiles (rows padded to 16 are written as zeros)
    double alpha, sprime, sgn;
    // support map for the row stage: kmax[t] = 1 + last folded row with data in folded-column tile t,
    // kmax[ntile + t] = Kf - first such row (both start at 0 = empty; columns c and their mirrors share a
    // tile: r2 = c - nhm  or  nhm - c - ncR2)
    int *kmax;
    int nhm, ncR2, ntile, pad2_;
    // fused pupil prep (K1 inside the fold): when amp != NULL the input element (i, c) is not read from D
    // but formed as amp * mask * exp(+2 pi i opd / lambda) at pupil pixel (pr0 + i, pc0 + c), lentil/plane.py:502-507
    const double *amp, *opd;
    const unsigned char *mask;   // this segment's mask plane, or NULL
    long long pld;               // pupil row stride (n_c)
    int pr0, pc0;
    double wavelength, inv_wavelength;
};

__device__ __forceinline__ double2 pupil_phasor(const FoldDesc &d, int i, int c) {
    const long long pix = (long long)(d.pr0 + i) * d.pld + (d.pc0 + c);
    double a = d.amp[pix];
    if (d.mask != nullptr && d.mask[pix] == 0) a = 0.0;
    if (a == 0.0) return make_double2(0.0, 0.0);
    const double tcyc = d.opd[pix] * d.inv_wavelength;  // phase in cycles (multiplication by the rounded reciprocal, as in K1), reduced exactly
    double sn, cs;
    cis_unit(tcyc - rint(tcyc), cs, sn);
    return make_double2(a * cs, a * sn);
}

struct FStageDesc {
    const double2 *G;   // blocked ge/go operand (blk_index, Ktiles K tiles)
    double2 *O;         // FOLD_OUT: blocked operand of the next stage (nKtiles K tiles); else out (ld = ldo)
    long long ldo;
    int Ktiles, nKtiles;
    int Kf, C, Rf, M, hM, cR2, cU2;
    int tiles_r, tiles_c;
    int nKf, nhm, ncR2;           // FOLD_OUT: folding of the data columns for the next stage
    double alpha, oprime, sprime, scale, sgn;
    double nalpha, nsprime;
    // per-(plane, stage) phase tables built by phase_table_kernel, so that no CTA spends FP64
    // pipe time on sincospi:  rot[Rfp] | seed[4][Rfp] | post+[Rfp] | post-[Rfp] | pre2[nKfp]
    const double2 *tab;
    int Rfp, nKfp;
    const int *kmax;    // row stage only: K rows that actually hold data, per column tile (NULL: all)
    int ntile;
    int intensity;      // column stage: write |F|^2 as float64 (ld = ldo) instead of the complex field
    // de-phasing of the two CTAs that share an SM (see mft_folded_kernel): per-launch slot counters
    // (one per SM, zeroed with the descriptor upload) and the skew in cycles (~ half a tile)
    unsigned *sm_slots;
    long long skew_cycles;
};

__host__ __device__ inline size_t table_elems(int Rfp, int nKfp) { return (size_t)7 * Rfp + nKfp; }

// one thread per table entry
__global__ void __launch_bounds__(256)
phase_table_kernel(const FStageDesc *__restrict__ descs) {
    const FStageDesc d = descs[blockIdx.y];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = (int)table_elems(d.Rfp, d.nKfp);
    if (e >= n) return;
    double2 *tab = const_cast<double2 *>(d.tab);
    const double cR = 0.5 * d.cR2, cU = 0.5 * d.cU2;
    const int sec = e / d.Rfp, u = e % d.Rfp;
    const double up = (double)u + cU;
    double c, s;
    if (sec == 0) {
        cis_cycles(d.alpha, 4.0, up, 1.0, c, s);                       // rotation for K += 4
    } else if (sec <= 4) {
        cis_cycles(d.alpha, (double)(sec - 1) + cR, up, 1.0, c, s);    // twiddle seed at K slot t = sec-1
    } else if (sec == 5) {
        cis_cycles(d.alpha, d.oprime, up - d.sprime, d.sgn, c, s);     // post(+U')
        c *= d.scale; s *= d.scale;
    } else if (sec == 6) {
        cis_cycles(d.alpha, d.oprime, -up - d.sprime, d.sgn, c, s);    // post(-U')
        c *= d.scale; s *= d.scale;
    } else {
        const int r2 = e - 7 * d.Rfp;
        cis_cycles(d.nalpha, d.nsprime, (double)r2 + 0.5 * d.ncR2, -d.sgn, c, s);   // pre2(R2')
    }
    tab[e] = make_double2(c, s);
}

#ifdef LFD_TILE_TIMING
// diagnostic build only: per-CTA timeline (entry, loop start, loop end, exit, smid) of lane 0 of warp 0
__device__ long long *g_tile_timing = nullptr;
#define LFD_TT(slot) if (g_tile_timing && threadIdx.x == 0) tt[slot] = clock64();
#else
#define LFD_TT(slot)
#endif

// z = x * y (complex)
__device__ __forceinline__ void cmul(double xr, double xi, double yr, double yi, double &zr, double &zi) {
    zr = xr * yr - xi * yi;
    zi = xr * yi + xi * yr;
}

// ---- fold: g = pre * f, ge/go = g[+R'] +- g[-R'], written in the row stage's blocked layout --------
// One CTA per folded K row (zero rows up to the next multiple of 16 included), one thread per slot:
// slot sl of column tile tl is data column j+ = nhm + r2 (sl < 16) or its mirror j- = nhm - r2 - ncR2
// (sl >= 16) for the folded column index r2 = 16 tl + (sl & 15), so that a lane of the row stage ends
// up holding T[., j+] and T[., j-] side by side.
__global__ void __launch_bounds__(256)
fold_kernel(const FoldDesc *__restrict__ descs) {
    const FoldDesc d = descs[blockIdx.y];
    const int r = blockIdx.x;
    if (r >= d.Kt * FBK) return;
    const bool live = r < d.Kf;
    const double Rp = (double)r + 0.5 * d.cR2;
    __shared__ double pre[2];
    if (threadIdx.x == 0) {
        double c, s;
        cis_cycles(d.alpha, d.sprime, Rp, -d.sgn, c, s);
        pre[0] = c;
        pre[1] = s;
    }
    __shared__ int tile_nz[FOLD_MAX_TILES];
    for (int t = threadIdx.x; t < d.ntile; t += blockDim.x) tile_nz[t] = 0;
    __syncthreads();
    const double pc = pre[0], ps = pre[1];
    const bool center = (d.cR2 == 0) && (r == 0);
    const int ip = d.hm + r, im = d.hm - r - d.cR2;
    const double2 *__restrict__ rowp = d.D + (long long)ip * d.ldd;
    const double2 *__restrict__ rowm = d.D + (long long)im * d.ldd;
    const bool has_p = ip < d.K;
    const int nKf2 = (d.C + 1) / 2;
    for (int s = threadIdx.x; s < d.ntile * FBC; s += blockDim.x) {
        const int tl = s / FBC, sl = s % FBC;
        const int r2 = tl * (FBC / 2) + (sl & 15);
        const int j = (sl < 16) ? (d.nhm + r2) : (d.nhm - r2 - d.ncR2);
        const bool ok = live && r2 < nKf2 && j >= 0 && j < d.C && !(sl >= 16 && d.ncR2 == 0 && r2 == 0);
        double2 ge = make_double2(0.0, 0.0), go = ge;
        if (ok) {
            double2 a, b;
            if (d.amp != nullptr) {
                a = has_p ? pupil_phasor(d, ip, j) : make_double2(0.0, 0.0);
                b = center ? make_double2(0.0, 0.0) : pupil_phasor(d, im, j);
            } else {
                a = has_p ? rowp[j] : make_double2(0.0, 0.0);
                b = center ? make_double2(0.0, 0.0) : rowm[j];
            }
            const double2 gp = make_double2(a.x * pc - a.y * ps, a.x * ps + a.y * pc);
            if (center) {
                ge = gp;
            } else {
                const double2 gm = make_double2(b.x * pc + b.y * ps, b.y * pc - b.x * ps);   // conj(pre) * b
                ge = make_double2(gp.x + gm.x, gp.y + gm.y);
                go = make_double2(gp.x - gm.x, gp.y - gm.y);
            }
            if (a.x != 0.0 || a.y != 0.0 || b.x != 0.0 || b.y != 0.0) tile_nz[tl] = 1;   // benign race
        }
        d.G[blk_index(d.Kt, tl, r, 0, sl)] = ge;
        d.G[blk_index(d.Kt, tl, r, 1, sl)] = go;
    }
    __syncthreads();
    if (live)
        for (int t = threadIdx.x; t < d.ntile; t += blockDim.x)
            if (tile_nz[t]) {
                atomicMax(d.kmax + t, r + 1);                        // 1 + last row with data
                atomicMax(d.kmax + d.ntile + t, d.Kf - r);          // Kf - first row with data
            }
}

// ---- folded MFT stage ---------------------------------------------------------------------------
// FOLD_OUT (row stage): slots 0..15 of a column tile hold data columns j+ = nhm + r2, slots 16..31
// their mirrors j- = nhm - r2 - ncR2, for the 16 folded column indices r2 = cf_base .. cf_base+15.
template <bool FOLD_OUT>
__global__ void __launch_bounds__(FTHREADS, 2)
mft_folded_kernel(const FStageDesc *__restrict__ descs) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *sD = reinterpret_cast<double2 *>(smem_raw);
    __shared__ __align__(8) uint64_t bar_full[FSTAGES], bar_empty[FSTAGES];

#ifdef LFD_TILE_TIMING
    long long tt[4];
#endif
    LFD_TT(0)
    const FStageDesc d = descs[blockIdx.y];
    const int tile = blockIdx.x;
    // Two CTAs share an SM and every tile of a batch takes the same time, so left alone they run in
    // lockstep and their prologues/epilogues (no MMAs) coincide.  The second CTA to arrive on an SM
    // in the first wave waits half a tile once; the offset then persists for the whole launch.
    unsigned nsm;                                // CTAs with linear id < 2 * (SMs of this device) form the first wave
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
    if (blockIdx.y * gridDim.x + blockIdx.x < 2 * nsm) {
        __shared__ unsigned s_slot;
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            s_slot = atomicAdd(d.sm_slots + (smid & 255), 1u);
        }
        __syncthreads();
        if (s_slot & 1u) {
            const long long t0 = clock64();
            while (clock64() - t0 < d.skew_cycles) __nanosleep(2000);
        }
    }
    if (tile >= d.tiles_r * d.tiles_c) return;
    const int tr = tile % d.tiles_r, tc = tile / d.tiles_r;
    const int r_base = tr * FBR;
    const int c_base = FOLD_OUT ? tc * (FBC / 2) : tc * FBC;   // FOLD_OUT: base of the folded column index

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    // the row stage stops at the last K tile that holds data for this column tile (a disc in its bounding
    // box leaves ~21% of them empty)
    // ... and starts at the first one that does (a central obscuration empties the first K tiles of the
    // central column tiles; folded K runs from the centre outwards)
    int Kneed = d.Kf, Kfirst = 0;
    if (FOLD_OUT && d.kmax != nullptr) {
        Kneed = min(d.Kf, d.kmax[tc]);
        Kfirst = max(0, d.Kf - d.kmax[d.ntile + tc]);
    }
    const int KT = (Kneed + FBK - 1) / FBK;
    const int KT0 = min(Kfirst / FBK, KT);
    const int nkt = KT - KT0;

    // operand pipeline: thread 0 issues one 16 KB bulk copy per K tile into a 4-slot ring; "full"
    // barriers carry the byte count, "empty" barriers collect one arrival per warp
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < FSTAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], FTHREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const double2 *gblk = d.G + ((size_t)tc * d.Ktiles + KT0) * FBLOCK_ELEMS;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < FSTAGES - 1; ++s)
            if (s < nkt) {
                mbar_expect_tx(&bar_full[s], FBLOCK_BYTES);
                bulk_g2s(sD + s * FBLOCK_ELEMS, gblk + (size_t)s * FBLOCK_ELEMS, FBLOCK_BYTES, &bar_full[s]);
            }
    }

    // accA[mb][q][part]: A = cos-GEMM on ge; accB: B = sin-GEMM on go.  part 0 = Re columns, 1 = Im.
    double accA[2][4][2][2], accB[2][4][2][2];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                accA[mb][q][p][0] = accA[mb][q][p][1] = 0.0;
                accB[mb][q][p][0] = accB[mb][q][p][1] = 0.0;
            }

    // twiddle state (cos, sin) of this lane's A-fragment slots and the per-row rotation for K += 4.
    // The lane's second row block is 8 rows below the first: one extra complex multiply, not a sincospi.
    double tc_[2], ts_[2], rc_[2], rs_[2];
    const double cR = 0.5 * d.cR2, cU = 0.5 * d.cU2;
    const double up0 = (double)(r_base + warp * 16 + g) + cU;            // U' of this lane's first row
    const int u0 = r_base + warp * 16 + g;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const double2 v = d.tab[u0 + 8 * mb];
        rc_[mb] = v.x;
        rs_[mb] = v.y;
    }

    const int gx = g ^ (t << 1);                                         // swizzled column of this lane's B slot
    for (int kt = KT0; kt < KT; ++kt) {
        const int j = kt - KT0, slot = j % FSTAGES;
        if (tid == 0) {
            const int nj = j + FSTAGES - 1;                              // refill the slot iteration j-1 used
            if (nj < nkt) {
                const int ns = nj % FSTAGES;
                if (nj >= FSTAGES) mbar_wait(&bar_empty[ns], ((nj / FSTAGES) - 1) & 1);
                mbar_expect_tx(&bar_full[ns], FBLOCK_BYTES);
                bulk_g2s(sD + ns * FBLOCK_ELEMS, gblk + (size_t)nj * FBLOCK_ELEMS, FBLOCK_BYTES, &bar_full[ns]);
            }
        }
        __syncwarp();
        mbar_wait(&bar_full[slot], (j / FSTAGES) & 1);
        if (kt == KT0) { LFD_TT(1) }
        if (kt == 0) {
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                const double2 v = d.tab[(size_t)(1 + t) * d.Rfp + u0 + 8 * mb];
                tc_[mb] = v.x;
                ts_[mb] = v.y;
            }
        } else if (kt == KT0 || (kt % FRESEED_TILES) == 0) {             // late start or very long K
            const double rp = (double)(kt * FBK + t) + cR;               // R' of this lane's K slot
            double sc, ss;
            cis_cycles(d.alpha, rp, up0, 1.0, tc_[0], ts_[0]);
            cis_cycles(d.alpha, rp, 8.0, 1.0, sc, ss);
            cmul(tc_[0], ts_[0], sc, ss, tc_[1], ts_[1]);
        }
        const double2 *se = sD + slot * FBLOCK_ELEMS + t * FBC + gx;
        const double2 *so = se + FBK * FBC;

#pragma unroll
        for (int ks = 0; ks < FBK / 4; ++ks) {
            double2 ve[4], vo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ve[q] = se[ks * 4 * FBC + q * 8];
                vo[q] = so[ks * 4 * FBC + q * 8];
            }
            // next k-step's twiddles, issued ahead of this step's MMAs (separate registers)
            double ntc[2], nts[2];
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) cmul(tc_[mb], ts_[mb], rc_[mb], rs_[mb], ntc[mb], nts[mb]);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    dmma884(accA[mb][q][0][0], accA[mb][q][0][1], tc_[mb], ve[q].x);
                    dmma884(accA[mb][q][1][0], accA[mb][q][1][1], tc_[mb], ve[q].y);
                }
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    dmma884(accB[mb][q][0][0], accB[mb][q][0][1], ts_[mb], vo[q].x);
                    dmma884(accB[mb][q][1][0], accB[mb][q][1][1], ts_[mb], vo[q].y);
                }
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                tc_[mb] = ntc[mb];
                ts_[mb] = nts[mb];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_empty[slot]);
    }
    LFD_TT(2)

    // ---- epilogue: unfold to the +U' and -U' output rows, post phase, scale, transposed store ----
    // FOLD_OUT: pre2(R2') for the lane's four folded columns r2 = c_base + 8*qq + 2t + i
    double p2c[2][2], p2s[2][2];
    if (FOLD_OUT) {
#pragma unroll
        for (int qq = 0; qq < 2; ++qq)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const double2 v = d.tab[(size_t)7 * d.Rfp + c_base + qq * 8 + 2 * t + i];
                p2c[qq][i] = v.x;
                p2s[qq][i] = v.y;
            }
    }

#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
        const double2 pp = d.tab[(size_t)5 * d.Rfp + u0 + 8 * mb], pm = d.tab[(size_t)6 * d.Rfp + u0 + 8 * mb];
        const double ppc = pp.x, pps = pp.y, pmc = pm.x, pms = pm.y;      // scale folded in
        const int u = r_base + warp * 16 + mb * 8 + g;
        if (u >= d.Rf) continue;
        const int kp = d.hM + u, km = d.hM - u - d.cU2;
        const bool has_p = kp < d.M;
        const bool has_m = (km >= 0) && !(d.cU2 == 0 && u == 0);

        if (!FOLD_OUT) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int c = c_base + q * 8 + 2 * t + i;
                    if (c >= d.C) continue;
                    const double Ar = accA[mb][q][0][i], Ai = accA[mb][q][1][i];
                    const double Br = d.sgn * accB[mb][q][0][i], Bi = d.sgn * accB[mb][q][1][i];
                    if (d.intensity) {
                        // |post| = scale, so |F|^2 = scale^2 |A +- i sgn B|^2: no phase needed
                        double *icol = reinterpret_cast<double *>(d.O) + (long long)c * d.ldo;
                        const double s2 = ppc * ppc + pps * pps;
                        if (has_p) { const double xr = Ar - Bi, xi = Ai + Br; icol[kp] = s2 * (xr * xr + xi * xi); }
                        if (has_m) { const double xr = Ar + Bi, xi = Ai - Br; icol[km] = s2 * (xr * xr + xi * xi); }
                        continue;
                    }
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
        } else {
#pragma unroll
            for (int qq = 0; qq < 2; ++qq)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r2 = c_base + qq * 8 + 2 * t + i;
                    const bool pad_row = r2 >= d.nKf;                     // zero K rows of the column stage
                    const bool center = (d.ncR2 == 0) && (r2 == 0);
                    // column j+ lives in q = qq, its mirror j- in q = qq + 2
                    const double Apr = accA[mb][qq][0][i], Api = accA[mb][qq][1][i];
                    const double Bpr = d.sgn * accB[mb][qq][0][i], Bpi = d.sgn * accB[mb][qq][1][i];
                    const double Amr = accA[mb][qq + 2][0][i], Ami = accA[mb][qq + 2][1][i];
                    const double Bmr = d.sgn * accB[mb][qq + 2][0][i], Bmi = d.sgn * accB[mb][qq + 2][1][i];
                    const double pc2 = p2c[qq][i], ps2 = p2s[qq][i];
#pragma unroll
                    for (int side = 0; side < 2; ++side) {
                        if (side == 0 ? !has_p : !has_m) continue;
                        const double sg = side == 0 ? 1.0 : -1.0;       // +U' row: A + i sgn B ; -U' row: A - i sgn B
                        const double pc = side == 0 ? ppc : pmc, ps = side == 0 ? pps : pms;
                        const int k = side == 0 ? kp : km;
                        double xr, xi, tpr, tpi, tmr, tmi, gpr, gpi, gmr, gmi;
                        xr = Apr - sg * Bpi; xi = Api + sg * Bpr;
                        cmul(xr, xi, pc, ps, tpr, tpi);                   // T[k][j+]
                        xr = Amr - sg * Bmi; xi = Ami + sg * Bmr;
                        cmul(xr, xi, pc, ps, tmr, tmi);                   // T[k][j-]
                        cmul(tpr, tpi, pc2, ps2, gpr, gpi);               // pre2 * T+
                        cmul(tmr, tmi, pc2, -ps2, gmr, gmi);              // conj(pre2) * T-
                        // column k of the column stage: tile k / 32, slot k % 32
                        double2 *ge = d.O + blk_index(d.nKtiles, k / FBC, r2, 0, k % FBC);
                        double2 *go = ge + FBK * FBC;
                        if (pad_row) {
                            *ge = make_double2(0.0, 0.0);
                            *go = make_double2(0.0, 0.0);
                        } else if (center) {
                            *ge = make_double2(gpr, gpi);
                            *go = make_double2(0.0, 0.0);
                        } else {
                            *ge = make_double2(gpr + gmr, gpi + gmi);
                            *go = make_double2(gpr - gmr, gpi - gmi);
                        }
                    }
                }
        }
    }
#ifdef LFD_TILE_TIMING
    if (g_tile_timing && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long *o = g_tile_timing + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 6;
        o[0] = tt[0]; o[1] = tt[1]; o[2] = tt[2]; o[3] = clock64(); o[4] = smid; o[5] = FOLD_OUT;
    }
#endif
}

static inline size_t f_align(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr long long SKEW_CYCLES_PER_KTILE = 2400;   // half of the ~4.7k cycles two co-resident CTAs spend per k-tile
constexpr size_t SLOT_BYTES = 2 * 256 * sizeof(unsigned);   // per-SM slot counters, one set per MFT launch

// blocked operand sizes: row stage = ceil(Kf2/16) column tiles x ceil(Kf1/16) K tiles, column stage =
// ceil(M/32) column tiles x ceil(Kf2/16) K tiles, 16 KB each
static inline size_t g1_bytes(const lfd_mft_desc &p) {
    return (size_t)(((p.n + 1) / 2 + FBC / 2 - 1) / (FBC / 2)) * (((p.m + 1) / 2 + FBK - 1) / FBK) * FBLOCK_BYTES;
}
static inline size_t g2_bytes(const lfd_mft_desc &p) {
    return (size_t)((p.M + FBC - 1) / FBC) * (((p.n + 1) / 2 + FBK - 1) / FBK) * FBLOCK_BYTES;
}

size_t folded_workspace_bytes(const lfd_mft_desc *descs, int count) {
    size_t kmax_ints = 0;
    for (int i = 0; i < count; ++i) kmax_ints += 2 * (((descs[i].n + 1) / 2 + FBC / 2 - 1) / (FBC / 2));
    size_t bytes = f_align((size_t)count * (sizeof(FoldDesc) + 2 * sizeof(FStageDesc)) + SLOT_BYTES + kmax_ints * sizeof(int), 256);
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        bytes += g1_bytes(p) + g2_bytes(p);
        const int Rfp1 = ((p.M + 1) / 2 + FBR - 1) / FBR * FBR, Rfp2 = ((p.N + 1) / 2 + FBR - 1) / FBR * FBR;
        const int nKfp = ((p.n + 1) / 2 + FBC / 2 - 1) / (FBC / 2) * (FBC / 2);
        bytes += f_align((table_elems(Rfp1, nKfp) + table_elems(Rfp2, 0)) * sizeof(double2), 256);
    }
    return bytes;
}

int launch_mft_folded(const lfd_mft_desc *descs, int count, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream, const lfd_pupil_src *src = nullptr, int intensity_out = 0) {
    size_t need = folded_workspace_bytes(descs, count);
    LFD_REQUIRE(workspace_bytes >= need, "lfd_mft_c128_batched: workspace too small (%zu < %zu)",
                workspace_bytes, need);
    LFD_REQUIRE(count <= 32767, "lfd_mft_c128_batched: at most 32767 planes per call (got %d)", count);
    int dev = 0;
    LFD_CUDA_OK(cudaGetDevice(&dev));
    if (ensure_dynamic_smem(dev, (const void *)mft_folded_kernel<true>, (int)FSMEM_BYTES)) return 1;
    if (ensure_dynamic_smem(dev, (const void *)mft_folded_kernel<false>, (int)FSMEM_BYTES)) return 1;
    const size_t desc_bytes = (size_t)count * (sizeof(FoldDesc) + 2 * sizeof(FStageDesc));
    size_t kmax_ints = 0;
    for (int i = 0; i < count; ++i) kmax_ints += 2 * (((descs[i].n + 1) / 2 + FBC / 2 - 1) / (FBC / 2));
    const size_t hdr_bytes = desc_bytes + SLOT_BYTES + kmax_ints * sizeof(int);   // counters and support maps start at zero
    char *h = (char *)calloc(hdr_bytes, 1);                 // slot counters start at zero
    LFD_REQUIRE(h != nullptr, "out of host memory");
    FoldDesc *hf = (FoldDesc *)h;
    FStageDesc *hs = (FStageDesc *)(h + (size_t)count * sizeof(FoldDesc));
    unsigned *slots_dev = (unsigned *)((char *)workspace + desc_bytes);
    int *kmax_dev = (int *)((char *)workspace + desc_bytes + SLOT_BYTES);
    char *ws = (char *)workspace;
    size_t off = f_align(hdr_bytes, 256);
    int max_rows = 0, max_t1 = 0, max_t2 = 0, max_tab = 0;
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        if (!(p.m > 0 && p.n > 0 && p.M > 0 && p.N > 0 && p.ldo >= p.N && p.out &&
              (src ? (src[i].amp && src[i].opd && src[i].wavelength != 0.0 && src[i].r0 >= 0 && src[i].c0 >= 0 &&
                      src[i].r0 + p.m <= src[i].n_r && src[i].c0 + p.n <= src[i].n_c)
                   : (p.f && p.ldf >= p.n)))) {
            free(h);
            LFD_REQUIRE(false, "lfd_mft_c128_batched: plane %d has invalid shape/ld/pointers", i);
        }
        const int Kf1 = (p.m + 1) / 2, Kf2 = (p.n + 1) / 2;
        const int Kt1 = (Kf1 + FBK - 1) / FBK, Kt2 = (Kf2 + FBK - 1) / FBK;
        double2 *G1 = (double2 *)(ws + off);
        off += g1_bytes(p);
        double2 *G2 = (double2 *)(ws + off);
        off += g2_bytes(p);
        const int Rfp1 = ((p.M + 1) / 2 + FBR - 1) / FBR * FBR, Rfp2 = ((p.N + 1) / 2 + FBR - 1) / FBR * FBR;
        const int nKfp = (Kf2 + FBC / 2 - 1) / (FBC / 2) * (FBC / 2);
        double2 *tab1 = (double2 *)(ws + off);
        double2 *tab2 = tab1 + table_elems(Rfp1, nKfp);
        off += f_align((table_elems(Rfp1, nKfp) + table_elems(Rfp2, 0)) * sizeof(double2), 256);
        if ((int)table_elems(Rfp1, nKfp) > max_tab) max_tab = (int)table_elems(Rfp1, nKfp);
        if ((int)table_elems(Rfp2, 0) > max_tab) max_tab = (int)table_elems(Rfp2, 0);
        const double sgn = p.inverse ? 1.0 : -1.0;
        double scale = p.unitary ? sqrt(fabs(p.alpha_r * p.alpha_c)) : 1.0;
        if (p.inverse) scale /= ((double)p.m * (double)p.n);
        const int cRm = (p.m % 2 == 0), cRn = (p.n % 2 == 0), cUM = (p.M % 2 == 0), cUN = (p.N % 2 == 0);

        FoldDesc &fd = hf[i];
        fd.D = (const double2 *)p.f; fd.ldd = p.ldf; fd.G = G1;
        fd.K = p.m; fd.C = p.n; fd.Kf = Kf1; fd.hm = p.m / 2; fd.cR2 = cRm; fd.Kt = Kt1;
        fd.alpha = p.alpha_r; fd.sprime = p.shift_r + 0.5 * cUM; fd.sgn = sgn;
        if (src) {
            fd.D = nullptr; fd.ldd = 0;
            fd.amp = src[i].amp; fd.opd = src[i].opd; fd.mask = src[i].mask;
            fd.pld = src[i].n_c; fd.pr0 = src[i].r0; fd.pc0 = src[i].c0; fd.wavelength = src[i].wavelength;
            fd.inv_wavelength = 1.0 / src[i].wavelength;
        } else {
            fd.amp = nullptr; fd.opd = nullptr; fd.mask = nullptr; fd.pld = 0; fd.pr0 = fd.pc0 = 0; fd.wavelength = 1.0; fd.inv_wavelength = 1.0;
        }
        fd.nhm = p.n / 2; fd.ncR2 = cRn; fd.ntile = (Kf2 + FBC / 2 - 1) / (FBC / 2); fd.pad2_ = 0;
        fd.kmax = kmax_dev;
        if (fd.ntile > FOLD_MAX_TILES) { free(h); LFD_REQUIRE(false, "lfd_mft_c128_batched: plane %d is too wide (%d columns)", i, p.n); }
        if (Kt1 * FBK > max_rows) max_rows = Kt1 * FBK;

        // stage 1 (rows): K = m, C = n, output rows M; result folded along its columns for stage 2
        FStageDesc &s1 = hs[i];
        s1.G = G1; s1.O = G2; s1.ldo = 0;
        s1.Kf = Kf1; s1.C = p.n; s1.Rf = (p.M + 1) / 2; s1.M = p.M; s1.hM = p.M / 2;
        s1.cR2 = cRm; s1.cU2 = cUM;
        s1.alpha = p.alpha_r; s1.oprime = p.off_r - 0.5 * cRm; s1.sprime = p.shift_r + 0.5 * cUM;
        s1.scale = 1.0; s1.sgn = sgn;
        s1.nKf = Kf2; s1.nhm = p.n / 2; s1.ncR2 = cRn; s1.Ktiles = Kt1; s1.nKtiles = Kt2;
        s1.nalpha = p.alpha_c; s1.nsprime = p.shift_c + 0.5 * cUN;
        s1.tab = tab1; s1.Rfp = Rfp1; s1.nKfp = nKfp;
        s1.kmax = kmax_dev; s1.ntile = fd.ntile; kmax_dev += 2 * fd.ntile;
        s1.sm_slots = slots_dev; s1.skew_cycles = (long long)((Kf1 + FBK - 1) / FBK) * SKEW_CYCLES_PER_KTILE;
        s1.tiles_r = (s1.Rf + FBR - 1) / FBR; s1.tiles_c = (Kf2 + FBC / 2 - 1) / (FBC / 2);
        if (s1.tiles_r * s1.tiles_c > max_t1) max_t1 = s1.tiles_r * s1.tiles_c;

        // stage 2 (columns): K = n, C = M, output rows N -> out (M x N after the transposed store)
        FStageDesc &s2 = hs[count + i];
        s2.G = G2; s2.O = (double2 *)p.out; s2.ldo = p.ldo;
        s2.Kf = Kf2; s2.C = p.M; s2.Rf = (p.N + 1) / 2; s2.M = p.N; s2.hM = p.N / 2;
        s2.cR2 = cRn; s2.cU2 = cUN;
        s2.alpha = p.alpha_c; s2.oprime = p.off_c - 0.5 * cRn; s2.sprime = p.shift_c + 0.5 * cUN;
        s2.scale = scale; s2.sgn = sgn;
        s2.nKf = 0; s2.nhm = 0; s2.ncR2 = 0; s2.Ktiles = Kt2; s2.nKtiles = 0; s2.nalpha = 0.0; s2.nsprime = 0.0;
        s2.tab = tab2; s2.Rfp = Rfp2; s2.nKfp = 0; s2.kmax = nullptr; s2.ntile = 0; s2.intensity = intensity_out; s1.intensity = 0;
        s2.sm_slots = slots_dev + 256; s2.skew_cycles = (long long)((Kf2 + FBK - 1) / FBK) * SKEW_CYCLES_PER_KTILE;
        s2.tiles_r = (s2.Rf + FBR - 1) / FBR; s2.tiles_c = (s2.C + FBC - 1) / FBC;
        if (s2.tiles_r * s2.tiles_c > max_t2) max_t2 = s2.tiles_r * s2.tiles_c;
    }
    cudaError_t e = cudaMemcpyAsync(workspace, h, hdr_bytes, cudaMemcpyHostToDevice, stream);
    free(h);
    LFD_CUDA_OK(e);
    const FoldDesc *df = (const FoldDesc *)workspace;
    const FStageDesc *ds = (const FStageDesc *)((char *)workspace + (size_t)count * sizeof(FoldDesc));
    phase_table_kernel<<<dim3((max_tab + 255) / 256, 2 * count), 256, 0, stream>>>(ds);
    LFD_CUDA_OK(cudaGetLastError());
    fold_kernel<<<dim3(max_rows, count), 256, 0, stream>>>(df);
    LFD_CUDA_OK(cudaGetLastError());
    mft_folded_kernel<true><<<dim3(max_t1, count), FTHREADS, FSMEM_BYTES, stream>>>(ds);
    LFD_CUDA_OK(cudaGetLastError());
    mft_folded_kernel<false><<<dim3(max_t2, count), FTHREADS, FSMEM_BYTES, stream>>>(ds + count);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch(4);
    return 0;
}

}  // namespace lfd

#ifdef LFD_TILE_TIMING
extern "C" int lfd_debug_tile_timing(long long *buf_dev) {
    return (int)cudaMemcpyToSymbol(lfd::g_tile_timing, &buf_dev, sizeof(buf_dev));
}
#endif

extern "C" int lfd_mft_c128_from_pupil(const lfd_mft_desc *descs, const lfd_pupil_src *src, int count,
                                       int intensity_out, void *workspace, size_t workspace_bytes, void *stream) {
    if (count == 0) return 0;
    LFD_REQUIRE(descs && src && workspace && count > 0, "lfd_mft_c128_from_pupil: bad arguments");
    if (lfd::mft_resolve_execution(descs, count) == LFD_MFT_CZT)
        return lfd::launch_mft_czt(descs, count, workspace, workspace_bytes, (cudaStream_t)stream, src, intensity_out, false);
    return lfd::launch_mft_folded(descs, count, workspace, workspace_bytes, (cudaStream_t)stream, src, intensity_out);
}
