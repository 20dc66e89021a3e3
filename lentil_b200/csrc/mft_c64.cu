// K2b — matrix Fourier transform, complex64 semantics, 3xTF32 split precision on the 5th-gen
// tensor cores: tcgen05.mma kind::tf32 issued by one thread, accumulators in TMEM.
//
// Same folded algorithm as mft_folded.cu (real cos/sin twiddles x folded complex data, see the
// derivation there); every real product a*b is evaluated as a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with
// a_hi = tf32(a), a_lo = a - a_hi (exact in fp32), accumulated in fp32 in TMEM.
//
// Accuracy note (measured, scripts/micro/umma_test.cu): the tensor core's fp32 accumulation
// TRUNCATES — a coherent sum loses ~4e-8 of its magnitude per MMA accumulation, always downwards.
// The small hi*lo and lo*hi corrections therefore go to their own TMEM accumulators and are added
// to the hi*hi sum once, in the epilogue (round-to-nearest), so the large accumulator sees one
// rounding per k-block instead of three.  What remains is a loss of TRUNC_LOSS_PER_PRODUCT x K
// (relative) on a fully coherent sum — the same constant for K = 121, 251 and 501 — and ~0 on an
// incoherent one; the epilogue multiplies by (1 + TRUNC_LOSS_PER_PRODUCT x K), which cancels the
// expected loss at a PSF peak and perturbs any other element by at most that fraction of its own
// magnitude.  Peak-normalised PSF error after compensation: ~1e-6 .. 4e-6 (gate 1e-5).
//
// Per CTA (one tile = 256 folded rows x 32 complex columns, K looped in blocks of 8):
//   warp 0        : TMEM alloc; one lane issues 12 UMMAs (128x64x8) per k-block and commits
//                   them to the stage's `empty` mbarrier
//   warps 1..8    : one thread per folded row.  Each k-block they generate the row's 8 cos/sin
//                   twiddles — an fp64 twiddle carried across k-blocks by one complex rotation,
//                   times 8 fp32 in-block factors — split them to tf32 hi/lo and store them in the
//                   UMMA canonical K-major (no-swizzle) layout, then fence.proxy.async and arrive on
//                   the stage's `fullA` mbarrier.  After the K loop the same warps are the epilogue:
//                   tcgen05.ld their TMEM lane, unfold (+U', -U'), apply the post phase and either
//                   write complex64 output or the hi/lo-split, column-folded operand of the next stage.
//   warp 9        : one lane moves the data operand of a k-block (ge/go, hi/lo: 8 KB) with a single
//                   cp.async.bulk into a 12-slot ring, signalling the slot's `fullB` mbarrier by tx count
// Data operand in HBM: PRE-BLOCKED in the UMMA canonical layout.  Block (column tile, k-block) is
// 8 KB = 4 planes (ge_hi, ge_lo, go_hi, go_lo) x [n/8][k/4][n%8][k%4] floats for 64 real columns
// x 8 K, so that one contiguous bulk copy lands it ready for the tensor core.  fold_split_kernel
// (stage 1) and the stage-1 epilogue (stage 2) write that layout directly.
#include "lfd_common.cuh"

namespace lfd {
namespace c64 {

constexpr int TM = 256;       // folded rows per CTA = 2 UMMA row tiles
constexpr int TN = 32;        // complex columns per CTA = 64 real columns (UMMA N)
constexpr int NR = 2 * TN;    // UMMA N
constexpr int HALF = TN / 2;  // FOLD_OUT: folded columns per tile (HALF columns j+ and their HALF mirrors)
constexpr int KB = 8;         // folded K per k-block (UMMA K for tf32)
// measured with scripts/gpu_c64.py (all-ones input, coherent everywhere): relative loss per accumulated product
constexpr double TRUNC_LOSS_PER_PRODUCT = 6.2e-9;
constexpr int NA = 3;         // twiddle (A operand) stages, 32 KB each
constexpr int NB = 12;        // data (B operand) slots, 8 KB each: a deeper ring, because the data comes from
                              // L2/HBM (~2000 cycles) while a k-block of MMAs lasts ~400
constexpr int A_ARR = 256 * KB * 4;            // one twiddle plane: 8 KB
constexpr int B_ARR = NR * KB * 4;             // one data plane: 2 KB
constexpr int A_BYTES = 4 * A_ARR, B_BYTES = 4 * B_ARR;
constexpr int B_RING = NA * A_BYTES;           // byte offset of the data ring behind the twiddle stages
constexpr int NGEN = 256;                      // generator / epilogue threads
constexpr int NTHREADS = 32 + NGEN + 32;         // MMA warp + generators + bulk-copy producer warp
constexpr size_t SMEM_BYTES = (size_t)NA * A_BYTES + (size_t)NB * B_BYTES + 1024;
// Operand tiles are K-major with 32-byte rows (8 tf32 = one UMMA K) in the SWIZZLE_32B canonical layout:
// row r of a tile sits at (r/8)*256 + (r%8)*32 and its two 16-byte K chunks are swapped when (r%8) >= 4
// (address bit 4 ^= bit 7), which is what keeps the tensor core's operand reads bank-conflict free.
constexpr int SBO = 256;
__host__ __device__ __forceinline__ int tile_off(int r, int kc) { return (r >> 3) * SBO + (r & 7) * 32 + ((kc ^ ((r >> 2) & 1)) << 4); }

struct CStage {
    const unsigned char *B;  // pre-blocked data operand, (Npad / 64) x (Kpad / 8) blocks of 8 KB
    long long plane;         // unused
    int Kf, Kpad, C, Npad;
    int Rf, M, hM, cR2, cU2, Rfp;
    int tiles_r, tiles_c;
    double sgn;
    // tables (phase_table_c64_kernel):  W0[Rfp] double2 | ROT8[Rfp] double2 | POST+[Rfp] double2 | POST-[Rfp] double2
    //                                   then S[Rfp][8] float2 | PRE2[nKfp] float2
    const double2 *tabd;
    const float2 *tabf;
    // outputs
    float2 *out; long long ldo;            // final stage: complex64 M x N (row-major), written transposed
    unsigned char *nB; long long nplane;   // FOLD_OUT: next stage's pre-blocked data operand
    int nKf, nKpad, nhm, ncR2, nKfp, ntile;
    const int *kmax;                       // row stage: support map of its input (NULL: all K blocks)
    double alpha, oprime, sprime, scale, nalpha, nsprime;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}

// one lane of a converged warp (the compiler can then keep tcgen05 operands in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                             // LBO: unused for swizzled K-major layouts
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                             // descriptor version 1 (sm_100)
    d |= (uint64_t)6 << 61;                             // layout type SWIZZLE_32B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void umma_commit(uint64_t *b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of element (plane a, real column n, K index r) inside a pre-blocked data operand with nkb k-blocks
__host__ __device__ __forceinline__ long long bop_offset(int a, int n, int r, int nkb) {
    return ((long long)(n / NR) * nkb + r / KB) * B_BYTES + a * B_ARR + tile_off(n % NR, (r % KB) / 4) + (r % 4) * 4;
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---- phase tables (all sincospi of a batch, in fp64) ------------------------------------------------
__global__ void __launch_bounds__(256)
phase_table_c64_kernel(const CStage *__restrict__ descs) {
    const CStage d = descs[blockIdx.y];
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const double cR = 0.5 * d.cR2, cU = 0.5 * d.cU2;
    double2 *td = const_cast<double2 *>(d.tabd);
    float2 *tf = const_cast<float2 *>(d.tabf);
    double c, s;
    if (e < 4 * d.Rfp) {
        const int sec = e / d.Rfp, u = e % d.Rfp;
        const double up = (double)u + cU;
        if (sec == 0) cis_cycles(d.alpha, cR, up, 1.0, c, s);                        // twiddle at K index 0
        else if (sec == 1) cis_cycles(d.alpha, (double)KB, up, 1.0, c, s);           // rotation per k-block
        else if (sec == 2) { cis_cycles(d.alpha, d.oprime, up - d.sprime, d.sgn, c, s); c *= d.scale; s *= d.scale; }
        else { cis_cycles(d.alpha, d.oprime, -up - d.sprime, d.sgn, c, s); c *= d.scale; s *= d.scale; }
        td[e] = make_double2(c, s);
    } else if (e < 4 * d.Rfp + 8 * d.Rfp) {
        const int i = e - 4 * d.Rfp, u = i / 8, j = i % 8;
        cis_cycles(d.alpha, (double)j, (double)u + cU, 1.0, c, s);                    // in-block factor
        tf[i] = make_float2((float)c, (float)s);
    } else if (e < 12 * d.Rfp + d.nKfp) {
        const int r2 = e - 12 * d.Rfp;
        cis_cycles(d.nalpha, d.nsprime, (double)r2 + 0.5 * d.ncR2, -d.sgn, c, s);     // pre2(R2')
        tf[8 * d.Rfp + r2] = make_float2((float)c, (float)s);
    }
}

// ---- fold + split + transpose: complex64 f (K x C) -> stage-1 data operand ----------------------------
struct FoldSplit {
    const float2 *D; long long ldd;
    unsigned char *B; long long plane;  // pre-blocked data operand of stage 1
    int K, C, Kf, Kpad, hm, cR2;
    int permute;                     // 1: columns arranged as HALF (j+) + HALF (j-) per TN-slot tile for FOLD_OUT
    int nhm, ncR2, nKf, slots;       // slots = Npad / 2
    double alpha, sprime, sgn;
    int *kmax;                       // support map, as in mft_folded.cu: [tile] = 1 + last row with data,
    int ntile, pad_;                 // [ntile + tile] = Kf - first row with data (column tile = 32 slots)
};

__global__ void __launch_bounds__(256)
fold_split_kernel(const FoldSplit *__restrict__ descs) {
    const FoldSplit d = descs[blockIdx.z];
    const int r0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
    if (r0 >= d.Kpad || s0 >= d.slots) return;
    __shared__ float tile[4][32][33];           // ge.re, ge.im, go.re, go.im  [r][slot]
    __shared__ float2 pre[32];
    __shared__ int row_nz[32];
    if (threadIdx.x < 32) row_nz[threadIdx.x] = 0;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    if (threadIdx.x < 32) {
        double c, s;
        cis_cycles(d.alpha, d.sprime, (double)(r0 + threadIdx.x) + 0.5 * d.cR2, -d.sgn, c, s);
        pre[threadIdx.x] = make_float2((float)c, (float)s);
    }
    __syncthreads();
    // slot -> source column
    const int slot = s0 + tx;
    int j;
    if (d.permute) {
        const int ct = slot / TN, q = slot % TN, r2 = ct * HALF + (q % HALF);
        j = (q < HALF) ? (d.nhm + r2) : (d.nhm - r2 - d.ncR2);
        if (r2 >= d.nKf) j = -1;
    } else {
        j = slot;
    }
    const bool col_ok = (j >= 0) && (j < d.C);
#pragma unroll
    for (int rr = ty; rr < 32; rr += 8) {
        const int r = r0 + rr;
        float ger = 0.f, gei = 0.f, gor = 0.f, goi = 0.f;
        if (col_ok && r < d.Kf) {
            const int ip = d.hm + r, im = d.hm - r - d.cR2;
            const float2 p = pre[rr];
            float2 a = (ip < d.K) ? d.D[(long long)ip * d.ldd + j] : make_float2(0.f, 0.f);
            const float gpr = a.x * p.x - a.y * p.y, gpi = a.x * p.y + a.y * p.x;
            bool nz = (a.x != 0.f) || (a.y != 0.f);
            if (d.cR2 == 0 && r == 0) {
                ger = gpr; gei = gpi;
            } else {
                float2 b = d.D[(long long)im * d.ldd + j];
                nz = nz || (b.x != 0.f) || (b.y != 0.f);
                const float gmr = b.x * p.x + b.y * p.y, gmi = b.y * p.x - b.x * p.y;
                ger = gpr + gmr; gei = gpi + gmi; gor = gpr - gmr; goi = gpi - gmi;
            }
            if (nz) row_nz[rr] = 1;                 // benign race
        }
        tile[0][rr][tx] = ger; tile[1][rr][tx] = gei; tile[2][rr][tx] = gor; tile[3][rr][tx] = goi;
    }
    __syncthreads();
    if (threadIdx.x < 32 && row_nz[threadIdx.x] && d.kmax != nullptr) {
        const int r = r0 + threadIdx.x, tl = s0 / 32;
        atomicMax(d.kmax + tl, r + 1);
        atomicMax(d.kmax + d.ntile + tl, d.Kf - r);
    }
    // this CUDA block owns one column tile (64 real columns) x 4 k-blocks = 4 x 8 KB of the pre-blocked
    // operand; consecutive threads write consecutive 16-byte chunks (4 K values of one real column)
    const int tileN = s0 / 32, nkb = d.Kpad / KB;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int c = threadIdx.x + 256 * it;          // 0 .. 2047
        const int kbl = c >> 9, a = (c >> 7) & 3, q = c & 127;
        const int n = q >> 1, kc = (q & 1) ^ ((n >> 2) & 1);       // physical chunk q & 1 holds logical K chunk kc
        const int sl = n >> 1, part = n & 1;
        float v[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float x = tile[(a >> 1) * 2 + part][kbl * 8 + kc * 4 + jj][sl];
            const float h = tf32_hi(x);
            v[jj] = (a & 1) ? (x - h) : h;
        }
        unsigned char *dst = d.B + ((long long)tileN * nkb + (r0 / KB + kbl)) * B_BYTES + a * B_ARR + q * 16;
        *(float4 *)dst = make_float4(v[0], v[1], v[2], v[3]);
    }
}

#ifdef LFD_TILE_TIMING
__device__ long long *g_c64_timing = nullptr;   // per CTA: [mma wait fullA, mma wait fullB, mma issue, gen wait emptyA, gen work, gen fence+arrive, prod wait, total]
#define TT_DECL long long tt_a = 0, tt_b = 0, tt_c = 0, tt_t0 = clock64(), tt_x;
#define TT_BEGIN tt_x = clock64();
#define TT_ADD(v) { long long n__ = clock64(); v += n__ - tt_x; tt_x = n__; }
#else
#define TT_DECL
#define TT_BEGIN
#define TT_ADD(v)
#endif

// ---- the tcgen05 stage kernel ------------------------------------------------------------------------
template <bool FOLD_OUT>
__global__ void __launch_bounds__(NTHREADS, 1)
mft_c64_kernel(const CStage *__restrict__ descs) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t fullA_bar[NA], emptyA_bar[NA], fullB_bar[NB], emptyB_bar[NB], tmem_full_bar;
    __shared__ uint32_t tmem_base_s;

    const CStage d = descs[blockIdx.y];
    const int tile = blockIdx.x;
    if (tile >= d.tiles_r * d.tiles_c) return;
    const int tr = tile % d.tiles_r, tc = tile / d.tiles_r;
    const int r_base = tr * TM;

    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = d.Kpad / KB;
    // k-block range that holds data for this column tile (support map of the fold kernel); at least one block
    int kb0 = 0, kb1 = nkb;
    if (FOLD_OUT && d.kmax != nullptr) {
        const int last = min(d.Kf, d.kmax[tc]), first = max(0, d.Kf - d.kmax[d.ntile + tc]);
        kb1 = min(nkb, (last + KB - 1) / KB);
        kb0 = min(first / KB, kb1);
        if (kb1 <= kb0) { kb0 = 0; kb1 = 1; }
    }

    if (tid == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(&fullA_bar[s], NGEN / 32); mbar_init(&emptyA_bar[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&fullB_bar[s], 1); mbar_init(&emptyB_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;

    if (warp == 0) {
        // ================= MMA issuer: the warp stays converged, one elected lane issues =================
        {
            TT_DECL
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int kb = kb0; kb < kb1; ++kb) {
                const int jb = kb - kb0, s = jb % NA;
                TT_BEGIN
                mbar_wait(&fullA_bar[s], (jb / NA) & 1);
                TT_ADD(tt_a)
                mbar_wait(&fullB_bar[jb % NB], (jb / NB) & 1);
                TT_ADD(tt_b)
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t sa = smem_u32(smem + (size_t)s * A_BYTES);
                const uint32_t sb = smem_u32(smem + B_RING + (size_t)(jb % NB) * B_BYTES);
                const uint32_t acc = jb > 0 ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) {
                    const uint32_t ro = rt * 4096;                       // rows 128..255 of each twiddle plane
                    const uint64_t ch = umma_desc(sa + 0 * A_ARR + ro), cl = umma_desc(sa + 1 * A_ARR + ro);
                    const uint64_t sh = umma_desc(sa + 2 * A_ARR + ro), sl = umma_desc(sa + 3 * A_ARR + ro);
                    const uint64_t eh = umma_desc(sb + 0 * B_ARR), el = umma_desc(sb + 1 * B_ARR);
                    const uint64_t oh = umma_desc(sb + 2 * B_ARR), ol = umma_desc(sb + 3 * B_ARR);
                    // TMEM columns per row tile: A_main | A_corr | B_main | B_corr, NR columns each
                    const uint32_t t0 = tmem + rt * 4 * NR;
                    umma_tf32(t0, ch, eh, idesc, acc);               // A_main += cos_hi * ge_hi
                    umma_tf32(t0 + NR, ch, el, idesc, acc);          // A_corr += cos_hi * ge_lo
                    umma_tf32(t0 + NR, cl, eh, idesc, 1u);           //         + cos_lo * ge_hi
                    umma_tf32(t0 + 2 * NR, sh, oh, idesc, acc);      // B_main += sin_hi * go_hi
                    umma_tf32(t0 + 3 * NR, sh, ol, idesc, acc);      // B_corr += sin_hi * go_lo
                    umma_tf32(t0 + 3 * NR, sl, oh, idesc, 1u);       //         + sin_lo * go_hi
                }
                umma_commit(&emptyA_bar[s]);                // frees the twiddle stage and the data slot
                umma_commit(&emptyB_bar[jb % NB]);          // when these MMAs retire
                }
                __syncwarp();
                TT_ADD(tt_c)
            }
            if (elect_one()) umma_commit(&tmem_full_bar);
            __syncwarp();
#ifdef LFD_TILE_TIMING
            if (g_c64_timing && lane == 0) { long long *o = g_c64_timing + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8; o[0] = tt_a; o[1] = tt_b; o[2] = tt_c; o[7] = clock64() - tt_t0; }
#endif
        }
    } else if (warp == 9) {
        // ================= data-operand producer: one bulk copy per k-block =================
        if (lane == 0) {
            const unsigned char *src = d.B + (long long)tc * nkb * B_BYTES;
            for (int kb = kb0; kb < kb1; ++kb) {
                const int jb = kb - kb0, sl = jb % NB;
                if (jb >= NB) mbar_wait(&emptyB_bar[sl], ((jb / NB) - 1) & 1);
                const uint32_t bar = smem_u32(&fullB_bar[sl]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)B_BYTES) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + B_RING + (size_t)sl * B_BYTES)), "l"(src + (long long)kb * B_BYTES),
                               "r"((uint32_t)B_BYTES), "r"(bar) : "memory");
            }
        }
    } else {
        // ================= twiddle generators =================
        const int g = tid - 32;                             // 0..255 = folded row within the tile
        const int u = r_base + g;
        double Wc, Ws, Rc, Rs;
        float2 S[KB];
        {
            const double2 w0 = d.tabd[u], r8 = d.tabd[d.Rfp + u];
            Wc = w0.x; Ws = w0.y; Rc = r8.x; Rs = r8.y;
#pragma unroll
            for (int j = 0; j < KB; ++j) S[j] = d.tabf[(size_t)u * KB + j];
        }
        const uint32_t adst = (g >> 3) * SBO + (g & 7) * 32;    // row g of a 256-row twiddle plane
        const uint32_t swz = ((g >> 2) & 1) << 4;
        TT_DECL
        if (kb0 > 0)   // late start: seed the carried twiddle at the first K block that is processed
            cis_cycles(d.alpha, (double)(kb0 * KB) + 0.5 * d.cR2, (double)u + 0.5 * d.cU2, 1.0, Wc, Ws);
        for (int kb = kb0; kb < kb1; ++kb) {
            const int jb = kb - kb0, s = jb % NA;
            TT_BEGIN
            // the twiddle stage must be free before it is overwritten
            if (jb >= NA) mbar_wait(&emptyA_bar[s], ((jb / NA) - 1) & 1);
            TT_ADD(tt_a)
            // twiddles of this row for the 8 K of the block
            {
                const float wc = (float)Wc, ws = (float)Ws;
                float ch[KB], cl[KB], sh[KB], sl[KB];
#pragma unroll
                for (int j = 0; j < KB; ++j) {
                    const float c = wc * S[j].x - ws * S[j].y, sn = wc * S[j].y + ws * S[j].x;
                    ch[j] = tf32_hi(c); cl[j] = c - ch[j];
                    sh[j] = tf32_hi(sn); sl[j] = sn - sh[j];
                }
                const uint32_t sa = smem_u32(smem + (size_t)s * A_BYTES) + adst;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t o = sa + ((h << 4) ^ swz);
                    st_shared_v4(o + 0 * A_ARR, ch[4 * h], ch[4 * h + 1], ch[4 * h + 2], ch[4 * h + 3]);
                    st_shared_v4(o + 1 * A_ARR, cl[4 * h], cl[4 * h + 1], cl[4 * h + 2], cl[4 * h + 3]);
                    st_shared_v4(o + 2 * A_ARR, sh[4 * h], sh[4 * h + 1], sh[4 * h + 2], sh[4 * h + 3]);
                    st_shared_v4(o + 3 * A_ARR, sl[4 * h], sl[4 * h + 1], sl[4 * h + 2], sl[4 * h + 3]);
                }
                const double nc = Wc * Rc - Ws * Rs, ns = Wc * Rs + Ws * Rc;
                Wc = nc; Ws = ns;
            }
            TT_ADD(tt_b)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy STS -> visible to the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&fullA_bar[s]);
            TT_ADD(tt_c)
        }
#ifdef LFD_TILE_TIMING
        if (g_c64_timing && g == 0) { long long *o = g_c64_timing + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8; o[3] = tt_a; o[4] = tt_b; o[5] = tt_c; }
#endif

        // ================= epilogue =================
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int gw = warp - 1;                            // 0..7
        const int rt = gw >> 2;                             // which 128-row half
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may read
        const int row = rt * 128 + quarter * 32 + lane;     // row within the tile
        const int ue = r_base + row;
        const uint32_t lane_addr = ((uint32_t)(quarter * 32) << 16);
        const uint32_t tA = tmem + lane_addr + rt * 4 * NR, tB = tA + 2 * NR;   // main; the corrections sit NR columns further
        const bool row_ok = ue < d.Rf;
        const int kp = d.hM + ue, km = d.hM - ue - d.cU2;
        const bool has_p = row_ok && (kp < d.M);
        const bool has_m = row_ok && (km >= 0) && !(d.cU2 == 0 && ue == 0);
        const double2 pp = d.tabd[2 * d.Rfp + (row_ok ? ue : 0)], pm = d.tabd[3 * d.Rfp + (row_ok ? ue : 0)];
        const float ppc = (float)pp.x, pps = (float)pp.y, pmc = (float)pm.x, pms = (float)pm.y;
        const float sg = (float)d.sgn;

        if (!FOLD_OUT) {
            // chunks of 16 complex columns; main + correction accumulators are summed here (round to nearest)
            for (int cc = 0; cc < TN / 16; ++cc) {
                float a[32], b[32];
                {
                    float t[32];
                    tmem_ld32(tA + cc * 32, a);
                    tmem_ld32(tA + NR + cc * 32, t);
#pragma unroll
                    for (int i = 0; i < 32; ++i) a[i] += t[i];
                    tmem_ld32(tB + cc * 32, b);
                    tmem_ld32(tB + NR + cc * 32, t);
#pragma unroll
                    for (int i = 0; i < 32; ++i) b[i] += t[i];
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = tc * TN + cc * 16 + i;
                    if (c >= d.C) continue;
                    const float Ar = a[2 * i], Ai = a[2 * i + 1], Br = sg * b[2 * i], Bi = sg * b[2 * i + 1];
                    float2 *col = d.out + (long long)c * d.ldo;
                    if (has_p) { const float xr = Ar - Bi, xi = Ai + Br; col[kp] = make_float2(xr * ppc - xi * pps, xr * pps + xi * ppc); }
                    if (has_m) { const float xr = Ar + Bi, xi = Ai - Br; col[km] = make_float2(xr * pmc - xi * pms, xr * pms + xi * pmc); }
                }
            }
        } else {
            // slots 0..HALF-1 of the tile are columns j+, slots HALF..TN-1 their mirrors j-: chunk cc pairs
            // TMEM columns [32cc, 32cc+32) with [2*HALF+32cc, ...)
            for (int cc = 0; cc < HALF / 16; ++cc) {
                float ap[32], am[32], bp[32], bm[32];
                {
                    float t[32];
                    tmem_ld32(tA + cc * 32, ap);
                    tmem_ld32(tA + NR + cc * 32, t);
#pragma unroll
                    for (int i = 0; i < 32; ++i) ap[i] += t[i];
                    tmem_ld32(tA + 2 * HALF + cc * 32, am);
                    tmem_ld32(tA + NR + 2 * HALF + cc * 32, t);
#pragma unroll
                    for (int i = 0; i < 32; ++i) am[i] += t[i];
                    tmem_ld32(tB + cc * 32, bp);
                    tmem_ld32(tB + NR + cc * 32, t);
#pragma unroll
                    for (int i = 0; i < 32; ++i) bp[i] += t[i];
                    tmem_ld32(tB + 2 * HALF + cc * 32, bm);
                    tmem_ld32(tB + NR + 2 * HALF + cc * 32, t);
#pragma unroll
                    for (int i = 0; i < 32; ++i) bm[i] += t[i];
                }
                const int r2_0 = tc * HALF + cc * 16;
#pragma unroll
                for (int side = 0; side < 2; ++side) {
                    if (side == 0 ? !has_p : !has_m) continue;
                    const float sd = side == 0 ? 1.f : -1.f;
                    const float pc = side == 0 ? ppc : pmc, ps = side == 0 ? pps : pms;
                    const int k = side == 0 ? kp : km;
                    // four K values (r2) at a time: one float4 per (plane, part)
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        float o[4][2][4];      // [ge_hi, ge_lo, go_hi, go_lo][re, im][r2 in group]
#pragma unroll
                        for (int w = 0; w < 4; ++w) {
                            const int i = 4 * v + w;
                            const int r2 = r2_0 + i;
                            const float2 p2 = d.tabf[(size_t)8 * d.Rfp + (r2 < d.nKfp ? r2 : 0)];
                            float xr, xi, tpr, tpi, tmr, tmi;
                            xr = ap[2 * i] - sd * sg * bp[2 * i + 1]; xi = ap[2 * i + 1] + sd * sg * bp[2 * i];
                            tpr = xr * pc - xi * ps; tpi = xr * ps + xi * pc;                 // T[k][j+]
                            xr = am[2 * i] - sd * sg * bm[2 * i + 1]; xi = am[2 * i + 1] + sd * sg * bm[2 * i];
                            tmr = xr * pc - xi * ps; tmi = xr * ps + xi * pc;                 // T[k][j-]
                            const float gpr = tpr * p2.x - tpi * p2.y, gpi = tpr * p2.y + tpi * p2.x;   // pre2 * T+
                            const float gmr = tmr * p2.x + tmi * p2.y, gmi = tmi * p2.x - tmr * p2.y;   // conj(pre2) * T-
                            float ger, gei, gor, goi;
                            if (d.ncR2 == 0 && r2 == 0) { ger = gpr; gei = gpi; gor = 0.f; goi = 0.f; }
                            else { ger = gpr + gmr; gei = gpi + gmi; gor = gpr - gmr; goi = gpi - gmi; }
                            if (r2 >= d.nKf) { ger = gei = gor = goi = 0.f; }
                            o[0][0][w] = tf32_hi(ger); o[1][0][w] = ger - o[0][0][w];
                            o[0][1][w] = tf32_hi(gei); o[1][1][w] = gei - o[0][1][w];
                            o[2][0][w] = tf32_hi(gor); o[3][0][w] = gor - o[2][0][w];
                            o[2][1][w] = tf32_hi(goi); o[3][1][w] = goi - o[2][1][w];
                        }
                        if (r2_0 + 4 * v < d.nKpad) {
#pragma unroll
                            for (int a4 = 0; a4 < 4; ++a4)
#pragma unroll
                                for (int part = 0; part < 2; ++part) {
                                    unsigned char *dst = d.nB + bop_offset(a4, 2 * k + part, r2_0 + 4 * v, d.nKpad / KB);
                                    *(float4 *)dst = make_float4(o[a4][part][0], o[a4][part][1], o[a4][part][2], o[a4][part][3]);
                                }
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static inline size_t al(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
static inline int rup(int v, int m) { return (v + m - 1) / m * m; }

struct Geo {
    int Kf1, Kpad1, slots1, Npad1, Rf1, Rfp1, Kf2, Kpad2, Npad2, Rf2, Rfp2, nKfp;
};
static Geo geo(const lfd_mft_desc &p) {
    Geo g;
    g.Kf1 = (p.m + 1) / 2; g.Kpad1 = rup(g.Kf1, 32);
    g.Kf2 = (p.n + 1) / 2; g.Kpad2 = rup(g.Kf2, 32);
    g.nKfp = rup(g.Kf2, 32);
    g.slots1 = rup(rup(g.Kf2, HALF) * 2, 32); // per HALF folded columns: HALF (j+) + HALF (j-) slots
    g.Npad1 = 2 * g.slots1;
    g.Rf1 = (p.M + 1) / 2; g.Rfp1 = rup(g.Rf1, TM);
    g.Npad2 = 2 * rup(p.M, 32);
    g.Rf2 = (p.N + 1) / 2; g.Rfp2 = rup(g.Rf2, TM);
    return g;
}
static size_t table_bytes(int Rfp, int nKfp) { return al((size_t)4 * Rfp * sizeof(double2)) + al(((size_t)8 * Rfp + nKfp) * sizeof(float2)); }

static size_t kmax_ints(const lfd_mft_desc *descs, int count) {
    size_t n = 0;
    for (int i = 0; i < count; ++i) n += 2 * (size_t)(geo(descs[i]).slots1 / TN);
    return n;
}

size_t c64_workspace_bytes(const lfd_mft_desc *descs, int count) {
    size_t bytes = al((size_t)count * (sizeof(FoldSplit) + 2 * sizeof(CStage)) + kmax_ints(descs, count) * sizeof(int));
    for (int i = 0; i < count; ++i) {
        Geo g = geo(descs[i]);
        bytes += al((size_t)4 * g.Npad1 * g.Kpad1 * sizeof(float));
        bytes += al((size_t)4 * g.Npad2 * g.Kpad2 * sizeof(float));
        bytes += table_bytes(g.Rfp1, g.nKfp) + table_bytes(g.Rfp2, 0);
    }
    return bytes;
}

int launch_mft_c64(const lfd_mft_desc *descs, int count, void *workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
    if (count == 0) return 0;
    LFD_REQUIRE(descs && workspace, "lfd_mft_c64x3_batched: NULL argument");
    LFD_REQUIRE(count <= 32767, "lfd_mft_c64x3_batched: at most 32767 planes per call (got %d)", count);
    const size_t need = c64_workspace_bytes(descs, count);
    LFD_REQUIRE(workspace_bytes >= need, "lfd_mft_c64x3_batched: workspace too small (%zu < %zu)", workspace_bytes, need);
    static bool attr_set = false;
    if (!attr_set) {
        LFD_CUDA_OK(cudaFuncSetAttribute(mft_c64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        LFD_CUDA_OK(cudaFuncSetAttribute(mft_c64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
        attr_set = true;
    }
    const size_t desc_bytes = (size_t)count * (sizeof(FoldSplit) + 2 * sizeof(CStage));
    const size_t hdr = desc_bytes + kmax_ints(descs, count) * sizeof(int);     // support maps start at zero
    int *kmax_dev = (int *)((char *)workspace + desc_bytes);
    char *h = (char *)calloc(hdr, 1);
    LFD_REQUIRE(h != nullptr, "out of host memory");
    FoldSplit *hf = (FoldSplit *)h;
    CStage *hs = (CStage *)(h + (size_t)count * sizeof(FoldSplit));
    char *ws = (char *)workspace;
    size_t off = al(hdr);
    int max_tab = 0, max_t1 = 0, max_t2 = 0, max_fs_x = 0, max_fs_y = 0;
    char *b2_begin = nullptr; size_t b2_total = 0;
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        if (!(p.m > 0 && p.n > 0 && p.M > 0 && p.N > 0 && p.ldf >= p.n && p.ldo >= p.N && p.f && p.out)) {
            free(h);
            LFD_REQUIRE(false, "lfd_mft_c64x3_batched: plane %d has invalid shape/ld/pointers", i);
        }
        const Geo g = geo(p);
        unsigned char *B1 = (unsigned char *)(ws + off); off += al((size_t)4 * g.Npad1 * g.Kpad1 * sizeof(float));
        unsigned char *B2 = (unsigned char *)(ws + off);
        const size_t b2_bytes = al((size_t)4 * g.Npad2 * g.Kpad2 * sizeof(float));
        if (!b2_begin) b2_begin = (char *)B2;
        off += b2_bytes;
        double2 *td1 = (double2 *)(ws + off); off += al((size_t)4 * g.Rfp1 * sizeof(double2));
        float2 *tf1 = (float2 *)(ws + off); off += al(((size_t)8 * g.Rfp1 + g.nKfp) * sizeof(float2));
        double2 *td2 = (double2 *)(ws + off); off += al((size_t)4 * g.Rfp2 * sizeof(double2));
        float2 *tf2 = (float2 *)(ws + off); off += al((size_t)8 * g.Rfp2 * sizeof(float2));
        b2_total = (size_t)((char *)B2 + b2_bytes - b2_begin);
        const double sgn = p.inverse ? 1.0 : -1.0;
        double scale = p.unitary ? sqrt(fabs(p.alpha_r * p.alpha_c)) : 1.0;
        if (p.inverse) scale /= ((double)p.m * (double)p.n);
        const int cRm = (p.m % 2 == 0), cRn = (p.n % 2 == 0), cUM = (p.M % 2 == 0), cUN = (p.N % 2 == 0);

        FoldSplit &f = hf[i];
        f.D = (const float2 *)p.f; f.ldd = p.ldf; f.B = B1; f.plane = (long long)g.Npad1 * g.Kpad1;
        f.K = p.m; f.C = p.n; f.Kf = g.Kf1; f.Kpad = g.Kpad1; f.hm = p.m / 2; f.cR2 = cRm;
        f.permute = 1; f.nhm = p.n / 2; f.ncR2 = cRn; f.nKf = g.Kf2; f.slots = g.slots1;
        f.alpha = p.alpha_r; f.sprime = p.shift_r + 0.5 * cUM; f.sgn = sgn;
        f.kmax = kmax_dev; f.ntile = g.slots1 / TN; f.pad_ = 0;
        if (g.slots1 / 32 > max_fs_x) max_fs_x = g.slots1 / 32;
        if (g.Kpad1 / 32 > max_fs_y) max_fs_y = g.Kpad1 / 32;

        CStage &s1 = hs[i];
        s1.B = B1; s1.plane = f.plane; s1.Kf = g.Kf1; s1.Kpad = g.Kpad1; s1.C = p.n; s1.Npad = g.Npad1;
        s1.Rf = g.Rf1; s1.M = p.M; s1.hM = p.M / 2; s1.cR2 = cRm; s1.cU2 = cUM; s1.Rfp = g.Rfp1;
        s1.tiles_r = g.Rfp1 / TM; s1.tiles_c = g.slots1 / TN;
        s1.sgn = sgn; s1.tabd = td1; s1.tabf = tf1;
        s1.out = nullptr; s1.ldo = 0; s1.nB = B2; s1.nplane = (long long)g.Npad2 * g.Kpad2;
        s1.nKf = g.Kf2; s1.nKpad = g.Kpad2; s1.nhm = p.n / 2; s1.ncR2 = cRn; s1.nKfp = g.nKfp;
        s1.alpha = p.alpha_r; s1.oprime = p.off_r - 0.5 * cRm; s1.sprime = p.shift_r + 0.5 * cUM; s1.scale = 1.0;
        s1.nalpha = p.alpha_c; s1.nsprime = p.shift_c + 0.5 * cUN;
        s1.kmax = kmax_dev; s1.ntile = f.ntile; kmax_dev += 2 * f.ntile;
        s1.scale *= 1.0 + TRUNC_LOSS_PER_PRODUCT * g.Kf1;

        CStage &s2 = hs[count + i];
        s2.B = B2; s2.plane = s1.nplane; s2.Kf = g.Kf2; s2.Kpad = g.Kpad2; s2.C = p.M; s2.Npad = g.Npad2;
        s2.Rf = g.Rf2; s2.M = p.N; s2.hM = p.N / 2; s2.cR2 = cRn; s2.cU2 = cUN; s2.Rfp = g.Rfp2;
        s2.tiles_r = g.Rfp2 / TM; s2.tiles_c = g.Npad2 / NR;
        s2.sgn = sgn; s2.tabd = td2; s2.tabf = tf2;
        s2.out = (float2 *)p.out; s2.ldo = p.ldo; s2.nB = nullptr; s2.nplane = 0;
        s2.nKf = 0; s2.nKpad = 0; s2.nhm = 0; s2.ncR2 = 0; s2.nKfp = 0;
        s2.alpha = p.alpha_c; s2.oprime = p.off_c - 0.5 * cRn; s2.sprime = p.shift_c + 0.5 * cUN; s2.scale = scale;
        s2.nalpha = 0.0; s2.nsprime = 0.0; s2.kmax = nullptr; s2.ntile = 0;
        s2.scale *= 1.0 + TRUNC_LOSS_PER_PRODUCT * g.Kf2;

        const int t1 = 12 * g.Rfp1 + g.nKfp, t2 = 12 * g.Rfp2;
        if (t1 > max_tab) max_tab = t1;
        if (t2 > max_tab) max_tab = t2;
        if (s1.tiles_r * s1.tiles_c > max_t1) max_t1 = s1.tiles_r * s1.tiles_c;
        if (s2.tiles_r * s2.tiles_c > max_t2) max_t2 = s2.tiles_r * s2.tiles_c;
    }
    cudaError_t e = cudaMemcpyAsync(workspace, h, hdr, cudaMemcpyHostToDevice, stream);
    free(h);
    LFD_CUDA_OK(e);
    const FoldSplit *df = (const FoldSplit *)workspace;
    const CStage *ds = (const CStage *)((char *)workspace + (size_t)count * sizeof(FoldSplit));
    // the stage-2 operand is only written where stage 1 has valid rows: clear the padding once
    LFD_CUDA_OK(cudaMemsetAsync(b2_begin, 0, b2_total, stream));
    phase_table_c64_kernel<<<dim3((max_tab + 255) / 256, 2 * count), 256, 0, stream>>>(ds);
    LFD_CUDA_OK(cudaGetLastError());
    fold_split_kernel<<<dim3(max_fs_x, max_fs_y, count), 256, 0, stream>>>(df);
    LFD_CUDA_OK(cudaGetLastError());
    mft_c64_kernel<true><<<dim3(max_t1, count), NTHREADS, SMEM_BYTES, stream>>>(ds);
    LFD_CUDA_OK(cudaGetLastError());
    mft_c64_kernel<false><<<dim3(max_t2, count), NTHREADS, SMEM_BYTES, stream>>>(ds + count);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch(4);
    return 0;
}

}  // namespace c64
}  // namespace lfd

extern "C" size_t lfd_mft_c64x3_workspace_bytes(const lfd_mft_desc *descs, int count) {
    return lfd::c64::c64_workspace_bytes(descs, count);
}

extern "C" int lfd_mft_c64x3_batched(const lfd_mft_desc *descs, int count, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    return lfd::c64::launch_mft_c64(descs, count, workspace, workspace_bytes, (cudaStream_t)stream);
}

#ifdef LFD_TILE_TIMING
extern "C" int lfd_debug_c64_timing(long long *buf_dev) {
    return (int)cudaMemcpyToSymbol(lfd::c64::g_c64_timing, &buf_dev, sizeof(buf_dev));
}
#endif
