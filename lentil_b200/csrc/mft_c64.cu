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
// Per CTA (one tile = 128 folded rows x 48 complex columns, K looped in blocks of 8):
//   warp 0        : TMEM alloc; one elected lane issues 6 UMMAs (128x96x8, "TS" form: A operand read from
//                   TENSOR MEMORY, B operand from shared memory) per k-block and commits them to the
//                   twiddle stage's `empty` mbarrier
//   warps 1..8    : two sets of four warps; set p generates the twiddles of the k-blocks jb = p (mod 2), one thread
//                   per folded row = one TMEM lane.  The row's 8 cos/sin twiddles — an fp64 twiddle carried from
//                   block to block by one complex rotation, times 8 fp32 in-block factors — are split to tf32
//                   hi/lo and written with ONE tcgen05.st (32 columns) straight into the A-operand staging
//                   columns of TMEM: the twiddles never touch shared memory.  tcgen05.wait::st + fence, then
//                   arrive on the stage's `fullA` mbarrier.  After the K loop the same warps are the epilogue:
//                   tcgen05.ld their TMEM lane, unfold (+U', -U'), apply the post phase and either write
//                   complex64 output or the hi/lo-split, column-folded operand of the next stage.
//   warp 9        : one lane moves the data operand of a k-block (ge/go, hi/lo: 12 KB) with a single
//                   cp.async.bulk into a 12-slot ring, signalling the slot's `fullB` mbarrier by tx count
// Why TS: measured with scripts/micro/umma_ts_test.cu, a 128xNx8 tf32 UMMA whose A tile comes from shared
// memory costs max(N/2, 48) cycles (the 4 KB A tile is re-read at 128 B/cycle for every instruction), from
// TMEM it costs the tensor floor N/2.  With A in TMEM the shared-memory traffic of a k-block is only the data
// operand.  Why N = 96 and four twiddle stages: the issuing thread blocks while the tensor queue is full, so every
// fixed cost of its loop (mbarrier waits, fences, descriptor moves; a tcgen05.commit alone costs ~28 tensor
// cycles) is exposed unless it is small against the k-block's MMA time and the operands are ready long before
// they are needed.
// TMEM map (512 columns): [0, 384) accumulators = {A_main, A_corr, B_main, B_corr} x 96;
//                         [384, 512) twiddles   = 4 stages x {cos_hi, cos_lo, sin_hi, sin_lo} x 8.
// Data operand in HBM: PRE-BLOCKED in the UMMA canonical layout.  Block (column tile, k-block) is
// 12 KB = 4 planes (ge_hi, ge_lo, go_hi, go_lo) x [n/8][n%8][k/4 (swizzled)][k%4] floats for 96 real columns
// x 8 K, so that one contiguous bulk copy lands it ready for the tensor core.  fold_split_kernel
// (stage 1) and the stage-1 epilogue (stage 2) write that layout directly.
#include "lfd_common.cuh"

namespace lfd {
namespace c64 {

constexpr int TM = 128;       // folded rows per CTA = one UMMA row tile = the 128 TMEM lanes
constexpr int TN = 48;        // complex columns per CTA = 96 real columns (UMMA N)
constexpr int NR = 2 * TN;    // UMMA N
constexpr int HALF = TN / 2;  // FOLD_OUT: folded columns per tile (HALF columns j+ and their HALF mirrors)
constexpr int KB = 8;         // folded K per k-block (UMMA K for tf32)
static_assert(HALF % KB == 0, "row-stage tiles must write whole k-blocks of the column stage's operand");
// measured with scripts/gpu_c64.py (all-ones input, coherent everywhere): relative loss per accumulated product
constexpr double TRUNC_LOSS_PER_PRODUCT = 6.2e-9;
constexpr int NA = 4;         // twiddle (A operand) stages in TMEM, 32 columns each
constexpr int NB = 12;        // data (B operand) slots in shared memory, 12 KB each: a deep ring, because the data
                              // comes from L2/HBM (~2000 cycles) while a k-block of MMAs lasts ~300
constexpr int B_ARR = NR * KB * 4;             // one data plane: 3 KB
constexpr int B_BYTES = 4 * B_ARR;
constexpr int ACC_COLS = 4 * NR;               // TMEM columns of the accumulators (4 x NR)
constexpr int A_STAGE_COLS = 32;               // 4 planes x 8 K
static_assert(ACC_COLS + NA * A_STAGE_COLS <= 512, "TMEM budget");
static_assert(NB == 12 && NA == 4, "the issuer's loop is unrolled over lcm(NA, NB) = 12 blocks, and 12 / NA must be odd");
#ifndef LFD_EB_QUARTER
#define LFD_EB_QUARTER 1      // lane quarter whose generator warps hand retired data slots back to the producer
#endif
#ifndef LFD_PROD_NS
#define LFD_PROD_NS 100
#endif
constexpr int NSETS = 2;                 // generator sets of four warps; set p makes the twiddles of k-blocks jb = p (mod NSETS)
constexpr int MMA_WARP = 0, PROD_WARP = 1 + 4 * NSETS;
constexpr int NTHREADS = 32 * (2 + 4 * NSETS);   // MMA warp + generators (warps 1 .. 4 NSETS) + bulk-copy producer warp
constexpr size_t SMEM_BYTES = (size_t)NB * B_BYTES + 1024;
// Data tiles are K-major with 32-byte rows (8 tf32 = one UMMA K) in the SWIZZLE_32B canonical layout:
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
    // tables (phase_table_c64_kernel):  W0[Rfp] double2 (unused) | ROT16[Rfp] double2 | POST+[Rfp] double2 | POST-[Rfp] double2
    //                                   then S[Rfp][8] float2 | PRE2[nKfp] float2
    const double2 *tabd;
    const float2 *tabf;
    // outputs
    float2 *out; long long ldo;            // final stage: complex64 M x N (row-major), written transposed
    unsigned char *nB; long long nplane;   // FOLD_OUT: next stage's pre-blocked data operand
    int nKf, nKpad, nhm, ncR2, nKfp, ntile;
    const int *kmax;                       // row stage: support map of its input (NULL: all K blocks)
    int intensity, pad3_;                  // final stage: write |F|^2 as float64 (ldo in doubles) instead of the complex64 field
    double alpha, oprime, sprime, scale, nalpha, nsprime;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// blocking wait: the thread is suspended in mbarrier.try_wait and resumes ~120 cycles after the phase completes
// (scripts/micro/wake_latency.cu); a try_wait on an already completed phase is a 67-cycle round trip (sync_cost.cu)
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}
// polling wait with a sleep between the (non-blocking) tests, for waiters that are far ahead of their consumer
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *b, uint32_t parity, unsigned ns) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}

// One k-block of the issuer in a single asm block, executed by the whole (converged) warp: probe the data barrier of
// the NEXT block (non-blocking test_wait), let the elected lane issue the six 128 x NR x 8 MMAs and the commit, and only
// then read the probe's predicate back — the probe's round trip overlaps the MMA issue instead of preceding it.
// U = jb mod 12 is a compile-time constant (the issuer's loop is unrolled over lcm(NA, NB) blocks), so the twiddle stage,
// the data slot, every barrier address and every descriptor offset are immediates: the six tcgen05.mma go out
// back to back (a loop that recomputes operands between them issues one MMA per ~90 cycles, umma_ts_test.cu).
//   bars : shared-memory address of the barrier block  (unused)[NA] | emptyA[NA] | fullB[NB] | emptyB[NB] | tmem_full
//   lap  : parity of jb / 12;  acc : 0 only for the very first block
template <int U>
__device__ __forceinline__ void issue_kblock(uint32_t tmem, uint64_t desc0, uint32_t idesc, uint32_t acc, uint32_t bars,
                                             uint32_t lap, uint32_t &okB) {
    constexpr int S = U % NA, SL = U % NB, NSL = (U + 1) % NB;
    constexpr int OFF_EMPTYA = 8 * (NA + S), OFF_NEXTB = 8 * (2 * NA + NSL);
    constexpr int PAR_B = ((U + 1) / NB) & 1;                                 // relative to the lap parity
    const uint32_t a0 = tmem + ACC_COLS + S * A_STAGE_COLS;
    const uint64_t eh = desc0 + (uint64_t)((SL * B_BYTES) >> 4), el = eh + (B_ARR >> 4), oh = el + (B_ARR >> 4), ol = oh + (B_ARR >> 4);
    // probe the data barrier of the NEXT block with the non-blocking test_wait; the answer is read after the MMAs
    asm volatile(
        "{\n\t"
        ".reg .pred pe, pacc, pb;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pb, [%15], %16;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pacc, %10, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%1], [%5], %11, %9, pacc;\n\t"     // A_main += cos_hi * ge_hi
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%2], [%5], %12, %9, pacc;\n\t"     // A_corr += cos_hi * ge_lo
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%2], [%6], %11, %9, 1;\n\t"        //         + cos_lo * ge_hi
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%3], [%7], %13, %9, pacc;\n\t"     // B_main += sin_hi * go_hi
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%4], [%7], %14, %9, pacc;\n\t"     // B_corr += sin_hi * go_lo
        "@pe tcgen05.mma.cta_group::1.kind::tf32 [%4], [%8], %13, %9, 1;\n\t"        //         + sin_lo * go_hi
        // ONE commit per k-block (each costs ~28 tensor cycles): it frees the twiddle stage when these MMAs retire;
        // the generators that see it pass the data slot on to the producer
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%17];\n\t"
        "selp.u32 %0, 1, 0, pb;\n\t"
        "}"
        : "=r"(okB)
        : "r"(tmem), "r"(tmem + NR), "r"(tmem + 2 * NR), "r"(tmem + 3 * NR), "r"(a0), "r"(a0 + 8), "r"(a0 + 16), "r"(a0 + 24),
          "r"(idesc), "r"(acc), "l"(eh), "l"(el), "l"(oh), "l"(ol),
          "r"(bars + OFF_NEXTB), "r"(lap ^ PAR_B), "r"(bars + OFF_EMPTYA)
        : "memory");
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
// A operand from tensor memory (lane = row, 8 consecutive columns = K), B operand from a shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void umma_commit(uint64_t *b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)));
}
// 32 columns of this thread's TMEM lane <- 32 registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
// tcgen05.ld of 8 / 16 columns of this thread's lane; the caller waits once for a group of loads
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// byte offset of element (plane a, real column n, K index r) inside a pre-blocked data operand with nkb k-blocks
__host__ __device__ __forceinline__ long long bop_offset(int a, int n, int r, int nkb) {
    return ((long long)(n / NR) * nkb + r / KB) * B_BYTES + a * B_ARR + tile_off(n % NR, (r % KB) / 4) + (r % 4) * 4;
}

// 32 contiguous bytes (one K-major row of a data tile) in one 256-bit store
__device__ __forceinline__ void st_global_v8(void *p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
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
        else if (sec == 1) cis_cycles(d.alpha, (double)(NSETS * KB), up, 1.0, c, s); // rotation per NSETS k-blocks (a generator set owns every NSETS-th block)
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
    int ntile, pad_;                 // [ntile + tile] = Kf - first row with data (column tile = TN slots)
    // fused pupil prep (K1 inside the fold, as in mft_folded.cu): when amp != NULL the input element (i, c) is not read
    // from D but formed as amp * mask * exp(+2 pi i opd / lambda) at pupil pixel (pr0 + i, pc0 + c), lentil/plane.py:502-507
    const double *amp, *opd;
    const unsigned char *mask;       // this segment's mask plane, or NULL
    long long pld;                   // pupil row stride (n_c)
    int pr0, pc0;
    double wavelength;               // 1 / lambda
};

// the phase is reduced in fp64 (in cycles) before the sine/cosine, the phasor is then rounded to complex64
__device__ __forceinline__ float2 pupil_phasor_c64(const FoldSplit &d, int i, int c) {
    const long long pix = (long long)(d.pr0 + i) * d.pld + (d.pc0 + c);
    double a = d.amp[pix];
    if (d.mask != nullptr && d.mask[pix] == 0) a = 0.0;
    if (a == 0.0) return make_float2(0.f, 0.f);
    // the phase is formed and reduced to (-1/2, 1/2] cycles in fp64 (opd / lambda is ~1e1 .. 1e3 cycles), the sine and
    // cosine of the reduced phase are then fp32: the phasor is rounded to complex64 anyway, and the fp64 sincospi
    // made this kernel FP64-issue bound
    const double tcyc = d.opd[pix] * d.wavelength;          // wavelength holds 1 / lambda here (set by the launcher)
    float sn, cs;
    sincospif(2.0f * (float)(tcyc - rint(tcyc)), &sn, &cs);
    const float af = (float)a;
    return make_float2(af * cs, af * sn);
}

__global__ void __launch_bounds__(256)
fold_split_kernel(const FoldSplit *__restrict__ descs) {
    const FoldSplit d = descs[blockIdx.z];
    const int r0 = blockIdx.y * 32, s0 = blockIdx.x * TN;      // 32 folded K rows x one column tile (TN slots)
    if (r0 >= d.Kpad || s0 >= d.slots) return;
    __shared__ float tile[4][32][TN + 1];       // ge.re, ge.im, go.re, go.im  [r][slot]
    __shared__ float2 pre[32];
    __shared__ int row_nz[32];
    if (threadIdx.x < 32) row_nz[threadIdx.x] = 0;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4, threads tx < TN carry a slot
    if (threadIdx.x < 32) {
        double c, s;
        cis_cycles(d.alpha, d.sprime, (double)(r0 + threadIdx.x) + 0.5 * d.cR2, -d.sgn, c, s);
        pre[threadIdx.x] = make_float2((float)c, (float)s);
    }
    __syncthreads();
    // slot -> source column
    const int slot = s0 + tx;
    int j = -1;
    if (tx < TN) {
        if (d.permute) {
            const int ct = slot / TN, q = slot % TN, r2 = ct * HALF + (q % HALF);
            j = (q < HALF) ? (d.nhm + r2) : (d.nhm - r2 - d.ncR2);
            if (r2 >= d.nKf) j = -1;
        } else {
            j = slot;
        }
    }
    const bool col_ok = (j >= 0) && (j < d.C);
    if (tx < TN) {
#pragma unroll
        for (int rr = ty; rr < 32; rr += 4) {
            const int r = r0 + rr;
            float ger = 0.f, gei = 0.f, gor = 0.f, goi = 0.f;
            if (col_ok && r < d.Kf) {
                const int ip = d.hm + r, im = d.hm - r - d.cR2;
                const float2 p = pre[rr];
                const bool pupil = d.amp != nullptr;
                float2 a = (ip < d.K) ? (pupil ? pupil_phasor_c64(d, ip, j) : d.D[(long long)ip * d.ldd + j]) : make_float2(0.f, 0.f);
                const float gpr = a.x * p.x - a.y * p.y, gpi = a.x * p.y + a.y * p.x;
                bool nz = (a.x != 0.f) || (a.y != 0.f);
                if (d.cR2 == 0 && r == 0) {
                    ger = gpr; gei = gpi;
                } else {
                    float2 b = pupil ? pupil_phasor_c64(d, im, j) : d.D[(long long)im * d.ldd + j];
                    nz = nz || (b.x != 0.f) || (b.y != 0.f);
                    const float gmr = b.x * p.x + b.y * p.y, gmi = b.y * p.x - b.x * p.y;
                    ger = gpr + gmr; gei = gpi + gmi; gor = gpr - gmr; goi = gpi - gmi;
                }
                if (nz) row_nz[rr] = 1;                 // benign race
            }
            tile[0][rr][tx] = ger; tile[1][rr][tx] = gei; tile[2][rr][tx] = gor; tile[3][rr][tx] = goi;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32 && row_nz[threadIdx.x] && d.kmax != nullptr) {
        const int r = r0 + threadIdx.x, tl = blockIdx.x;
        atomicMax(d.kmax + tl, r + 1);
        atomicMax(d.kmax + d.ntile + tl, d.Kf - r);
    }
    // this CUDA block owns one column tile (NR real columns) x 4 k-blocks = 4 x B_BYTES of the pre-blocked
    // operand; consecutive threads write consecutive 16-byte chunks (4 K values of one real column)
    const int tileN = blockIdx.x, nkb = d.Kpad / KB;
    constexpr int CH_PLANE = 2 * NR, CH_KB = 4 * CH_PLANE;      // 16-byte chunks per plane / per k-block
    for (int c = threadIdx.x; c < 4 * CH_KB; c += 256) {
        const int kbl = c / CH_KB, rem = c % CH_KB, a = rem / CH_PLANE, q = rem % CH_PLANE;
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
__device__ long long *g_c64_timing_buf = nullptr;
__device__ int g_c64_timing_stage = -1;        // -1: both stages record (the column stage overwrites), 1: row stage only, 0: column stage only
__device__ long long *g_c64_trace = nullptr;    // event trace of CTA (0, 0): [6][128] clocks
#define TRACE(ev, j) { if (trace_p && (j) < 128) trace_p[(ev) * 128 + (j)] = clock64(); }
#define TRACE_IF(cond, ev, j) { if (cond) TRACE(ev, j) }
#define g_c64_timing ((g_c64_timing_stage < 0 || g_c64_timing_stage == (int)FOLD_OUT) ? g_c64_timing_buf : (long long *)nullptr)   // per CTA: [mma wait fullA, mma wait fullB, mma issue, gen wait emptyA, gen work, gen fence+arrive, prod wait, total]
#define TT_DECL long long tt_a = 0, tt_b = 0, tt_c = 0, tt_t0 = clock64(), tt_x;
#define TT_BEGIN tt_x = clock64();
#define TT_ADD(v) { long long n__ = clock64(); v += n__ - tt_x; tt_x = n__; }
#else
#define TT_DECL
#define TT_BEGIN
#define TT_ADD(v)
#define TRACE(ev, j) {}
#define TRACE_IF(cond, ev, j) {}
#endif

// ---- the tcgen05 stage kernel ------------------------------------------------------------------------
template <bool FOLD_OUT>
__global__ void __launch_bounds__(NTHREADS, 1)
mft_c64_kernel(const CStage *__restrict__ descs) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * NA + 2 * NB + 1];     // (unused) | emptyA | fullB | emptyB | tmem_full
    uint64_t *const emptyA_bar = bars + NA, *const fullB_bar = bars + 2 * NA,
                   *const emptyB_bar = bars + 2 * NA + NB, *const tmem_full_bar_p = bars + 2 * NA + 2 * NB;
    __shared__ uint32_t tmem_base_s;

#ifdef LFD_TILE_TIMING
    const long long t_entry = clock64();
    long long *const trace_p = (blockIdx.x == 5 && blockIdx.y == 0 && (g_c64_timing_stage < 0 || g_c64_timing_stage == (int)FOLD_OUT)) ? g_c64_trace : nullptr;
#endif
    const CStage d = descs[blockIdx.y];
    const int tile = blockIdx.x;
    if (tile >= d.tiles_r * d.tiles_c) return;
    const int tr = tile % d.tiles_r, tc = tile / d.tiles_r;     // the row tiles of one column tile are neighbours: they
    const int r_base = tr * TM;                                 // read the same data blocks through L2 at the same time

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
    const int nblk = kb1 - kb0;

    if (tid == 0) {
        for (int s = 0; s < NA; ++s) mbar_init(&emptyA_bar[s], 1);
        for (int s = 0; s < NB; ++s) { mbar_init(&fullB_bar[s], 1); mbar_init(&emptyB_bar[s], 1); }
        mbar_init(tmem_full_bar_p, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;

    if (warp == MMA_WARP) {
        // ================= MMA issuer: the warp stays converged, one elected lane issues =================
        {
            TT_DECL
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t desc0 = umma_desc(smem_u32(smem));          // data slot 0, plane ge_hi
            const uint32_t bars_a = smem_u32(bars);
#define LFD_C64_BLOCK(U)                                                                                          \
            if (base + U < nblk) {                                                                                \
                /* data: probed one block ahead inside issue_kblock (the producer runs ~10 blocks ahead, so the   \
                   fallback wait is rare); twiddles: named barrier 1 + stage, on which the four generator warps   \
                   of the block arrive - the issuer sleeps in the barrier unit instead of polling an mbarrier    \
                   next to the generator warps of its scheduler (measured: +6 % over an mbarrier hand-over) */    \
                if (!okB) mbar_wait(&fullB_bar[U % NB], (lap ^ (U / NB)) & 1);                                    \
                TT_ADD(tt_b)                                                                                      \
                asm volatile("bar.sync %0, %1;" ::"n"(1 + U % NA), "n"(32 + 128) : "memory");                     \
                asm volatile("tcgen05.fence::after_thread_sync;");                                                \
                TT_ADD(tt_a)                                                                                      \
                TRACE_IF(lane == 0, 0, base + U)                                                                  \
                issue_kblock<U>(tmem, desc0, idesc, (uint32_t)(base + U), bars_a, lap, okB);                      \
                TRACE_IF(lane == 0, 1, base + U)                                                                  \
                TT_ADD(tt_c)                                                                                      \
            }
            uint32_t lap = 0, okB = 0;
            for (int base = 0; base < nblk; base += 12, lap ^= 1u) {
                TT_BEGIN
                LFD_C64_BLOCK(0) LFD_C64_BLOCK(1) LFD_C64_BLOCK(2) LFD_C64_BLOCK(3) LFD_C64_BLOCK(4) LFD_C64_BLOCK(5)
                LFD_C64_BLOCK(6) LFD_C64_BLOCK(7) LFD_C64_BLOCK(8) LFD_C64_BLOCK(9) LFD_C64_BLOCK(10) LFD_C64_BLOCK(11)
            }
#undef LFD_C64_BLOCK
            if (elect_one()) umma_commit(tmem_full_bar_p);
            __syncwarp();
#ifdef LFD_TILE_TIMING
            if (g_c64_timing && lane == 0) { long long *o = g_c64_timing + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16; o[0] = tt_a; o[1] = tt_b; o[2] = tt_c; o[7] = clock64() - tt_t0; o[13] = nblk; }
#endif
        }
    } else if (warp == PROD_WARP) {
        // ================= data-operand producer: one bulk copy per k-block =================
        if (lane == 0) {
            const unsigned char *src = d.B + ((long long)tc * nkb + kb0) * B_BYTES;
            for (int jb = 0; jb < nblk; ++jb) {
                const int sl = jb % NB;
                if (jb >= NB) mbar_wait_sleep(&emptyB_bar[sl], ((jb / NB) - 1) & 1, LFD_PROD_NS);
                const uint32_t bar = smem_u32(&fullB_bar[sl]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)B_BYTES) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(smem + (size_t)sl * B_BYTES)), "l"(src + (long long)jb * B_BYTES),
                               "r"((uint32_t)B_BYTES), "r"(bar) : "memory");
            }
        }
    } else {
        // ================= twiddle generators: thread <-> folded row <-> TMEM lane =================
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
        const int set = (warp - 1) >> 2;                    // generator set: k-blocks jb = set (mod 2)
        const int row = quarter * 32 + lane;                // row within the tile
        const int u = r_base + row;
        const uint32_t lane_addr = ((uint32_t)(quarter * 32) << 16);
        double Wc, Ws, Rc, Rs;
        float2 S[KB];
        {
            const double2 r16 = d.tabd[d.Rfp + u];
            Rc = r16.x; Rs = r16.y;
#pragma unroll
            for (int j = 0; j < KB; ++j) S[j] = d.tabf[(size_t)u * KB + j];
        }
        // the carried twiddle, seeded at the first K block this set processes
        cis_cycles(d.alpha, (double)((kb0 + set) * KB) + 0.5 * d.cR2, (double)u + 0.5 * d.cU2, 1.0, Wc, Ws);
        const uint32_t tw = tmem + lane_addr + ACC_COLS;
        TT_DECL
        for (int jb = set; jb < nblk; jb += NSETS) {
            const int s = jb % NA;
            TRACE_IF(lane == 0 && warp == 3, 5, jb)
            TRACE_IF(lane == 0 && warp == 1, 10, jb)          // warp 1's timeline goes to the set-1 rows (free at even jb)
            TT_BEGIN
            // twiddles of this row for the 8 K of the block: columns cos_hi[8] | cos_lo[8] | sin_hi[8] | sin_lo[8]
            float v[32];
#ifdef LFD_X_NOCOMPUTE      /* timing experiment (wrong results): constant twiddles, no arithmetic in the K loop */
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = S[j & 7].x;
#else
            {
#ifdef LFD_X_NOFP64         /* timing experiment (wrong results): no FP64 instruction in the K loop */
                const float wc = S[1].x + (float)jb, ws = S[1].y;
#else
                const float wc = (float)Wc, ws = (float)Ws;
#endif
                v[0] = tf32_hi(wc); v[8] = wc - v[0]; v[16] = tf32_hi(ws); v[24] = ws - v[16];     // S[0] = 1
#pragma unroll
                for (int j = 1; j < KB; ++j) {
                    const float c = wc * S[j].x - ws * S[j].y, sn = wc * S[j].y + ws * S[j].x;
                    v[j] = tf32_hi(c); v[8 + j] = c - v[j];
                    v[16 + j] = tf32_hi(sn); v[24 + j] = sn - v[16 + j];
                }
#ifndef LFD_X_NOFP64
                const double nc = Wc * Rc - Ws * Rs, ns = Wc * Rs + Ws * Rc;
                Wc = nc; Ws = ns;
#endif
            }
#endif
            TT_ADD(tt_b)
            TRACE_IF(lane == 0 && warp == 1, 11, jb)
            // the twiddle stage must have been consumed before it is overwritten
            if (jb >= NA) {
                mbar_wait(&emptyA_bar[s], ((jb / NA) - 1) & 1);
                TRACE_IF(lane == 0 && warp == 3, 2, jb)
                asm volatile("tcgen05.fence::after_thread_sync;");
                // block jb - NA has retired: hand its data slot back to the producer
                if (quarter == LFD_EB_QUARTER && lane == 0) mbar_arrive(&emptyB_bar[(jb - NA) % NB]);
            }
            TRACE_IF(lane == 0 && (warp == 3 || warp == 4), warp, jb)
            TRACE_IF(lane == 0 && warp == 1, 12, jb)
            TT_ADD(tt_a)
#ifdef LFD_X_NOST           /* timing experiment (wrong results): the twiddles are computed but not written to TMEM */
            { float acc = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) acc += v[j];
              if (acc == 123.456f) tmem_st32(tw + s * A_STAGE_COLS, v); }
#else
            tmem_st32(tw + s * A_STAGE_COLS, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#endif
            TRACE_IF(lane == 0 && (warp == 3 || warp == 4), warp + 11, jb)
            TRACE_IF(lane == 0 && warp == 1, 13, jb)
            asm volatile("tcgen05.fence::before_thread_sync;");
            asm volatile("bar.arrive %0, %1;" ::"r"(1 + s), "r"(32 + 128) : "memory");
            TRACE_IF(lane == 0, 5 + warp, jb)
            TT_ADD(tt_c)
        }
#ifdef LFD_TILE_TIMING
        if (g_c64_timing && warp == 1 && lane == 0) { long long *o = g_c64_timing + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16; o[3] = tt_a; o[4] = tt_b; o[5] = tt_c; o[8] = tt_t0 - t_entry; o[9] = clock64() - tt_t0; }
        const long long t_e0 = clock64();
#endif

        // ================= epilogue: the two sets split the column chunks =================
        mbar_wait(tmem_full_bar_p, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
#ifdef LFD_TILE_TIMING
        const long long t_e1 = clock64();
#endif
        const int ue = u;
        const uint32_t tA = tmem + lane_addr, tB = tA + 2 * NR;   // main; the corrections sit NR columns further
        const bool row_ok = ue < d.Rf;
        const int kp = d.hM + ue, km = d.hM - ue - d.cU2;
        const bool has_p = row_ok && (kp < d.M);
        const bool has_m = row_ok && (km >= 0) && !(d.cU2 == 0 && ue == 0);
        const double2 pp = d.tabd[2 * d.Rfp + (row_ok ? ue : 0)], pm = d.tabd[3 * d.Rfp + (row_ok ? ue : 0)];
        const float ppc = (float)pp.x, pps = (float)pp.y, pmc = (float)pm.x, pms = (float)pm.y;
        const float sg = (float)d.sgn;

        if (!FOLD_OUT) {
            // chunks of 8 complex columns; main + correction accumulators are summed here (round to nearest)
            constexpr int NCH = TN / 8;
            static_assert(NCH % NSETS == 0, "column chunks must split evenly over the generator sets");
#pragma unroll 1
            for (int cc = set * (NCH / NSETS); cc < (set + 1) * (NCH / NSETS); ++cc) {
                uint32_t am[16], ac[16], bm[16], bc[16];
                tmem_ld16_nowait(tA + cc * 16, am);
                tmem_ld16_nowait(tA + NR + cc * 16, ac);
                tmem_ld16_nowait(tB + cc * 16, bm);
                tmem_ld16_nowait(tB + NR + cc * 16, bc);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = tc * TN + cc * 8 + i;
                    if (c >= d.C) continue;
                    const float Ar = __uint_as_float(am[2 * i]) + __uint_as_float(ac[2 * i]);
                    const float Ai = __uint_as_float(am[2 * i + 1]) + __uint_as_float(ac[2 * i + 1]);
                    const float Br = sg * (__uint_as_float(bm[2 * i]) + __uint_as_float(bc[2 * i]));
                    const float Bi = sg * (__uint_as_float(bm[2 * i + 1]) + __uint_as_float(bc[2 * i + 1]));
                    if (d.intensity) {      // |post|^2 = scale^2 is folded into the squared modulus of the rotated value
                        double *col = (double *)d.out + (long long)c * d.ldo;
                        if (has_p) { const float xr = Ar - Bi, xi = Ai + Br, yr = xr * ppc - xi * pps, yi = xr * pps + xi * ppc; col[kp] = (double)yr * yr + (double)yi * yi; }
                        if (has_m) { const float xr = Ar + Bi, xi = Ai - Br, yr = xr * pmc - xi * pms, yi = xr * pms + xi * pmc; col[km] = (double)yr * yr + (double)yi * yi; }
                        continue;
                    }
                    float2 *col = d.out + (long long)c * d.ldo;
                    if (has_p) { const float xr = Ar - Bi, xi = Ai + Br; col[kp] = make_float2(xr * ppc - xi * pps, xr * pps + xi * ppc); }
                    if (has_m) { const float xr = Ar + Bi, xi = Ai - Br; col[km] = make_float2(xr * pmc - xi * pms, xr * pms + xi * pmc); }
                }
            }
        } else {
            // Slots 0..HALF-1 of the tile are columns j+, slots HALF..TN-1 their mirrors j-.  The output is the data
            // operand of the column stage: K index = folded column r2, real columns n = 2k + part for output row k.
            // One k-block of it (8 r2) is a 32-byte row per n, so a thread that owns row k writes, per plane and
            // part, ONE 256-bit store; set 0 writes the rows +U' (k = kp), set 1 the rows -U' (k = km).
            static_assert(NSETS == 2, "the two generator sets take one output side each");
            const bool has = set == 0 ? has_p : has_m;
            const float sd = set == 0 ? 1.f : -1.f;
            const float pc = set == 0 ? ppc : pmc, ps = set == 0 ? pps : pms;
            const int k = set == 0 ? kp : km;
            const bool swz = ((2 * k) >> 2) & 1;        // SWIZZLE_32B: the two 16-byte K chunks of rows n%8 >= 4 are swapped
#pragma unroll 1
            for (int kbo = 0; kbo < HALF / KB; ++kbo) {
                float ap[16], am[16], bp[16], bm[16];
                {
                    uint32_t x0[16], x1[16], x2[16], x3[16];
                    tmem_ld16_nowait(tA + kbo * 16, x0);
                    tmem_ld16_nowait(tA + NR + kbo * 16, x1);
                    tmem_ld16_nowait(tA + 2 * HALF + kbo * 16, x2);
                    tmem_ld16_nowait(tA + NR + 2 * HALF + kbo * 16, x3);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        ap[i] = __uint_as_float(x0[i]) + __uint_as_float(x1[i]);
                        am[i] = __uint_as_float(x2[i]) + __uint_as_float(x3[i]);
                    }
                    tmem_ld16_nowait(tB + kbo * 16, x0);
                    tmem_ld16_nowait(tB + NR + kbo * 16, x1);
                    tmem_ld16_nowait(tB + 2 * HALF + kbo * 16, x2);
                    tmem_ld16_nowait(tB + NR + 2 * HALF + kbo * 16, x3);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        bp[i] = __uint_as_float(x0[i]) + __uint_as_float(x1[i]);
                        bm[i] = __uint_as_float(x2[i]) + __uint_as_float(x3[i]);
                    }
                }
                const int r2_0 = tc * HALF + kbo * KB;
                if (!has || r2_0 >= d.nKpad) continue;
                // g[0..3][w] = ge.re, ge.im, go.re, go.im of K index r2_0 + w (reusing ap/am as storage would not save registers)
                float g[4][KB];
#pragma unroll
                for (int w = 0; w < KB; ++w) {
                    const int r2 = r2_0 + w;
                    const float2 p2 = d.tabf[(size_t)8 * d.Rfp + (r2 < d.nKfp ? r2 : 0)];
                    float xr, xi, tpr, tpi, tmr, tmi;
                    xr = ap[2 * w] - sd * sg * bp[2 * w + 1]; xi = ap[2 * w + 1] + sd * sg * bp[2 * w];
                    tpr = xr * pc - xi * ps; tpi = xr * ps + xi * pc;                 // T[k][j+]
                    xr = am[2 * w] - sd * sg * bm[2 * w + 1]; xi = am[2 * w + 1] + sd * sg * bm[2 * w];
                    tmr = xr * pc - xi * ps; tmi = xr * ps + xi * pc;                 // T[k][j-]
                    const float gpr = tpr * p2.x - tpi * p2.y, gpi = tpr * p2.y + tpi * p2.x;   // pre2 * T+
                    const float gmr = tmr * p2.x + tmi * p2.y, gmi = tmi * p2.x - tmr * p2.y;   // conj(pre2) * T-
                    float ger, gei, gor, goi;
                    if (d.ncR2 == 0 && r2 == 0) { ger = gpr; gei = gpi; gor = 0.f; goi = 0.f; }
                    else { ger = gpr + gmr; gei = gpi + gmi; gor = gpr - gmr; goi = gpi - gmi; }
                    if (r2 >= d.nKf) { ger = gei = gor = goi = 0.f; }
                    g[0][w] = ger; g[1][w] = gei; g[2][w] = gor; g[3][w] = goi;
                }
#pragma unroll
                for (int a4 = 0; a4 < 4; ++a4)          // planes ge_hi, ge_lo, go_hi, go_lo
#pragma unroll
                    for (int part = 0; part < 2; ++part) {
                        float v[KB], o[KB];
#pragma unroll
                        for (int w = 0; w < KB; ++w) {
                            const float x = g[(a4 >> 1) * 2 + part][w], h = tf32_hi(x);
                            v[w] = (a4 & 1) ? (x - h) : h;
                        }
#pragma unroll
                        for (int w = 0; w < KB; ++w) o[w] = swz ? v[w ^ 4] : v[w];
                        const int n = 2 * k + part;
                        unsigned char *dst = d.nB + ((long long)(n / NR) * (d.nKpad / KB) + r2_0 / KB) * B_BYTES + a4 * B_ARR
                                             + ((n % NR) >> 3) * SBO + ((n % NR) & 7) * 32;
                        st_global_v8(dst, o);
                    }
            }
        }
#ifdef LFD_TILE_TIMING
        if (g_c64_timing && warp == 1 && lane == 0) { long long *o = g_c64_timing + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16; o[10] = t_e1 - t_e0; o[11] = clock64() - t_e1; }
#endif
    }
#ifdef LFD_TILE_TIMING
    if (g_c64_timing && warp == 1 && lane == 0) { long long *o = g_c64_timing + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16; o[12] = clock64() - t_entry; }
#endif
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static inline size_t al(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
static inline int rup(int v, int m) { return (v + m - 1) / m * m; }

struct Geo {
    int Kf1, Kpad1, slots1, Npad1, Rf1, Rfp1, Kf2, Kpad2, Npad2, Rf2, Rfp2, nKfp;
};
static Geo geo(const lfd_mft_desc &p) {
    Geo g;
    g.Kf1 = (p.m + 1) / 2; g.Kpad1 = rup(g.Kf1, 32);
    g.Kf2 = (p.n + 1) / 2; g.Kpad2 = rup(g.Kf2, HALF);   // = what the row-stage tiles write (HALF is a multiple of KB)
    g.nKfp = rup(g.Kf2, 32);
    g.slots1 = rup(g.Kf2, HALF) * 2;          // per HALF folded columns: HALF (j+) + HALF (j-) slots = one tile
    g.Npad1 = 2 * g.slots1;
    g.Rf1 = (p.M + 1) / 2; g.Rfp1 = rup(g.Rf1, TM);
    g.Npad2 = 2 * rup(p.M, TN);
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
                   cudaStream_t stream, const lfd_pupil_src *src = nullptr, int intensity_out = 0) {
    if (count == 0) return 0;
    LFD_REQUIRE(descs && workspace, "lfd_mft_c64x3_batched: NULL argument");
    LFD_REQUIRE(count <= 32767, "lfd_mft_c64x3_batched: at most 32767 planes per call (got %d)", count);
    const size_t need = c64_workspace_bytes(descs, count);
    LFD_REQUIRE(workspace_bytes >= need, "lfd_mft_c64x3_batched: workspace too small (%zu < %zu)", workspace_bytes, need);
    int dev = 0;
    LFD_CUDA_OK(cudaGetDevice(&dev));
    if (ensure_dynamic_smem(dev, (const void *)mft_c64_kernel<true>, (int)SMEM_BYTES)) return 1;
    if (ensure_dynamic_smem(dev, (const void *)mft_c64_kernel<false>, (int)SMEM_BYTES)) return 1;
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
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        const bool src_ok = !src || (src[i].amp && src[i].opd && src[i].wavelength != 0.0 && src[i].r0 >= 0 && src[i].c0 >= 0 &&
                                     src[i].r0 + p.m <= src[i].n_r && src[i].c0 + p.n <= src[i].n_c);
        if (!(p.m > 0 && p.n > 0 && p.M > 0 && p.N > 0 && (src || (p.ldf >= p.n && p.f)) && p.ldo >= p.N && p.out && src_ok)) {
            free(h);
            LFD_REQUIRE(false, "lfd_mft_c64x3_batched: plane %d has invalid shape/ld/pointers", i);
        }
        const Geo g = geo(p);
        unsigned char *B1 = (unsigned char *)(ws + off); off += al((size_t)4 * g.Npad1 * g.Kpad1 * sizeof(float));
        unsigned char *B2 = (unsigned char *)(ws + off);
        const size_t b2_bytes = al((size_t)4 * g.Npad2 * g.Kpad2 * sizeof(float));
        off += b2_bytes;
        double2 *td1 = (double2 *)(ws + off); off += al((size_t)4 * g.Rfp1 * sizeof(double2));
        float2 *tf1 = (float2 *)(ws + off); off += al(((size_t)8 * g.Rfp1 + g.nKfp) * sizeof(float2));
        double2 *td2 = (double2 *)(ws + off); off += al((size_t)4 * g.Rfp2 * sizeof(double2));
        float2 *tf2 = (float2 *)(ws + off); off += al((size_t)8 * g.Rfp2 * sizeof(float2));
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
        if (src) {
            f.amp = src[i].amp; f.opd = src[i].opd; f.mask = src[i].mask;
            f.pld = src[i].n_c; f.pr0 = src[i].r0; f.pc0 = src[i].c0; f.wavelength = 1.0 / src[i].wavelength;
        }
        if (g.slots1 / TN > max_fs_x) max_fs_x = g.slots1 / TN;
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
        s2.out = (float2 *)p.out; s2.ldo = p.ldo; s2.nB = nullptr; s2.nplane = 0; s2.intensity = intensity_out;
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
    // every block of both pre-blocked operands is written in full (zeros in the K padding; the column padding of the
    // stage-2 operand only feeds output columns that are never stored), so the workspace needs no clearing
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

// Two executions of the complex64 mode: the folded 3xTF32 form on tcgen05 (this file) and the FP32 build of the chirp-z
// row transform (mft_czt.cu).  lfd_mft_desc.execution of the first plane, else the process default, decides:
// DIRECT / FOLDED -> tcgen05, CZT / AUTO -> chirp-z whenever every plane fits it (FFT length <= 8192).
namespace lfd {
int mft_resolve_execution(const lfd_mft_desc *descs, int count);
bool czt_supported(const lfd_mft_desc *descs, int count);
size_t czt_workspace_bytes(const lfd_mft_desc *descs, int count, bool c64);
int launch_mft_czt(const lfd_mft_desc *descs, int count, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                   const lfd_pupil_src *src, int intensity_out, bool c64);
static bool c64_runs_czt(const lfd_mft_desc *descs, int count) {
    return descs && count > 0 && mft_resolve_execution(descs, count) == LFD_MFT_CZT && czt_supported(descs, count);
}
}  // namespace lfd

extern "C" int lfd_mft_c64_execution(const lfd_mft_desc *descs, int count) {
    return lfd::c64_runs_czt(descs, count) ? LFD_MFT_CZT : LFD_MFT_FOLDED;
}

extern "C" size_t lfd_mft_c64x3_workspace_bytes(const lfd_mft_desc *descs, int count) {
    if (lfd::c64_runs_czt(descs, count)) return lfd::czt_workspace_bytes(descs, count, true);
    return lfd::c64::c64_workspace_bytes(descs, count);
}

extern "C" int lfd_mft_c64x3_batched(const lfd_mft_desc *descs, int count, void *workspace,
                                     size_t workspace_bytes, void *stream) {
    if (lfd::c64_runs_czt(descs, count))
        return lfd::launch_mft_czt(descs, count, workspace, workspace_bytes, (cudaStream_t)stream, nullptr, 0, true);
    return lfd::c64::launch_mft_c64(descs, count, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int lfd_mft_c64x3_from_pupil(const lfd_mft_desc *descs, const lfd_pupil_src *src, int count, int intensity_out,
                                        void *workspace, size_t workspace_bytes, void *stream) {
    LFD_REQUIRE(descs && src && workspace && count > 0, "lfd_mft_c64x3_from_pupil: bad arguments");
    if (lfd::c64_runs_czt(descs, count))
        return lfd::launch_mft_czt(descs, count, workspace, workspace_bytes, (cudaStream_t)stream, src, intensity_out, true);
    return lfd::c64::launch_mft_c64(descs, count, workspace, workspace_bytes, (cudaStream_t)stream, src, intensity_out);
}

#ifdef LFD_TILE_TIMING
extern "C" int lfd_debug_c64_trace(long long *buf_dev) {
    return (int)cudaMemcpyToSymbol(lfd::c64::g_c64_trace, &buf_dev, sizeof(buf_dev));
}
extern "C" int lfd_debug_c64_timing(long long *buf_dev, int stage) {
    cudaMemcpyToSymbol(lfd::c64::g_c64_timing_stage, &stage, sizeof(stage));
    return (int)cudaMemcpyToSymbol(lfd::c64::g_c64_timing_buf, &buf_dev, sizeof(buf_dev));
}
#endif
