// K2a, chirp-z execution (LFD_MFT_CZT) — the same matrix Fourier transform, lentil/fourier.py:5-121,
//     F = E1 f E2 sqrt|a_r a_c|,  E1[u,i] = exp(sgn 2 pi i a_r (R_i + o_r)(U_u - s_r)),
// evaluated as two passes of length-L convolutions instead of two GEMMs.  With R' = R + o, U' = U - s and
// D = U' - R' (D depends on u - i only),  R'U' = (R'^2 + U'^2 - D^2) / 2, hence along one axis
//     y[u] = post[u] * sum_i (pre[i] x[i]) h[u - i],   pre = cis(sgn pi a R'^2), post = cis(sgn pi a U'^2), h = cis(-sgn pi a D^2)
// (Bluestein / chirp-z).  The convolution runs per row as FFT_L -> x H -> IFFT_L in shared memory, L the power of two
// >= n_in + n_out - 1 (2048 for the 1001 -> 1024 bench shape, 8192 for the 4081 -> 2048 planes of BASELINE configs[4]),
// H = FFT_L(h) built once per (plane, axis).  One plane of 1001^2 -> 1024^2 is 2025 row transforms ~ 0.5 GFLOP of FP64
// instead of the 4.15 GFLOP the folded DMMA kernel executes (16.6 GFLOP algorithmic); exact for any alpha / shift /
// offset / parity like the other executions.
//
//  * every chirp phase is formed in cycles with error-free products and reduced exactly before sincospi (cis_cycles),
//    the FFT twiddles come from per-pass tables of exactly reduced sincospi values: parity vs the oracle ~1e-14.
//  * FFT: Stockham auto-sort in shared memory, RADIX-16 passes, L/16 threads per row, one radix-16 butterfly per thread
//    and pass: log2 L = 4 a + b  ->  a radix-16 passes and one pass of radix 2^b (b = 1..4) in the middle of the row (the
//    "turn").  Whatever the radix, thread t owns the elements t + s L/16 (s = 0..15) on the natural-order side of every
//    pass, so a pass can run in place in ONE buffer (read own 16 -> barrier -> scatter) or ping-pong between two.
//  * the inverse transform is run as the adjoints of the forward passes in reverse order, so the last forward pass, the
//    product with H and the first inverse pass of a thread happen in registers (turn); the first forward pass loads
//    straight from global memory and the last inverse pass stores straight to it.
//    L = 2048: 8 shared-memory sweeps per row (4 written, 4 read) where the radix-8 form of round 1 needed 12.
//  * ROWS rows of a plane are transformed by one CTA, interleaved in the lane index (lane % ROWS = row): twiddles,
//    chirps and H are loaded once per ROWS rows (same address across those lanes), and the transposed store of stage A /
//    the column store of stage B write ROWS adjacent elements (32-byte sectors are filled from ROWS = 2 on).
//    Shared-memory element i of row c lives at (i + i / 16) * ROWS + c: one element of padding per 16 makes the
//    stride-16 scatter of the first pass conflict-free, every other access of a quarter-warp is contiguous.
//  * stage A transforms the rows of f (K1 can be fused: the phasor amp*mask*exp(2 pi i opd / lambda) is formed in the
//    load, lentil/plane.py:502-507) and stores its result transposed, stage B transforms the rows of that (= the
//    columns of the plane) and stores F, or |F|^2 as float64 when the caller only wants intensities.
//  * L <= 8192: one buffer of L complex128 (139 KB with padding) is the most that fits the 227 KB of an SM; longer
//    transforms use the folded DMMA execution (the dispatcher in mft_c128.cu decides).
#include "lfd_common.cuh"
#include <mutex>

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
__device__ double2 g_tw[MAX_LOG2L - MIN_LOG2L + 1][TW_PER_LEN];

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
#ifndef LFD_CZT_PPRE
#define LFD_CZT_PPRE 4        // post-chirp factors prefetched before the last pass (0 .. 8)
#endif
__host__ __device__ constexpr int rows_for(int lg) {
    return LFD_CZT_ELEMS ? ((1 << lg) >= LFD_CZT_ELEMS ? 1 : LFD_CZT_ELEMS >> lg) : (lg <= 10 ? 2048 >> lg : (lg == 11 ? 2 : 1));
}
__host__ __device__ constexpr int nbuf_for(int lg) { return ((size_t)LFD_CZT_NBUF * rows_for(lg) * ((1 << lg) + (1 << lg) / 16) * 16 > 200 * 1024) ? 1 : LFD_CZT_NBUF; }

__global__ void roots_kernel() {
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
            g_tw[blockIdx.y][e] = make_double2(c, s);
        }
    }
}

struct Plane {
    const double2 *f; long long ldf;
    double2 *Gt;                  // stage A result, transposed: N x mpad
    void *out; long long ldo;
    int m, n, M, N, mpad, logLA, logLB, intensity;
    double2 *preA, *postA, *HA;   // axis 1 (n -> N): pre[n], post[N], H[LA]
    double2 *preB, *postB, *HB;   // axis 0 (m -> M)
    double alpha_r, alpha_c, x0r, y0r, x0c, y0c, sgn, scale;
    // fused pupil prep: when amp != NULL, f(i, c) = amp * mask * exp(+2 pi i opd / lambda) at pupil pixel (pr0 + i, pc0 + c)
    const double *amp, *opd;
    const unsigned char *mask;
    long long pld;
    int pr0, pc0;
    double wavelength;
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i (S = +1: forward transform) or by +i (S = -1: inverse)
template <int S> __device__ __forceinline__ double2 mul_mi(double2 a) { return S > 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x); }
// multiply by the constant (wr, -S wi): a root of unity of the forward (S = +1) or inverse transform
template <int S> __device__ __forceinline__ double2 mul_root(double2 a, double wr, double wi) {
    return S > 0 ? make_double2(a.x * wr + a.y * wi, a.y * wr - a.x * wi) : make_double2(a.x * wr - a.y * wi, a.y * wr + a.x * wi);
}

template <int S> __device__ __forceinline__ void dft2p(double2 &x0, double2 &x1) {
    const double2 s = cadd(x0, x1), d = csub(x0, x1);
    x0 = s; x1 = d;
}
template <int S> __device__ __forceinline__ void dft4(double2 &x0, double2 &x1, double2 &x2, double2 &x3) {
    const double2 s0 = cadd(x0, x2), s1 = csub(x0, x2), s2 = cadd(x1, x3), s3 = mul_mi<S>(csub(x1, x3));
    x0 = cadd(s0, s2); x2 = csub(s0, s2); x1 = cadd(s1, s3); x3 = csub(s1, s3);
}
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
// 16-point DFT, natural order in and out, as 4 x 4: n = 4 n1 + n2, k = k1 + 4 k2,
//   X[k1 + 4 k2] = sum_n2 w4^(n2 k2) [ w16^(n2 k1) sum_n1 x[4 n1 + n2] w4^(n1 k1) ]
template <int S> __device__ __forceinline__ void dft16(double2 (&v)[16]) {
    const double h = 0.70710678118654752440, c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<S>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);      // -> y[n2][k1] at v[4 k1 + n2]
    v[5]  = mul_root<S>(v[5], c1, s1);        // k1 = 1: w16^1, w16^2, w16^3
    v[6]  = mul_root<S>(v[6], h, h);
    v[7]  = mul_root<S>(v[7], s1, c1);
    v[9]  = mul_root<S>(v[9], h, h);          // k1 = 2: w16^2, w16^4, w16^6
    v[10] = mul_mi<S>(v[10]);
    v[11] = mul_root<S>(v[11], -h, h);
    v[13] = mul_root<S>(v[13], s1, c1);       // k1 = 3: w16^3, w16^6, w16^9
    v[14] = mul_root<S>(v[14], -h, h);
    v[15] = mul_root<S>(v[15], -c1, -s1);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<S>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // X[k1 + 4 k2] at v[4 k1 + k2]
    // 4 x 4 transpose of the register names
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b) { const double2 x = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = x; }
}
template <int R, int S> __device__ __forceinline__ void dftR(double2 (&v)[R]) {
    if constexpr (R == 16) dft16<S>(v);
    else if constexpr (R == 8) dft8<S>(v);
    else if constexpr (R == 4) dft4<S>(v[0], v[1], v[2], v[3]);
    else dft2p<S>(v[0], v[1]);
}

// v[r] *= w^r (CONJ: conj(w)^r), r = 1 .. R - 1: four interleaved chains of powers, each stepping by w^4
template <int R, bool CONJ> __device__ __forceinline__ void twiddle_powers(double2 (&v)[R], double2 w1) {
    if (CONJ) w1.y = -w1.y;
    v[1] = cmul(v[1], w1);
    if constexpr (R >= 4) {
        const double2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        if constexpr (R >= 8) {
            const double2 w4 = cmul(w2, w2);
            double2 q0 = w4, q1 = w1, q2 = w2, q3 = w3;
#pragma unroll
            for (int a = 4; a < R; a += 4) {
                q1 = cmul(q1, w4); q2 = cmul(q2, w4); q3 = cmul(q3, w4);
                v[a] = cmul(v[a], q0); v[a + 1] = cmul(v[a + 1], q1); v[a + 2] = cmul(v[a + 2], q2); v[a + 3] = cmul(v[a + 3], q3);
                if (a + 4 < R) q0 = cmul(q0, w4);
            }
        }
    }
}

// shared-memory slot of element i (per row; rows are interleaved with stride ROWS)
template <int ROWS> __device__ __forceinline__ int slot(int i) { return (i + (i >> 4)) * ROWS; }

// last forward pass, product with H and first adjoint pass of one butterfly of the turn, in registers and in place;
// hv[r] = H[j + r L/RT] for this butterfly (j = t + q L/16), fetched by the caller before the barrier
template <int LOG2L, int ROWS>
__device__ __forceinline__ void turn(double2 *X, const double2 *hv, int t, int q, double2 w1) {
    constexpr int L = 1 << LOG2L, T = L / 16, RT = turn_radix(LOG2L), NB = 16 / RT;
    double2 v[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) v[r] = X[slot<ROWS>(t + (q + r * NB) * T)];
    twiddle_powers<RT, false>(v, w1);
    dftR<RT, 1>(v);
#pragma unroll
    for (int r = 0; r < RT; ++r) v[r] = cmul(v[r], hv[r]);
    dftR<RT, -1>(v);
    twiddle_powers<RT, true>(v, w1);
#pragma unroll
    for (int r = 0; r < RT; ++r) X[slot<ROWS>(t + (q + r * NB) * T)] = v[r];
}

// One row (per thread: its 16 elements of one row) of the chirp-z convolution y = IFFT(FFT(x) * H).  `load16(v)` fills
// v[s] with x[t + s L/16] (global memory), y leaves through `store16(v)` (v[s] = y[t + s L/16]).  X / Y are this
// thread's row base pointers in the two buffers (the same buffer when NBUF == 1); with two buffers the roles alternate
// from row to row, so the next row's first scatter never meets this row's last reads.  Everything a pass needs from
// global memory — its twiddle, the 16 values of H for the turn, the first eight post-chirp factors for the last pass — is
// fetched BEFORE the barrier that precedes the pass: the data registers are dead there (the row lives in shared memory),
// and the load latency overlaps the barrier wait instead of the arithmetic.
template <int LOG2L, int ROWS, int NBUF, class Load16, class Store16>
__device__ __forceinline__ void czt_row(double2 *&X, double2 *&Y, int t, const Load16 &load16, const double2 *__restrict__ H,
                                        const double2 *__restrict__ post, int nout, const Store16 &store16) {
    constexpr int L = 1 << LOG2L, T = L / 16, NREG = nreg(LOG2L), RT = turn_radix(LOG2L), NB = 16 / RT, NS = L / RT;
    const double2 *__restrict__ tw = g_tw[LOG2L - MIN_LOG2L];
    double2 v[16];
    load16(v);
    dft16<1>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) X[(17 * t + r) * ROWS] = v[r];          // = slot(16 t + r)
    int Ns = 16;
#pragma unroll
    for (int p = 1; p < NREG; ++p) {
        const int k = t & (Ns - 1), j0 = (t - k) * 16 + k;
        const double2 w1 = tw[tw_offset(p) + k];
        __syncthreads();
#pragma unroll
        for (int s = 0; s < 16; ++s) v[s] = X[slot<ROWS>(t + s * T)];
        twiddle_powers<16, false>(v, w1);
        dft16<1>(v);
        if (NBUF == 1) __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) Y[slot<ROWS>(j0 + r * Ns)] = v[r];
        if (NBUF == 2) { double2 *x = X; X = Y; Y = x; }
        Ns *= 16;
    }
    {
        // the first HN of this thread's 16 values of H before the barrier, the others once the turn is under way
        constexpr int HN = LOG2L <= LFD_CZT_PRE_MAXLG ? LFD_CZT_HPRE : 0;
        double2 wt[NB], hv[16];
#pragma unroll
        for (int q = 0; q < NB; ++q) wt[q] = tw[tw_offset(NREG) + t + q * T];
#pragma unroll
        for (int e = 0; e < HN; ++e) hv[e] = H[t + (e / RT) * T + (e % RT) * NS];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < NB; ++q) {
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (q * RT + r >= HN) hv[q * RT + r] = H[t + q * T + r * NS];
            turn<LOG2L, ROWS>(X, hv + q * RT, t, q, wt[q]);
        }
    }
#pragma unroll
    for (int p = NREG - 1; p >= 1; --p) {
        Ns /= 16;
        const int k = t & (Ns - 1), j0 = (t - k) * 16 + k;
        const double2 w1 = tw[tw_offset(p) + k];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = X[slot<ROWS>(j0 + r * Ns)];
        dft16<-1>(v);
        twiddle_powers<16, true>(v, w1);
        if (NBUF == 1) __syncthreads();
#pragma unroll
        for (int s = 0; s < 16; ++s) Y[slot<ROWS>(t + s * T)] = v[s];
        if (NBUF == 2) { double2 *x = X; X = Y; Y = x; }
    }
    constexpr int PP = LOG2L <= LFD_CZT_PRE_MAXLG ? LFD_CZT_PPRE : 0;   // post-chirp factors fetched before the barrier (outputs t + s L/16, s < PP)
    double2 pv[PP > 0 ? PP : 1];
#pragma unroll
    for (int sI = 0; sI < PP; ++sI) pv[sI] = (t + sI * T < nout) ? post[t + sI * T] : make_double2(0.0, 0.0);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = X[(17 * t + r) * ROWS];
    dft16<-1>(v);
#pragma unroll
    for (int sI = 0; sI < PP; ++sI) v[sI] = cmul(v[sI], pv[sI]);
#pragma unroll
    for (int sI = PP; sI < 8; ++sI) v[sI] = (t + sI * T < nout) ? cmul(v[sI], post[t + sI * T]) : v[sI];
    if (8 * T < nout) {                                    // uniform; outputs beyond half the transform length are rare
#pragma unroll
        for (int sI = 8; sI < 16; ++sI) v[sI] = (t + sI * T < nout) ? cmul(v[sI], post[t + sI * T]) : v[sI];
    }
    store16(v);
    if (NBUF == 1) __syncthreads();
    else { double2 *x = X; X = Y; Y = x; }
}

// ---- per (plane, axis): pre / post chirps and the transformed chirp filter H = FFT_L(h) -----------------
template <int LOG2L>
__global__ void __launch_bounds__((1 << LOG2L) / 16)
czt_tables_kernel(const Plane *__restrict__ descs) {
    constexpr int L = 1 << LOG2L, T = L / 16, NREG = nreg(LOG2L), RT = turn_radix(LOG2L), NB = 16 / RT, NS = L / RT;
    extern __shared__ double2 sm[];
    const Plane &d = descs[blockIdx.x];
    const bool axisA = blockIdx.y == 0;
    if ((axisA ? d.logLA : d.logLB) != LOG2L) return;
    const int nin = axisA ? d.n : d.m, nout = axisA ? d.N : d.M;
    const double alpha = axisA ? d.alpha_c : d.alpha_r, x0 = axisA ? d.x0c : d.x0r, y0 = axisA ? d.y0c : d.y0r;
    double2 *pre = axisA ? d.preA : d.preB, *post = axisA ? d.postA : d.postB, *H = axisA ? d.HA : d.HB;
    const double post_scale = (axisA ? 1.0 : d.scale) / (double)L;
    const int t = threadIdx.x;
    double c, s;
    for (int i = t; i < nin; i += T) {
        const double r = (double)i + x0;
        cis_cycles(alpha, r, 0.5 * r, d.sgn, c, s);
        pre[i] = make_double2(c, s);
    }
    for (int u = t; u < nout; u += T) {
        const double v = (double)u + y0;
        cis_cycles(alpha, v, 0.5 * v, d.sgn, c, s);
        post[u] = make_double2(c * post_scale, s * post_scale);
    }
    const double dd = y0 - x0;                       // D = (u - i) + (y0 - x0)
    const double sg = d.sgn;
    // plain forward FFT of the chirp: circular position q holds the lag p = u - i (p = q for q < nout, p = q - L for the
    // negative lags); one buffer, in place
    double2 *X = sm;
    const double2 *__restrict__ tw = g_tw[LOG2L - MIN_LOG2L];
    double2 v[16];
#pragma unroll
    for (int sI = 0; sI < 16; ++sI) {
        const int q = t + sI * T;
        const int p = q < nout ? q : q - L;
        v[sI] = make_double2(0.0, 0.0);
        if (p > -nin && p < nout) {
            const double D = (double)p + dd;
            double cc, ss;
            cis_cycles(alpha, D, 0.5 * D, -sg, cc, ss);
            v[sI] = make_double2(cc, ss);
        }
    }
    dft16<1>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) X[17 * t + r] = v[r];
    int Ns = 16;
#pragma unroll
    for (int p = 1; p < NREG; ++p) {
        const int k = t & (Ns - 1), j0 = (t - k) * 16 + k;
        const double2 w1 = tw[tw_offset(p) + k];
        __syncthreads();
#pragma unroll
        for (int sI = 0; sI < 16; ++sI) v[sI] = X[slot<1>(t + sI * T)];
        twiddle_powers<16, false>(v, w1);
        dft16<1>(v);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) X[slot<1>(j0 + r * Ns)] = v[r];
        Ns *= 16;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const int j = t + q * T;
        double2 u[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) u[r] = X[slot<1>(t + (q + r * NB) * T)];
        twiddle_powers<RT, false>(u, tw[tw_offset(NREG) + j]);
        dftR<RT, 1>(u);
#pragma unroll
        for (int r = 0; r < RT; ++r) H[j + r * NS] = u[r];
    }
}

// ---- one stage: every row of every plane whose FFT length is L ----------------------------------------------
// STAGE_A: row i of f (n elements, or the fused phasor) -> Gt[:, i] (N outputs, transposed store)
// else   : row v of Gt (m elements)                     -> out[:, v] (M outputs; complex128 or |.|^2 float64)
// A CTA takes ROWS consecutive rows at a time (lane % ROWS = row).  Register budget: 128 per thread (512 threads per SM).
__host__ __device__ constexpr int cta_threads(int lg) { return ((1 << lg) / 16) * rows_for(lg); }
__host__ __device__ constexpr int min_ctas(int lg) { return cta_threads(lg) >= 512 ? 1 : 512 / cta_threads(lg); }

// Work units (ROWS rows of one plane) are numbered plane-major over the planes of THIS length; starts[p] = units in planes
// < p (count + 1 entries).  Units are dealt round-robin: at any moment the CTAs of the grid work on ADJACENT rows, so the
// 16-byte pieces they scatter into the transposed intermediate (stage A) or the output columns (stage B) complete their
// 128-byte lines in L2 within one unit time.  (A contiguous run of units per CTA keeps a plane's tables in L1 but leaves
// N x CTAs partially written lines in flight — 77 MB for 2048-point planes — which L2 evicts half filled: measured
// 1.6x slower at 2001^2 -> 2048^2, LFD_CZT_CONTIG=1.)
template <int LOG2L, bool STAGE_A>
__global__ void __launch_bounds__(cta_threads(LOG2L), min_ctas(LOG2L))
czt_stage_kernel(const Plane *__restrict__ descs, const int *__restrict__ starts, int count) {
    constexpr int L = 1 << LOG2L, T = L / 16, ROWS = rows_for(LOG2L), NBUF = nbuf_for(LOG2L), NT = T * ROWS;
    extern __shared__ double2 sm[];
    const int c = threadIdx.x % ROWS, t = threadIdx.x / ROWS;
    double2 *X = sm + c, *Y = sm + (NBUF - 1) * ROWS * (L + L / 16) + c;
    const int total = starts[count];
#if LFD_CZT_CONTIG
    const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x, stride = 1;
    const int w0 = (int)blockIdx.x * per, w1 = min(total, w0 + per);
#else
    const int stride = (int)gridDim.x, w0 = (int)blockIdx.x, w1 = total;
#endif
    if (w0 >= w1) return;
    __shared__ Plane sd;                 // the plane this CTA is working on
    int plane = 0, cur = -1;
    {   // last plane whose first unit is <= w0 (planes of another length own no units: their start equals the next one's)
        int lo = 0, hi = count;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (starts[mid] <= w0) lo = mid; else hi = mid; }
        plane = lo;
    }
    int pbeg = starts[plane], pend = starts[plane + 1];
    for (int w = w0; w < w1; w += stride) {
        while (w >= pend) { ++plane; pbeg = pend; pend = starts[plane + 1]; }
        const int row0 = (w - pbeg) * ROWS;
        if (plane != cur) {              // uniform over the CTA
            __syncthreads();
            const unsigned long long *g = reinterpret_cast<const unsigned long long *>(descs + plane);
            unsigned long long *sdw = reinterpret_cast<unsigned long long *>(&sd);
            for (int i = threadIdx.x; i < (int)(sizeof(Plane) / 8); i += NT) sdw[i] = g[i];
            __syncthreads();
            cur = plane;
        }
        const Plane &d = sd;
        const int nrows = STAGE_A ? d.m : d.N;
        const int row = row0 + c;
        const bool rv = row < nrows;
        const int nin = STAGE_A ? d.n : d.m, nout = STAGE_A ? d.N : d.M;
        const double2 *__restrict__ pre = STAGE_A ? d.preA : d.preB;
        const double2 *__restrict__ post = STAGE_A ? d.postA : d.postB;
        const double2 *__restrict__ H = STAGE_A ? d.HA : d.HB;
        const Plane *dp = &d;
        if (LFD_CZT_L2PRE && LOG2L <= LFD_CZT_PRE_MAXLG && w + stride < w1 && w + stride < pend && row + stride * ROWS < nrows) {   // next unit in the same plane: pull its input rows into L2 now
            const long long nrow = row + stride * ROWS;
#pragma unroll
            for (int sI = 0; sI < 16; sI += 2) {                         // one prefetch per 32-byte sector pair of this thread's elements
                const int i = t + sI * T;
                if (i < nin) {
                    if (STAGE_A && dp->amp != nullptr) {
                        const long long e = (nrow + dp->pr0) * dp->pld + dp->pc0 + i;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(dp->amp + e));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(dp->opd + e));
                    } else {
                        const double2 *nsrc = STAGE_A ? dp->f + nrow * dp->ldf : dp->Gt + nrow * dp->mpad;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + i));
                    }
                }
            }
        }
        // first forward pass: the inputs of this thread straight from global memory (x pre-chirp; zero beyond the input
        // length), in two halves of eight so that every load of a half is issued before its arithmetic
        auto load = [=](double2 (&v)[16]) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                if (hf * 8 * T >= nin) {                                 // uniform: this half lies beyond the input
#pragma unroll
                    for (int s = 0; s < 8; ++s) v[hf * 8 + s] = make_double2(0.0, 0.0);
                    continue;
                }
                double2 pr[8];
                if (STAGE_A && dp->amp != nullptr) {
                    double am[8], op[8];
                    const long long base = (long long)(dp->pr0 + row) * dp->pld + dp->pc0;
                    const unsigned char *mk = dp->mask;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const int i = t + (hf * 8 + s) * T;
                        const bool in = rv && i < nin;
                        am[s] = in ? dp->amp[base + i] : 0.0;
                        op[s] = in ? dp->opd[base + i] : 0.0;
                        pr[s] = in ? pre[i] : make_double2(0.0, 0.0);
                        if (in && mk != nullptr && mk[base + i] == 0) am[s] = 0.0;
                    }
                    const double lam = dp->wavelength;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        double2 x = make_double2(0.0, 0.0);
                        if (am[s] != 0.0) {                        // same arithmetic as K1 (pupil_prep.cu): phase in cycles, reduced exactly
                            const double tcyc = op[s] / lam;
                            double sn, cs;
                            sincospi(2.0 * (tcyc - rint(tcyc)), &sn, &cs);
                            x = make_double2(am[s] * cs, am[s] * sn);
                        }
                        v[hf * 8 + s] = cmul(x, pr[s]);
                    }
                } else {
                    const double2 *src = STAGE_A ? dp->f + (long long)row * dp->ldf : dp->Gt + (long long)row * dp->mpad;
                    double2 x[8];
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        const int i = t + (hf * 8 + s) * T;
                        const bool in = rv && i < nin;
                        x[s] = in ? src[i] : make_double2(0.0, 0.0);
                        pr[s] = in ? pre[i] : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int s = 0; s < 8; ++s) v[hf * 8 + s] = cmul(x[s], pr[s]);
                }
            }
        };
        if (STAGE_A) {
            double2 *Gt = dp->Gt; const long long mpad = dp->mpad;
            auto store = [=](double2 (&v)[16]) {
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    const int i = t + s * T;
                    if (rv && i < nout) Gt[(long long)i * mpad + row] = v[s];
                }
            };
            czt_row<LOG2L, ROWS, NBUF>(X, Y, t, load, H, post, nout, store);
        } else if (dp->intensity) {
            double *out = (double *)dp->out; const long long ldo = dp->ldo;
            auto store = [=](double2 (&v)[16]) {
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    const int i = t + s * T;
                    if (rv && i < nout) out[(long long)i * ldo + row] = v[s].x * v[s].x + v[s].y * v[s].y;
                }
            };
            czt_row<LOG2L, ROWS, NBUF>(X, Y, t, load, H, post, nout, store);
        } else {
            double2 *out = (double2 *)dp->out; const long long ldo = dp->ldo;
            auto store = [=](double2 (&v)[16]) {
#pragma unroll
                for (int s = 0; s < 16; ++s) {
                    const int i = t + s * T;
                    if (rv && i < nout) out[(long long)i * ldo + row] = v[s];
                }
            };
            czt_row<LOG2L, ROWS, NBUF>(X, Y, t, load, H, post, nout, store);
        }
    }
}

static inline size_t al(size_t v) { return (v + 255) / 256 * 256; }
static inline int log2_len(int nin, int nout) {
    int lg = MIN_LOG2L;
    while ((1LL << lg) < (long long)nin + nout - 1) ++lg;
    return lg;
}
static inline int pad_rows(int m) { return (m + 7) & ~7; }
static inline int rows_for_length(int lg) { return rows_for(lg); }
constexpr int NLEN = MAX_LOG2L - MIN_LOG2L + 1;
// workspace header: the plane descriptors, then per (stage, length) the unit start table of count + 1 ints
static inline size_t header_bytes(int count) { return al((size_t)count * sizeof(Plane) + (size_t)2 * NLEN * (count + 1) * sizeof(int)); }
static inline size_t starts_offset(int count, int stage, int lg) {
    return (size_t)count * sizeof(Plane) + ((size_t)stage * NLEN + (lg - MIN_LOG2L)) * (count + 1) * sizeof(int);
}

}  // namespace czt

using namespace czt;

// true when every plane of the batch fits the shared-memory FFT (both axes)
bool czt_supported(const lfd_mft_desc *descs, int count) {
    for (int i = 0; i < count; ++i)
        if (log2_len(descs[i].n, descs[i].N) > MAX_LOG2L || log2_len(descs[i].m, descs[i].M) > MAX_LOG2L) return false;
    return true;
}

// LFD_MFT_AUTO: chirp-z wherever it can run (measured against the folded DMMA form: DESIGN.md section 4).
bool czt_preferred(const lfd_mft_desc *descs, int count) { return czt_supported(descs, count); }

size_t czt_workspace_bytes(const lfd_mft_desc *descs, int count) {
    size_t bytes = header_bytes(count);
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        bytes += al((size_t)p.N * pad_rows(p.m) * sizeof(double2));
        bytes += al(((size_t)p.n + p.N + ((size_t)1 << log2_len(p.n, p.N))) * sizeof(double2));
        bytes += al(((size_t)p.m + p.M + ((size_t)1 << log2_len(p.m, p.M))) * sizeof(double2));
    }
    return bytes;
}

template <int LOG2L>
static int launch_for_length(const Plane *dd, int count, const int *starts_a, int total_a, const int *starts_b, int total_b,
                             int phase, int dev, int nsm, cudaStream_t stream) {
    constexpr int L = 1 << LOG2L, ROWS = rows_for(LOG2L), NBUF = nbuf_for(LOG2L), NT = cta_threads(LOG2L);
    const int smem_tab = (L + L / 16) * (int)sizeof(double2);
    const int smem = NBUF * ROWS * smem_tab;
    if (phase == 0) {
        if (ensure_dynamic_smem(dev, (const void *)czt_tables_kernel<LOG2L>, smem_tab)) return 1;
        czt_tables_kernel<LOG2L><<<dim3(count, 2), L / 16, smem_tab, stream>>>(dd);
        LFD_CUDA_OK(cudaGetLastError());
        count_launch();
        return 0;
    }
    int occ = 1;
    if (phase == 1 && total_a > 0) {
        if (ensure_dynamic_smem(dev, (const void *)czt_stage_kernel<LOG2L, true>, smem)) return 1;
        LFD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, czt_stage_kernel<LOG2L, true>, NT, smem));
        const int grid = total_a < nsm * occ ? total_a : nsm * occ;
        czt_stage_kernel<LOG2L, true><<<grid, NT, smem, stream>>>(dd, starts_a, count);
        LFD_CUDA_OK(cudaGetLastError());
        count_launch();
    }
    if (phase == 2 && total_b > 0) {
        if (ensure_dynamic_smem(dev, (const void *)czt_stage_kernel<LOG2L, false>, smem)) return 1;
        LFD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, czt_stage_kernel<LOG2L, false>, NT, smem));
        const int grid = total_b < nsm * occ ? total_b : nsm * occ;
        czt_stage_kernel<LOG2L, false><<<grid, NT, smem, stream>>>(dd, starts_b, count);
        LFD_CUDA_OK(cudaGetLastError());
        count_launch();
    }
    (void)ROWS;
    return 0;
}

int launch_mft_czt(const lfd_mft_desc *descs, int count, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                   const lfd_pupil_src *src, int intensity_out) {
    if (count == 0) return 0;
    LFD_REQUIRE(descs && workspace, "lfd_mft_c128 (chirp-z): NULL argument");
    LFD_REQUIRE(czt_supported(descs, count), "lfd_mft_c128 (chirp-z): a plane needs an FFT longer than %d", 1 << MAX_LOG2L);
    const size_t need = czt_workspace_bytes(descs, count);
    LFD_REQUIRE(workspace_bytes >= need, "lfd_mft_c128 (chirp-z): workspace too small (%zu < %zu)", workspace_bytes, need);

    int dev = 0, nsm = 0;
    LFD_CUDA_OK(cudaGetDevice(&dev));
    nsm = sm_count(dev);
    LFD_REQUIRE(nsm > 0, "lfd_mft_c128 (chirp-z): cannot query device %d", dev);
    {   // FFT roots: once per device and process
        static std::mutex mu;
        static bool ready[64] = {false};
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 64 || !ready[dev]) {
            roots_kernel<<<dim3((TW_PER_LEN + 255) / 256, MAX_LOG2L - MIN_LOG2L + 1), 256, 0, stream>>>();
            LFD_CUDA_OK(cudaGetLastError());
            LFD_CUDA_OK(cudaStreamSynchronize(stream));
            count_launch();
            if (dev < 64) ready[dev] = true;
        }
    }

    const size_t hdr = header_bytes(count);
    char *hbuf = (char *)calloc(hdr, 1);
    LFD_REQUIRE(hbuf != nullptr, "out of host memory");
    Plane *h = (Plane *)hbuf;
    char *ws = (char *)workspace;
    size_t off = hdr;
    bool useA[MAX_LOG2L + 1] = {false}, useB[MAX_LOG2L + 1] = {false};
    long long unitsA[MAX_LOG2L + 1] = {0}, unitsB[MAX_LOG2L + 1] = {0};
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        const bool ok = p.m > 0 && p.n > 0 && p.M > 0 && p.N > 0 && p.ldo >= p.N && p.out &&
                        (src ? (src[i].amp && src[i].opd && src[i].wavelength != 0.0 && src[i].r0 >= 0 && src[i].c0 >= 0 &&
                                src[i].r0 + p.m <= src[i].n_r && src[i].c0 + p.n <= src[i].n_c)
                             : (p.f && p.ldf >= p.n));
        if (!ok) {
            free(hbuf);
            LFD_REQUIRE(false, "lfd_mft_c128 (chirp-z): plane %d has invalid shape/ld/pointers", i);
        }
        Plane &d = h[i];
        d.f = (const double2 *)p.f; d.ldf = p.ldf; d.out = p.out; d.ldo = p.ldo;
        d.m = p.m; d.n = p.n; d.M = p.M; d.N = p.N; d.mpad = pad_rows(p.m);
        d.logLA = log2_len(p.n, p.N); d.logLB = log2_len(p.m, p.M); d.intensity = intensity_out;
        d.Gt = (double2 *)(ws + off); off += al((size_t)p.N * d.mpad * sizeof(double2));
        const size_t LA = (size_t)1 << d.logLA, LB = (size_t)1 << d.logLB;
        d.preA = (double2 *)(ws + off); d.postA = d.preA + p.n; d.HA = d.postA + p.N;
        off += al(((size_t)p.n + p.N + LA) * sizeof(double2));
        d.preB = (double2 *)(ws + off); d.postB = d.preB + p.m; d.HB = d.postB + p.M;
        off += al(((size_t)p.m + p.M + LB) * sizeof(double2));
        d.alpha_r = p.alpha_r; d.alpha_c = p.alpha_c;
        d.x0r = -floor(p.m / 2.0) + p.off_r; d.y0r = -floor(p.M / 2.0) - p.shift_r;
        d.x0c = -floor(p.n / 2.0) + p.off_c; d.y0c = -floor(p.N / 2.0) - p.shift_c;
        d.sgn = p.inverse ? 1.0 : -1.0;
        d.scale = p.unitary ? sqrt(fabs(p.alpha_r * p.alpha_c)) : 1.0;
        if (p.inverse) d.scale /= ((double)p.m * (double)p.n);
        if (src) {
            d.amp = src[i].amp; d.opd = src[i].opd; d.mask = src[i].mask;
            d.pld = src[i].n_c; d.pr0 = src[i].r0; d.pc0 = src[i].c0; d.wavelength = src[i].wavelength;
        }
        useA[d.logLA] = true; useB[d.logLB] = true;
        // unit start tables: stage A deals the m input rows of the plane, stage B the N columns, ROWS at a time
        for (int lg = MIN_LOG2L; lg <= MAX_LOG2L; ++lg) {
            int *sa = (int *)(hbuf + starts_offset(count, 0, lg)), *sb = (int *)(hbuf + starts_offset(count, 1, lg));
            sa[i] = (int)unitsA[lg]; sb[i] = (int)unitsB[lg];
        }
        const int ra = rows_for_length(d.logLA), rb = rows_for_length(d.logLB);
        unitsA[d.logLA] += (p.m + ra - 1) / ra;
        unitsB[d.logLB] += (p.N + rb - 1) / rb;
    }
    bool too_many = false;
    for (int lg = MIN_LOG2L; lg <= MAX_LOG2L; ++lg) {
        ((int *)(hbuf + starts_offset(count, 0, lg)))[count] = (int)unitsA[lg];
        ((int *)(hbuf + starts_offset(count, 1, lg)))[count] = (int)unitsB[lg];
        too_many |= unitsA[lg] > 0x7fffffffLL || unitsB[lg] > 0x7fffffffLL;
    }
    if (too_many) {
        free(hbuf);
        LFD_REQUIRE(false, "lfd_mft_c128 (chirp-z): too many rows in one batch");
    }
    cudaError_t e = cudaMemcpyAsync(workspace, hbuf, hdr, cudaMemcpyHostToDevice, stream);
    free(hbuf);
    LFD_CUDA_OK(e);
    const Plane *dd = (const Plane *)workspace;
    for (int phase = 0; phase < 3; ++phase) {
        for (int lg = MIN_LOG2L; lg <= MAX_LOG2L; ++lg) {
            if (!(useA[lg] || useB[lg])) continue;
            int rc = 0;
            switch (lg) {
#define LFD_CZT_CASE(LG) case LG: rc = launch_for_length<LG>(dd, count, (const int *)(ws + starts_offset(count, 0, LG)), (int)unitsA[LG], \
                                                             (const int *)(ws + starts_offset(count, 1, LG)), (int)unitsB[LG], phase, dev, nsm, stream); break;
                LFD_CZT_CASE(6) LFD_CZT_CASE(7) LFD_CZT_CASE(8) LFD_CZT_CASE(9) LFD_CZT_CASE(10) LFD_CZT_CASE(11) LFD_CZT_CASE(12)
                LFD_CZT_CASE(13)
#undef LFD_CZT_CASE
            default: break;
            }
            if (rc) return rc;
        }
    }
    return 0;
}

}  // namespace lfd
