// K2a, chirp-z execution (LFD_MFT_CZT) — the same matrix Fourier transform, lentil/fourier.py:5-121,
//     F = E1 f E2 sqrt|a_r a_c|,  E1[u,i] = exp(sgn 2 pi i a_r (R_i + o_r)(U_u - s_r)),
// evaluated as two passes of length-L convolutions instead of two GEMMs.  With R' = R + o, U' = U - s and
// D = U' - R' (D depends on u - i only),  R'U' = (R'^2 + U'^2 - D^2) / 2, hence along one axis
//     y[u] = post[u] * sum_i (pre[i] x[i]) h[u - i],   pre = cis(sgn pi a R'^2), post = cis(sgn pi a U'^2), h = cis(-sgn pi a D^2)
// (Bluestein / chirp-z).  The convolution runs per row as FFT_L -> x H -> IFFT_L in shared memory, L the power of two
// >= n_in + n_out - 1 (2048 for the 1001 -> 1024 bench shape), H = FFT_L(h) built once per (plane, axis).
// One plane of 1001^2 -> 1024^2 is 2025 row transforms ~ 0.5 GFLOP of FP64 instead of the 4.15 GFLOP the folded DMMA
// kernel executes (16.6 GFLOP algorithmic); exact for any alpha / shift / offset / parity like the other executions.
//
//  * every chirp phase is formed in cycles with error-free products and reduced exactly before sincospi (cis_cycles),
//    the FFT twiddles come from per-pass tables of exactly reduced sincospi values: parity vs the oracle ~1e-13.
//  * FFT: Stockham auto-sort in shared memory, radix 8 passes (+ one radix 4 or 2 pass), L/8 threads, one butterfly per
//    thread and pass, two padded buffers (one 16-byte element of padding per 8: the stride-8 stores of the first pass
//    are bank-conflict free).
//  * the inverse transform is run as the adjoints of the forward passes in reverse order, so the last forward pass, the
//    product with H and the first inverse pass of a thread happen in registers (pass_turn); the first forward pass loads
//    straight from global memory and the last inverse pass stores straight to it (czt_row).
//  * stage A transforms the rows of f (contiguous loads; K1 can be fused: the phasor amp*mask*exp(2 pi i opd / lambda)
//    is formed in the load, lentil/plane.py:502-507) and stores its result transposed, stage B transforms the rows of
//    that (= the columns of the plane) and stores F, or |F|^2 as float64 when the caller only wants intensities.
//  * L <= 4096 (two buffers of L complex128 must fit in 227 KB of shared memory): larger planes use the folded DMMA
//    execution (the dispatcher in mft_c128.cu decides).
#include "lfd_common.cuh"
#include <mutex>

namespace lfd {
namespace czt {

constexpr int MAX_LOG2L = 12, MIN_LOG2L = 6;
// Pass twiddles, one contiguous run per (FFT length, pass): g_tw[lg][(Ns - 1) / 7 + k] = exp(-2 pi i k / (Ns R)) for the pass
// whose sub-transform length is Ns = 8^p (R = 8, or the 4 / 2 of the last pass), k = 0 .. Ns - 1.  Consecutive butterflies
// read consecutive entries (a warp: 512 contiguous bytes) instead of gathering from one table of L-th roots.
constexpr int TW_PER_LEN = 1024;                 // 1 + 8 + 64 + 512 = 585 entries for the longest transform
__device__ double2 g_tw[MAX_LOG2L + 1][TW_PER_LEN];

__global__ void roots_kernel() {
    const int lg = MIN_LOG2L + blockIdx.y;
    const int npass = lg / 3 + (lg % 3 ? 1 : 0), rlast = lg % 3 == 0 ? 8 : (lg % 3 == 2 ? 4 : 2);
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    int Ns = 1, off = 0;
    for (int p = 0; p < npass; ++p) {
        if (e >= off && e < off + Ns) {
            const int k = e - off, R = (p == npass - 1) ? rlast : 8;
            double s, c;
            sincospi(-2.0 * (double)k / (double)(Ns * R), &s, &c);      // k / (Ns R) is exact (power-of-two denominator)
            g_tw[lg][e] = make_double2(c, s);
        }
        off += Ns; Ns *= 8;
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

__device__ __forceinline__ int P(int i) { return i + (i >> 3); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i (S = +1: forward transform) or by +i (S = -1: inverse)
template <int S> __device__ __forceinline__ double2 mul_mi(double2 a) { return S > 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x); }

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

// Element sources / sinks of a pass: shared memory (padded layout), or the caller's functor (global memory).
struct SmemIn  { const double2 *p; __device__ __forceinline__ double2 operator()(int i) const { return p[P(i)]; } };
struct SmemOut { double2 *p; __device__ __forceinline__ void operator()(int i, double2 v) const { p[P(i)] = v; } };
// twiddle of butterfly j in the pass with sub-transform length Ns (callers fetch it before the barrier that precedes the pass)
template <int LOG2L> __device__ __forceinline__ double2 pass_twiddle(int j, int Ns) { return g_tw[LOG2L][(Ns - 1) / 7 + (j & (Ns - 1))]; }

template <int R, int S, int LOG2L, class In, class Out>
__device__ __forceinline__ void pass(const In &in, const Out &out, int j, int Ns, double2 w1 = make_double2(1.0, 0.0)) {
    constexpr int L = 1 << LOG2L;
    const int k = j & (Ns - 1);
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in(j + r * (L / R));
    if (Ns > 1) {
        if (S < 0) w1.y = -w1.y;
        double2 w = w1;
#pragma unroll
        for (int r = 1; r < R; ++r) { v[r] = cmul(v[r], w); if (r + 1 < R) w = cmul(w, w1); }
    }
    if constexpr (R == 8) dft8<S>(v);
    else if constexpr (R == 4) dft4<S>(v[0], v[1], v[2], v[3]);
    else dft2p<S>(v[0], v[1]);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out(j0 + r * Ns, v[r]);
}

// all butterflies of one pass that thread t owns (L / 8 threads: one radix-8, two radix-4 or four radix-2 butterflies)
template <int R, int S, int LOG2L, class In, class Out>
__device__ __forceinline__ void pass_all(const In &in, const Out &out, int t, int Ns) {
    constexpr int T = (1 << LOG2L) / 8;
#pragma unroll
    for (int q = 0; q < 8 / R; ++q) pass<R, S, LOG2L>(in, out, t + q * T, Ns, pass_twiddle<LOG2L>(t + q * T, Ns));
}

// Plain FFT of L elements with L / 8 threads (used for the chirp filter H; the row transforms use czt_row below).  The first
// pass reads through `first`, the last pass writes through `last` (functors: shared or global memory); the passes in
// between ping-pong between the padded buffers a and b, starting by WRITING a.  Every pass is followed by a __syncthreads().
template <int S, int LOG2L, class In, class Out>
__device__ __forceinline__ void fft(double2 *a, double2 *b, int t, const In &first, const Out &last) {
    constexpr int N8 = LOG2L / 3, REM = LOG2L % 3, NPASS = N8 + (REM ? 1 : 0);
    constexpr int RLAST = REM == 0 ? 8 : (REM == 2 ? 4 : 2);
    static_assert(NPASS >= 2, "at least two passes");
    int Ns = 1;
    // first pass: radix 8 from `first` into a
    pass_all<8, S, LOG2L>(first, SmemOut{a}, t, Ns);
    __syncthreads();
    Ns *= 8;
    double2 *src = a, *dst = b;
#pragma unroll
    for (int p = 1; p < NPASS - 1; ++p) {
        pass_all<8, S, LOG2L>(SmemIn{src}, SmemOut{dst}, t, Ns);
        __syncthreads();
        double2 *x = src; src = dst; dst = x;
        Ns *= 8;
    }
    pass_all<RLAST, S, LOG2L>(SmemIn{src}, last, t, Ns);
    __syncthreads();
}

// The adjoint of a forward pass: reads where the forward pass writes, takes the conjugate butterfly, multiplies by the
// conjugate twiddles and writes where the forward pass reads.  The forward passes in reverse order, each replaced by its
// adjoint, are the conjugate (= unnormalised inverse) transform — and the first of them reads exactly the elements the
// last forward pass of the same thread produced, so that hand-over needs no trip through shared memory.
template <int R, int LOG2L, class In, class Out>
__device__ __forceinline__ void pass_adj(const In &in, const Out &out, int j, int Ns, double2 w1) {
    constexpr int L = 1 << LOG2L;
    const int k = j & (Ns - 1);
    const int j0 = (j - k) * R + k;
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = in(j0 + r * Ns);
    if constexpr (R == 8) dft8<-1>(v);
    else if constexpr (R == 4) dft4<-1>(v[0], v[1], v[2], v[3]);
    else dft2p<-1>(v[0], v[1]);
    if (Ns > 1) {
        w1.y = -w1.y;
        double2 w = w1;
#pragma unroll
        for (int r = 1; r < R; ++r) { v[r] = cmul(v[r], w); if (r + 1 < R) w = cmul(w, w1); }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out(j + r * (L / R), v[r]);
}

// last forward pass, product with H and first adjoint pass of one butterfly, in registers and in place in `buf`
template <int R, int LOG2L>
__device__ __forceinline__ void pass_turn(double2 *buf, const double2 *__restrict__ H, int j, int Ns, double2 w1) {
    constexpr int L = 1 << LOG2L;
    const int k = j & (Ns - 1);
    const int j0 = (j - k) * R + k;
    double2 v[R], w[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = buf[P(j + r * (L / R))];
    w[1] = w1;
#pragma unroll
    for (int r = 2; r < R; ++r) w[r] = cmul(w[r - 1], w1);
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = cmul(v[r], w[r]);
    if constexpr (R == 8) dft8<1>(v);
    else if constexpr (R == 4) dft4<1>(v[0], v[1], v[2], v[3]);
    else dft2p<1>(v[0], v[1]);
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = cmul(v[r], H[j0 + r * Ns]);
    if constexpr (R == 8) dft8<-1>(v);
    else if constexpr (R == 4) dft4<-1>(v[0], v[1], v[2], v[3]);
    else dft2p<-1>(v[0], v[1]);
#pragma unroll
    for (int r = 1; r < R; ++r) v[r] = cmul(v[r], make_double2(w[r].x, -w[r].y));
#pragma unroll
    for (int r = 0; r < R; ++r) buf[P(j + r * (L / R))] = v[r];
}

// One row of the chirp-z convolution: y = IFFT(FFT(x) * H).  `load8(j, v)` fills v[r] with x[j + r L/8] (global memory,
// all eight loads issued before any arithmetic), y leaves through `store`.  The twiddle of the next pass is fetched
// before the barrier that precedes it, so its latency overlaps the barrier wait.
template <int LOG2L, class Load8, class Out>
__device__ __forceinline__ void czt_row(double2 *a, double2 *b, int t, const Load8 &load8, const double2 *__restrict__ H, const Out &store) {
    constexpr int T = (1 << LOG2L) / 8, N8 = LOG2L / 3, REM = LOG2L % 3, NPASS = N8 + (REM ? 1 : 0);
    constexpr int RLAST = REM == 0 ? 8 : (REM == 2 ? 4 : 2);
    static_assert(NPASS >= 2, "at least two passes");
    int Ns = 8;
    {   // first pass (Ns = 1: no twiddles): global -> registers -> a
        double2 v[8];
        load8(t, v);
        dft8<1>(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) a[P(8 * t + r)] = v[r];
    }
    double2 *src = a, *dst = b;
#pragma unroll
    for (int p = 1; p < NPASS - 1; ++p) {
        const double2 w1 = pass_twiddle<LOG2L>(t, Ns);
        __syncthreads();
        pass<8, 1, LOG2L>(SmemIn{src}, SmemOut{dst}, t, Ns, w1);
        double2 *x = src; src = dst; dst = x;
        Ns *= 8;
    }
    {
        double2 wt[8 / RLAST];
#pragma unroll
        for (int q = 0; q < 8 / RLAST; ++q) wt[q] = pass_twiddle<LOG2L>(t + q * T, Ns);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8 / RLAST; ++q) pass_turn<RLAST, LOG2L>(src, H, t + q * T, Ns, wt[q]);
    }
#pragma unroll
    for (int p = NPASS - 2; p >= 1; --p) {
        Ns /= 8;
        const double2 w1 = pass_twiddle<LOG2L>(t, Ns);
        __syncthreads();
        pass_adj<8, LOG2L>(SmemIn{src}, SmemOut{dst}, t, Ns, w1);
        double2 *x = src; src = dst; dst = x;
    }
    __syncthreads();
    pass_adj<8, LOG2L>(SmemIn{src}, store, t, 1, make_double2(1.0, 0.0));
    __syncthreads();
}

// ---- per (plane, axis): pre / post chirps and the transformed chirp filter H = FFT_L(h) -----------------
template <int LOG2L>
__global__ void __launch_bounds__((1 << LOG2L) / 8)
czt_tables_kernel(const Plane *__restrict__ descs) {
    constexpr int L = 1 << LOG2L, T = L / 8;
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
    double2 *a = sm, *b = sm + (L + L / 8);
    const double dd = y0 - x0;                       // D = (u - i) + (y0 - x0)
    const double sg = d.sgn;
    // circular position q holds the lag p = u - i: p = q for q < nout, p = q - L for the negative lags
    auto chirp = [=](int q) -> double2 {
        const int p = q < nout ? q : q - L;
        if (p <= -nin || p >= nout) return make_double2(0.0, 0.0);
        const double D = (double)p + dd;
        double cc, ss;
        cis_cycles(alpha, D, 0.5 * D, -sg, cc, ss);
        return make_double2(cc, ss);
    };
    auto to_H = [=](int q, double2 v) { H[q] = v; };
    fft<1, LOG2L>(a, b, t, chirp, to_H);
}

// ---- one stage: every row of every plane whose FFT length is L ----------------------------------------------
// STAGE_A: row i of f (n elements, or the fused phasor) -> Gt[:, i] (N outputs, transposed store)
// else   : row v of Gt (m elements)                     -> out[:, v] (M outputs; complex128 or |.|^2 float64)
// resident CTAs the register allocation must allow: ~768 threads per SM (3 CTAs of the 2048-point transform)
constexpr int min_ctas(int log2l) { return (1 << log2l) / 8 >= 768 ? 1 : (768 / ((1 << log2l) / 8) > 16 ? 16 : 768 / ((1 << log2l) / 8)); }

template <int LOG2L, bool STAGE_A>
__global__ void __launch_bounds__((1 << LOG2L) / 8, min_ctas(LOG2L))
czt_stage_kernel(const Plane *__restrict__ descs, int count, int max_rows) {
    constexpr int L = 1 << LOG2L, T = L / 8;
    extern __shared__ double2 sm[];
    double2 *a = sm, *b = sm + (L + L / 8);
    const int t = threadIdx.x;
    const long long total = (long long)count * max_rows;
    __shared__ Plane sd;                 // the plane this CTA is working on (rows are dealt plane-major: it changes rarely)
    int cur = -1;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const int plane = (int)(w / max_rows);
        const int row = (int)(w - (long long)plane * max_rows);
        if (plane != cur) {              // uniform over the CTA
            __syncthreads();
            const unsigned long long *g = reinterpret_cast<const unsigned long long *>(descs + plane);
            unsigned long long *sdw = reinterpret_cast<unsigned long long *>(&sd);
            for (int i = t; i < (int)(sizeof(Plane) / 8); i += T) sdw[i] = g[i];
            __syncthreads();
            cur = plane;
        }
        const Plane &d = sd;
        const int nrows = STAGE_A ? d.m : d.N;
        if (row >= nrows || (STAGE_A ? d.logLA : d.logLB) != LOG2L) continue;     // uniform over the CTA
        const int nin = STAGE_A ? d.n : d.m, nout = STAGE_A ? d.N : d.M;
        const double2 *__restrict__ pre = STAGE_A ? d.preA : d.preB;
        const double2 *__restrict__ post = STAGE_A ? d.postA : d.postB;
        const double2 *__restrict__ H = STAGE_A ? d.HA : d.HB;
        const Plane *dp = &d;
        // first forward pass: the eight inputs of this thread straight from global memory (x pre-chirp; zero beyond the
        // input length), every load issued before the arithmetic
        auto load = [=](int j, double2 (&v)[8]) {
            double2 pr[8];
            if (STAGE_A && dp->amp != nullptr) {
                double am[8], op[8];
                const long long base = (long long)(dp->pr0 + row) * dp->pld + dp->pc0;
                const unsigned char *mk = dp->mask;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = j + r * (L / 8);
                    const bool in = i < nin;
                    am[r] = in ? dp->amp[base + i] : 0.0;
                    op[r] = in ? dp->opd[base + i] : 0.0;
                    pr[r] = in ? pre[i] : make_double2(0.0, 0.0);
                    if (in && mk != nullptr && mk[base + i] == 0) am[r] = 0.0;
                }
                const double lam = dp->wavelength;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    double2 x = make_double2(0.0, 0.0);
                    if (am[r] != 0.0) {                        // same arithmetic as K1 (pupil_prep.cu): phase in cycles, reduced exactly
                        const double tcyc = op[r] / lam;
                        double sn, cs;
                        sincospi(2.0 * (tcyc - rint(tcyc)), &sn, &cs);
                        x = make_double2(am[r] * cs, am[r] * sn);
                    }
                    v[r] = cmul(x, pr[r]);
                }
            } else {
                const double2 *src = STAGE_A ? dp->f + (long long)row * dp->ldf : dp->Gt + (long long)row * dp->mpad;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int i = j + r * (L / 8);
                    const bool in = i < nin;
                    v[r] = in ? src[i] : make_double2(0.0, 0.0);
                    pr[r] = in ? pre[i] : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) v[r] = cmul(v[r], pr[r]);
            }
        };
        if (STAGE_A) {
            double2 *Gt = dp->Gt; const long long mpad = dp->mpad;
            auto store = [=](int i, double2 v) { if (i < nout) Gt[(long long)i * mpad + row] = cmul(v, post[i]); };
            czt_row<LOG2L>(a, b, t, load, H, store);
        } else if (dp->intensity) {
            double *out = (double *)dp->out; const long long ldo = dp->ldo;
            auto store = [=](int i, double2 v) {
                if (i < nout) { const double2 z = cmul(v, post[i]); out[(long long)i * ldo + row] = z.x * z.x + z.y * z.y; }
            };
            czt_row<LOG2L>(a, b, t, load, H, store);
        } else {
            double2 *out = (double2 *)dp->out; const long long ldo = dp->ldo;
            auto store = [=](int i, double2 v) { if (i < nout) out[(long long)i * ldo + row] = cmul(v, post[i]); };
            czt_row<LOG2L>(a, b, t, load, H, store);
        }
    }
}

static inline size_t al(size_t v) { return (v + 255) / 256 * 256; }
static inline int log2_len(int nin, int nout) {
    int lg = MIN_LOG2L;
    while ((1 << lg) < nin + nout - 1) ++lg;
    return lg;
}

}  // namespace czt

using namespace czt;

// true when every plane of the batch fits the shared-memory FFT (both axes)
bool czt_supported(const lfd_mft_desc *descs, int count) {
    for (int i = 0; i < count; ++i)
        if (log2_len(descs[i].n, descs[i].N) > MAX_LOG2L || log2_len(descs[i].m, descs[i].M) > MAX_LOG2L) return false;
    return true;
}

// LFD_MFT_AUTO: chirp-z wherever it can run.  Measured against the folded DMMA form on dense random planes, batched
// (scripts/small_plane_timing.py, r01o): 121^2 -> 128^2 0.66 vs 0.73 us, 241^2 -> 256^2 2.9 vs 3.1 us, 501^2 -> 512^2 12.4 vs
// 19.3 us, 1001^2 -> 1024^2 53.6 vs 139 us, 2001^2 -> 2048^2 251 vs 1036 us per plane; cfg3's 900 windows 9.5 vs 12.6 ms.
bool czt_preferred(const lfd_mft_desc *descs, int count) { return czt_supported(descs, count); }

size_t czt_workspace_bytes(const lfd_mft_desc *descs, int count) {
    size_t bytes = al((size_t)count * sizeof(Plane));
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        const int mpad = (p.m + 1) & ~1;
        bytes += al((size_t)p.N * mpad * sizeof(double2));
        bytes += al(((size_t)p.n + p.N + ((size_t)1 << log2_len(p.n, p.N))) * sizeof(double2));
        bytes += al(((size_t)p.m + p.M + ((size_t)1 << log2_len(p.m, p.M))) * sizeof(double2));
    }
    return bytes;
}

template <int LOG2L>
static int launch_for_length(const Plane *dd, int count, int max_rows_a, int max_rows_b, bool any_a, bool any_b,
                             int phase, int nsm, cudaStream_t stream) {
    constexpr int L = 1 << LOG2L, T = L / 8;
    const int smem = 2 * (L + L / 8) * (int)sizeof(double2);
    static bool attr_set = false;
    if (!attr_set) {
        LFD_CUDA_OK(cudaFuncSetAttribute(czt_tables_kernel<LOG2L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        LFD_CUDA_OK(cudaFuncSetAttribute(czt_stage_kernel<LOG2L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        LFD_CUDA_OK(cudaFuncSetAttribute(czt_stage_kernel<LOG2L, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set = true;
    }
    if (phase == 0) {
        czt_tables_kernel<LOG2L><<<dim3(count, 2), T, smem, stream>>>(dd);
        LFD_CUDA_OK(cudaGetLastError());
        count_launch();
        return 0;
    }
    int occ = 1;
    if (phase == 1 && any_a) {
        LFD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, czt_stage_kernel<LOG2L, true>, T, smem));
        const long long total = (long long)count * max_rows_a;
        const int grid = (int)(total < (long long)nsm * occ ? total : (long long)nsm * occ);
        czt_stage_kernel<LOG2L, true><<<grid, T, smem, stream>>>(dd, count, max_rows_a);
        LFD_CUDA_OK(cudaGetLastError());
        count_launch();
    }
    if (phase == 2 && any_b) {
        LFD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, czt_stage_kernel<LOG2L, false>, T, smem));
        const long long total = (long long)count * max_rows_b;
        const int grid = (int)(total < (long long)nsm * occ ? total : (long long)nsm * occ);
        czt_stage_kernel<LOG2L, false><<<grid, T, smem, stream>>>(dd, count, max_rows_b);
        LFD_CUDA_OK(cudaGetLastError());
        count_launch();
    }
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
    LFD_CUDA_OK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    {   // FFT roots: once per device and process
        static std::mutex mu;
        static bool ready[64] = {false};
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 64 && !ready[dev]) {
            roots_kernel<<<dim3(TW_PER_LEN / 256, MAX_LOG2L - MIN_LOG2L + 1), 256, 0, stream>>>();
            LFD_CUDA_OK(cudaGetLastError());
            LFD_CUDA_OK(cudaStreamSynchronize(stream));
            count_launch();
            ready[dev] = true;
        }
    }

    Plane *h = (Plane *)calloc(count, sizeof(Plane));
    LFD_REQUIRE(h != nullptr, "out of host memory");
    char *ws = (char *)workspace;
    size_t off = al((size_t)count * sizeof(Plane));
    bool useA[MAX_LOG2L + 1] = {false}, useB[MAX_LOG2L + 1] = {false};
    int max_m = 0, max_N = 0;
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        const bool ok = p.m > 0 && p.n > 0 && p.M > 0 && p.N > 0 && p.ldo >= p.N && p.out &&
                        (src ? (src[i].amp && src[i].opd && src[i].wavelength != 0.0 && src[i].r0 >= 0 && src[i].c0 >= 0 &&
                                src[i].r0 + p.m <= src[i].n_r && src[i].c0 + p.n <= src[i].n_c)
                             : (p.f && p.ldf >= p.n));
        if (!ok) {
            free(h);
            LFD_REQUIRE(false, "lfd_mft_c128 (chirp-z): plane %d has invalid shape/ld/pointers", i);
        }
        Plane &d = h[i];
        d.f = (const double2 *)p.f; d.ldf = p.ldf; d.out = p.out; d.ldo = p.ldo;
        d.m = p.m; d.n = p.n; d.M = p.M; d.N = p.N; d.mpad = (p.m + 1) & ~1;
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
        if (p.m > max_m) max_m = p.m;
        if (p.N > max_N) max_N = p.N;
    }
    cudaError_t e = cudaMemcpyAsync(workspace, h, (size_t)count * sizeof(Plane), cudaMemcpyHostToDevice, stream);
    free(h);
    LFD_CUDA_OK(e);
    const Plane *dd = (const Plane *)workspace;
    for (int phase = 0; phase < 3; ++phase) {
        for (int lg = MIN_LOG2L; lg <= MAX_LOG2L; ++lg) {
            if (!(useA[lg] || useB[lg])) continue;
            int rc = 0;
            switch (lg) {
#define LFD_CZT_CASE(LG) case LG: rc = launch_for_length<LG>(dd, count, max_m, max_N, useA[LG], useB[LG], phase, nsm, stream); break;
                LFD_CZT_CASE(6) LFD_CZT_CASE(7) LFD_CZT_CASE(8) LFD_CZT_CASE(9) LFD_CZT_CASE(10) LFD_CZT_CASE(11) LFD_CZT_CASE(12)
#undef LFD_CZT_CASE
            default: break;
            }
            if (rc) return rc;
        }
    }
    return 0;
}

}  // namespace lfd
