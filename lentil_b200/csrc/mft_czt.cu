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
//  * work units (ROWS rows of a plane) are dealt round-robin over exact per-length unit tables (adjacent rows are written
//    concurrently, so the scattered stores complete their lines in L2); planes of a batch whose row stage is the same
//    computation (a grid of field points) share one stage-A result.
//  * half-length fast path: when n_in and n_out both fit L / 2 (the usual case) the first butterfly skips its zero half, the
//    last one forms only the wanted half, and interior threads load / store without predicates.
//  * from L = 2048 up the NEXT unit's input rows are copied into a shared-memory staging buffer by the bulk-copy engine
//    (cp.async.bulk + mbarrier) while the current unit computes; the pass twiddle, half of H and part of the post-chirp are
//    fetched before the barrier that precedes their pass.
//  * the same device code is compiled twice (mft_czt_body.cuh): complex128 / FP64, and complex64 / FP32 for the optional
//    complex64 mode (tables still built in float64 and rounded once).
//  * L <= 8192: one buffer of L complex128 (139 KB with padding) is the most that fits the 227 KB of an SM; longer
//    transforms use the folded DMMA execution (the dispatcher in mft_c128.cu decides).
#include "mft_czt_common.cuh"
#include <map>
#include <mutex>
#include <tuple>

namespace lfd {
namespace czt {

__device__ double2 g_tw[MAX_LOG2L - MIN_LOG2L + 1][TW_PER_LEN];
__global__ void roots_kernel() { fill_roots(g_tw); }

// ---- the row transform, once per precision ----------------------------------------------------------------------
namespace f64 {
using RL = double;
using V2 = double2;
constexpr int REG_THREADS = LFD_CZT_F64_THREADS;  // threads per SM the register allocation must allow (512 -> 128 registers)
__device__ __forceinline__ V2 mk2(RL x, RL y) { return make_double2(x, y); }
__device__ __forceinline__ const V2 *tw_table(int lg) { return g_tw[lg - MIN_LOG2L]; }
// amp * exp(2 pi i opd / lambda): phase in cycles, reduced exactly (K1's arithmetic, pupil_prep.cu, with the division by
// lambda replaced by a multiplication with its rounded reciprocal: at most one ulp of the phase in cycles, ~1e-16 rad)
__device__ __forceinline__ V2 phasor(double am, double op, double, double inv_lam) {
    const double tcyc = op * inv_lam;
    double sn, cs;
    cis_unit(tcyc - rint(tcyc), cs, sn);
    return make_double2(am * cs, am * sn);
}
#include "mft_czt_body.cuh"
}  // namespace f64

using namespace f64;      // the tables below are always computed in float64 (and rounded once for the complex64 build)

// ---- per (plane, axis): pre / post chirps and the transformed chirp filter H = FFT_L(h) -----------------
template <class OUT> __device__ __forceinline__ OUT to_out(double2 v);
template <> __device__ __forceinline__ double2 to_out<double2>(double2 v) { return v; }
template <> __device__ __forceinline__ float2 to_out<float2>(double2 v) { return make_float2((float)v.x, (float)v.y); }

template <int LOG2L, class OUT>
__global__ void __launch_bounds__((1 << LOG2L) / 16)
czt_tables_kernel(const Plane *__restrict__ descs) {
    constexpr int L = 1 << LOG2L, T = L / 16, NREG = nreg(LOG2L), RT = turn_radix(LOG2L), NB = 16 / RT, NS = L / RT;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    double2 *sm = reinterpret_cast<double2 *>(sm_raw);
    const Plane &d = descs[blockIdx.x];
    const bool axisA = blockIdx.y == 0;
    if ((axisA ? d.logLA : d.logLB) != LOG2L) return;
    const int nin = axisA ? d.n : d.m, nout = axisA ? d.N : d.M;
    const double alpha = axisA ? d.alpha_c : d.alpha_r, x0 = axisA ? d.x0c : d.x0r, y0 = axisA ? d.y0c : d.y0r;
    OUT *pre = (OUT *)(axisA ? d.preA : d.preB), *post = (OUT *)(axisA ? d.postA : d.postB), *H = (OUT *)(axisA ? d.HA : d.HB);
    const double post_scale = (axisA ? 1.0 : d.scale) / (double)L;
    const int t = threadIdx.x;
    double c, s;
    for (int i = t; i < nin; i += T) {
        const double r = (double)i + x0;
        cis_cycles(alpha, r, 0.5 * r, d.sgn, c, s);
        pre[i] = to_out<OUT>(make_double2(c, s));
    }
    for (int u = t; u < nout; u += T) {
        const double v = (double)u + y0;
        cis_cycles(alpha, v, 0.5 * v, d.sgn, c, s);
        post[u] = to_out<OUT>(make_double2(c * post_scale, s * post_scale));
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
        for (int r = 0; r < RT; ++r) H[j + r * NS] = to_out<OUT>(u[r]);
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

// c64: complex64 planes (the FP32 build of the row transform); element size 8 instead of 16 bytes
size_t czt_workspace_bytes(const lfd_mft_desc *descs, int count, bool c64) {
    const size_t es = c64 ? sizeof(float2) : sizeof(double2);
    size_t bytes = header_bytes(count);
    for (int i = 0; i < count; ++i) {
        const lfd_mft_desc &p = descs[i];
        bytes += al((size_t)p.N * pad_rows(p.m) * es);
        bytes += al(((size_t)p.n + p.N + ((size_t)1 << log2_len(p.n, p.N))) * es);
        bytes += al(((size_t)p.m + p.M + ((size_t)1 << log2_len(p.m, p.M))) * es);
    }
    return bytes;
}

template <int LOG2L>
static int launch_for_length(const Plane *dd, int count, const int *starts_a, int total_a, const int *starts_b, int total_b,
                             int phase, int dev, int nsm, bool c64, cudaStream_t stream) {
    constexpr int L = 1 << LOG2L, NT = cta_threads(LOG2L);
    const int smem_tab = (L + L / 16) * (int)sizeof(double2);                   // the tables are always built in float64
    const int smem = stage_smem_bytes(LOG2L, sizeof(double2));
    if (phase == 0) {
        if (c64) {
            if (ensure_dynamic_smem(dev, (const void *)czt_tables_kernel<LOG2L, float2>, smem_tab)) return 1;
            czt_tables_kernel<LOG2L, float2><<<dim3(count, 2), L / 16, smem_tab, stream>>>(dd);
        } else {
            if (ensure_dynamic_smem(dev, (const void *)czt_tables_kernel<LOG2L, double2>, smem_tab)) return 1;
            czt_tables_kernel<LOG2L, double2><<<dim3(count, 2), L / 16, smem_tab, stream>>>(dd);
        }
        LFD_CUDA_OK(cudaGetLastError());
        count_launch();
        return 0;
    }
    if (phase == 1 && total_a > 0)
        return c64 ? czt_f32_launch_stage(LOG2L, true, total_a, dev, nsm, dd, starts_a, count, stream)
                   : launch_stage(f64::czt_stage_kernel<LOG2L, true>, NT, smem, total_a, dev, nsm, dd, starts_a, count, stream);
    if (phase == 2 && total_b > 0)
        return c64 ? czt_f32_launch_stage(LOG2L, false, total_b, dev, nsm, dd, starts_b, count, stream)
                   : launch_stage(f64::czt_stage_kernel<LOG2L, false>, NT, smem, total_b, dev, nsm, dd, starts_b, count, stream);
    return 0;
}

int launch_mft_czt(const lfd_mft_desc *descs, int count, void *workspace, size_t workspace_bytes, cudaStream_t stream,
                   const lfd_pupil_src *src, int intensity_out, bool c64) {
    const size_t es = c64 ? sizeof(float2) : sizeof(double2);
    if (count == 0) return 0;
    LFD_REQUIRE(descs && workspace, "lfd_mft_c128 (chirp-z): NULL argument");
    LFD_REQUIRE(czt_supported(descs, count), "lfd_mft_c128 (chirp-z): a plane needs an FFT longer than %d", 1 << MAX_LOG2L);
    const size_t need = czt_workspace_bytes(descs, count, c64);
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
    if (c64 && czt_f32_ensure_roots(dev, stream)) return 1;

    const size_t hdr = header_bytes(count);
    char *hbuf = (char *)calloc(hdr, 1);
    LFD_REQUIRE(hbuf != nullptr, "out of host memory");
    Plane *h = (Plane *)hbuf;
    char *ws = (char *)workspace;
    size_t off = hdr;
    bool useA[MAX_LOG2L + 1] = {false}, useB[MAX_LOG2L + 1] = {false};
    long long unitsA[MAX_LOG2L + 1] = {0}, unitsB[MAX_LOG2L + 1] = {0};
    // Planes whose row stage is the SAME computation — same input (array or pupil window and wavelength) and the same column
    // geometry (n -> N, alpha_c, offset, shift) — share one stage-A result: field points that differ only in their row shift
    // (a grid of field points, BASELINE configs[4]) transform their rows once.  Same arithmetic, so results are unchanged.
    typedef std::tuple<const void *, long long, const void *, const void *, const void *, int, int, int, double,
                       int, int, int, double, double, double, int> StageAKey;
    std::map<StageAKey, int> leaders;
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
        d.f = p.f; d.ldf = p.ldf; d.out = p.out; d.ldo = p.ldo;
        d.m = p.m; d.n = p.n; d.M = p.M; d.N = p.N; d.mpad = pad_rows(p.m);
        d.logLA = log2_len(p.n, p.N); d.logLB = log2_len(p.m, p.M); d.intensity = intensity_out;
        d.Gt = ws + off; off += al((size_t)p.N * d.mpad * es);
        const size_t LA = (size_t)1 << d.logLA, LB = (size_t)1 << d.logLB;
        d.preA = ws + off; d.postA = ws + off + (size_t)p.n * es; d.HA = ws + off + ((size_t)p.n + p.N) * es;
        off += al(((size_t)p.n + p.N + LA) * es);
        d.preB = ws + off; d.postB = ws + off + (size_t)p.m * es; d.HB = ws + off + ((size_t)p.m + p.M) * es;
        off += al(((size_t)p.m + p.M + LB) * es);
        d.alpha_r = p.alpha_r; d.alpha_c = p.alpha_c;
        d.x0r = -floor(p.m / 2.0) + p.off_r; d.y0r = -floor(p.M / 2.0) - p.shift_r;
        d.x0c = -floor(p.n / 2.0) + p.off_c; d.y0c = -floor(p.N / 2.0) - p.shift_c;
        d.sgn = p.inverse ? 1.0 : -1.0;
        d.scale = p.unitary ? sqrt(fabs(p.alpha_r * p.alpha_c)) : 1.0;
        if (p.inverse) d.scale /= ((double)p.m * (double)p.n);
        if (src) {
            d.amp = src[i].amp; d.opd = src[i].opd; d.mask = src[i].mask;
            d.pld = src[i].n_c; d.pr0 = src[i].r0; d.pc0 = src[i].c0; d.wavelength = src[i].wavelength;
            d.inv_wavelength = 1.0 / src[i].wavelength;
        }
        useA[d.logLA] = true; useB[d.logLB] = true;
        const StageAKey key(src ? nullptr : p.f, src ? 0 : p.ldf, d.amp, d.opd, d.mask, src ? src[i].n_c : 0, d.pr0, d.pc0,
                            src ? d.wavelength : 0.0, p.m, p.n, p.N, p.alpha_c, p.off_c, p.shift_c, p.inverse ? 1 : 0);
        const auto lead = leaders.find(key);
        const bool shared = lead != leaders.end();
        if (shared) d.Gt = h[lead->second].Gt;          // stage B reads the leader's transposed intermediate
        else leaders.emplace(key, i);
        // unit start tables: stage A deals the m input rows of the plane, stage B the N columns, ROWS at a time
        for (int lg = MIN_LOG2L; lg <= MAX_LOG2L; ++lg) {
            int *sa = (int *)(hbuf + starts_offset(count, 0, lg)), *sb = (int *)(hbuf + starts_offset(count, 1, lg));
            sa[i] = (int)unitsA[lg]; sb[i] = (int)unitsB[lg];
        }
        const int ra = rows_for_length(d.logLA), rb = rows_for_length(d.logLB);
        unitsA[d.logLA] += shared ? 0 : (p.m + ra - 1) / ra;
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
                                                             (const int *)(ws + starts_offset(count, 1, LG)), (int)unitsB[LG], phase, dev, nsm, c64, stream); break;
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
