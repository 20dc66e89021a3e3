// Shared device/host helpers for the lentil_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/lentil_b200.h"

namespace lfd {

// ---- error plumbing (no exceptions across the C ABI) ------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Per-DEVICE bookkeeping (a process may drive several GPUs; cudaFuncSetAttribute applies to the current device only).
// ensure_dynamic_smem raises a kernel's dynamic shared-memory limit once per (device, kernel); sm_count caches
// cudaDevAttrMultiProcessorCount.  Both are thread-safe; ensure_dynamic_smem returns non-zero (error set) on failure.
int ensure_dynamic_smem(int device, const void *kernel, int bytes);
int sm_count(int device);
int current_sm_count();          // of the calling thread's current device; 0 on error
// grid-size cap helper: SMs of the current device (148 on B200); never 0
inline int sm_or_default() { const int n = current_sm_count(); return n > 0 ? n : 148; }

#define LFD_CUDA_OK(expr)                                                                   \
    do {                                                                                    \
        cudaError_t e__ = (expr);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            lfd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),         \
                           __FILE__, __LINE__);                                             \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

#define LFD_REQUIRE(cond, ...)                                                              \
    do {                                                                                    \
        if (!(cond)) {                                                                      \
            lfd::set_error(__VA_ARGS__);                                                    \
            return 2;                                                                       \
        }                                                                                   \
    } while (0)

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__

// exp(sgn * 2 pi i * alpha * x * y): the phase is formed and range-reduced in CYCLES with the
// rounding error of both products carried along, and only then handed to sincospi.  |phase|
// reaches 1e2..1e4 rad on this path (SURVEY.md hard part 2), so reducing in radians would
// throw away 3-4 digits.
__device__ __forceinline__ void cis_cycles(double alpha, double x, double y, double sgn,
                                           double &c, double &s) {
    double p  = x * y;
    double pe = fma(x, y, -p);
    double t  = alpha * p;
    double te = fma(alpha, p, -t) + alpha * pe;
    double r  = (t - rint(t)) + te;
    sincospi(2.0 * r, &s, &c);
    s *= sgn;
}

// cos and sin of 2 pi r for a phase r already reduced to [-0.5, 0.5] CYCLES (what the phasor kernels have after
// r = t - rint(t)).  Octant reduction k = rint(4 r), f = r - k/4 (exact), theta = 2 pi f in [-pi/4, pi/4], then the classic
// minimax kernels in theta^2 (the published fdlibm __kernel_sin / __kernel_cos coefficients, error < 1 ulp there) and a
// rotation by k quarter turns.  ~25 FP64 instructions instead of the ~60 of a general sincospi call, max error 1.6e-16
// against 40-digit arithmetic (checked over 2e6 random phases and every octant boundary).
__device__ __forceinline__ void cis_unit(double r, double &c, double &s) {
    const double kd = rint(4.0 * r);
    const double f = fma(-0.25, kd, r);
    const double th = f * 6.283185307179586476925;
    const double z = th * th;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sn = fma(th * z, ps, th);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cs = fma(z * z, pc, fma(-0.5, z, 1.0));
    const int k = (int)kd & 3;                          // quarter turns: 0, 1, 2 (= -2), 3 (= -1)
    c = (k & 1) ? sn : cs;
    s = (k & 1) ? cs : sn;
    if (k == 1 || k == 2) c = -c;
    if (k >= 2) s = -s;
}

__device__ __forceinline__ double neg_f64(double v) {
    // sign flip on the integer pipe: keeps the FP64 pipe for DMMA
    return __hiloint2double(__double2hiint(v) ^ 0x80000000, __double2loint(v));
}

// D(8x8) += A(8x4) * B(4x8), all fp64.  sm_100a lowers every mma.sync f64 shape to this one
// SASS instruction (DMMA.8x8x4), so it is the native granule.
//   A: lane holds A[g][t]        (g = lane / 4, t = lane % 4)
//   B: lane holds B[t][g]
//   C/D: lane holds D[g][2t], D[g][2t+1]
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1)
        : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool pred) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    int bytes = pred ? 16 : 0;  // src-size 0 => 16 bytes of zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src),
                 "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// ---- mbarrier + bulk-copy (TMA engine) helpers ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_addr(b)), "r"(parity) : "memory");
    }
}
// global -> shared bulk copy (multiple of 16 bytes, both sides 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

#endif  // __CUDACC__

}  // namespace lfd
