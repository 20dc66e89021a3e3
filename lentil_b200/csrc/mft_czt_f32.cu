// K2b, chirp-z execution in complex64: the FP32 build of the row transform of mft_czt.cu (same device code,
// mft_czt_body.cuh, compiled with float2).  Chirps and the transformed chirp filter H are built in float64 by the tables
// kernel of mft_czt.cu and rounded once; the phasor phase of the fused pupil path is formed and reduced in float64.
// A separate translation unit only so that the two precisions compile in parallel.
#include "mft_czt_common.cuh"
#include <mutex>

namespace lfd {
namespace czt {

__device__ float2 g_twf[MAX_LOG2L - MIN_LOG2L + 1][TW_PER_LEN];      // the pass twiddles of mft_czt.cu rounded once
__global__ void roots_f32_kernel() { fill_roots(g_twf); }

namespace f32 {
using RL = float;
using V2 = float2;
constexpr int REG_THREADS = LFD_CZT_F32_THREADS;  // threads per SM the register allocation must allow (768 -> 85 registers)
__device__ __forceinline__ V2 mk2(RL x, RL y) { return make_float2(x, y); }
__device__ __forceinline__ const V2 *tw_table(int lg) { return g_twf[lg - MIN_LOG2L]; }
// the phase is still formed and reduced in float64 (it reaches 1e2 .. 1e4 cycles); sine and cosine of the reduced phase in fp32
__device__ __forceinline__ V2 phasor(double am, double op, double, double inv_lam) {
    const double tcyc = op * inv_lam;             // one FP64 multiply instead of a division: 1 ulp of the phase in cycles
    float sn, cs;
    sincospif((float)(2.0 * (tcyc - rint(tcyc))), &sn, &cs);
    const float a = (float)am;
    return make_float2(a * cs, a * sn);
}
#include "mft_czt_body.cuh"
}  // namespace f32

int czt_f32_ensure_roots(int dev, cudaStream_t stream) {
    static std::mutex mu;
    static bool ready[64] = {false};
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 64 || !ready[dev]) {
        roots_f32_kernel<<<dim3((TW_PER_LEN + 255) / 256, MAX_LOG2L - MIN_LOG2L + 1), 256, 0, stream>>>();
        LFD_CUDA_OK(cudaGetLastError());
        LFD_CUDA_OK(cudaStreamSynchronize(stream));
        count_launch();
        if (dev < 64) ready[dev] = true;
    }
    return 0;
}

template <int LOG2L>
static int launch_f32(bool stage_a, int total, int dev, int nsm, const Plane *dd, const int *starts, int count, cudaStream_t stream) {
    const int smem = stage_smem_bytes(LOG2L, sizeof(float2));
    return stage_a ? launch_stage(f32::czt_stage_kernel<LOG2L, true>, cta_threads(LOG2L), smem, total, dev, nsm, dd, starts, count, stream)
                   : launch_stage(f32::czt_stage_kernel<LOG2L, false>, cta_threads(LOG2L), smem, total, dev, nsm, dd, starts, count, stream);
}

int czt_f32_launch_stage(int lg, bool stage_a, int total, int dev, int nsm, const Plane *dd, const int *starts, int count, cudaStream_t stream) {
    switch (lg) {
#define LFD_CZT_CASE(LG) case LG: return launch_f32<LG>(stage_a, total, dev, nsm, dd, starts, count, stream);
        LFD_CZT_CASE(6) LFD_CZT_CASE(7) LFD_CZT_CASE(8) LFD_CZT_CASE(9) LFD_CZT_CASE(10) LFD_CZT_CASE(11) LFD_CZT_CASE(12) LFD_CZT_CASE(13)
#undef LFD_CZT_CASE
    default: break;
    }
    LFD_REQUIRE(false, "lfd_mft (chirp-z, complex64): unsupported transform length 2^%d", lg);
}

}  // namespace czt
}  // namespace lfd
