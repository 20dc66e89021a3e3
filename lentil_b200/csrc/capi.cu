// C-ABI plumbing: error state, launch counter, device info, the host-buffer context layer and
// the FP64 issue-rate probe.  See include/lentil_b200.h for the contract of each entry point.
#include "lfd_common.cuh"

#include <string.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <utility>

namespace lfd {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::mutex g_dev_mu;
static std::map<std::pair<int, const void *>, int> g_smem_set;      // (device, kernel) -> bytes granted
static std::map<int, int> g_sm_count;

int ensure_dynamic_smem(int device, const void *kernel, int bytes) {
    std::lock_guard<std::mutex> lock(g_dev_mu);
    auto key = std::make_pair(device, kernel);
    auto it = g_smem_set.find(key);
    if (it != g_smem_set.end() && it->second >= bytes) return 0;
    LFD_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    g_smem_set[key] = bytes;
    return 0;
}

int sm_count(int device) {
    std::lock_guard<std::mutex> lock(g_dev_mu);
    auto it = g_sm_count.find(device);
    if (it != g_sm_count.end()) return it->second;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) return 0;
    g_sm_count[device] = n;
    return n;
}

int current_sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    return sm_count(dev);
}

}  // namespace lfd

using namespace lfd;

extern "C" int lfd_abi_version(void) { return LFD_ABI_VERSION; }
extern "C" const char *lfd_last_error(void) { return g_err; }
extern "C" uint64_t lfd_launch_count(void) { return g_launches.load(); }

extern "C" size_t lfd_struct_size(int which) {
    switch (which) {
        case 0: return sizeof(lfd_mft_desc);
        case 1: return sizeof(lfd_segment);
        case 2: return sizeof(lfd_window);
        default: return 0;
    }
}

extern "C" int lfd_device_info(int device, int *out3) {
    LFD_REQUIRE(out3 != nullptr, "lfd_device_info: NULL output");
    cudaDeviceProp p;
    LFD_CUDA_OK(cudaGetDeviceProperties(&p, device));
    out3[0] = p.multiProcessorCount;
    out3[1] = p.major;
    out3[2] = p.minor;
    return 0;
}

// ---- host-buffer context ---------------------------------------------------------------------
struct lfd_ctx {
    int device;
    cudaStream_t stream;
    void *dev_in = nullptr;   size_t dev_in_bytes = 0;
    void *dev_out = nullptr;  size_t dev_out_bytes = 0;
    void *dev_ws = nullptr;   size_t dev_ws_bytes = 0;
    void *pin = nullptr;      size_t pin_bytes = 0;
};

static int grow_dev(void **p, size_t *have, size_t need) {
    if (*have >= need) return 0;
    if (*p) LFD_CUDA_OK(cudaFree(*p));
    *p = nullptr; *have = 0;
    LFD_CUDA_OK(cudaMalloc(p, need));
    *have = need;
    return 0;
}
static int grow_pin(void **p, size_t *have, size_t need) {
    if (*have >= need) return 0;
    if (*p) LFD_CUDA_OK(cudaFreeHost(*p));
    *p = nullptr; *have = 0;
    LFD_CUDA_OK(cudaMallocHost(p, need));
    *have = need;
    return 0;
}

extern "C" lfd_ctx *lfd_ctx_create(int device) {
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("lfd_ctx_create: cudaSetDevice(%d) failed", device);
        return nullptr;
    }
    lfd_ctx *ctx = new lfd_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("lfd_ctx_create: cannot create a stream on device %d", device);
        delete ctx;
        return nullptr;
    }
    return ctx;
}

extern "C" void lfd_ctx_destroy(lfd_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->dev_in) cudaFree(ctx->dev_in);
    if (ctx->dev_out) cudaFree(ctx->dev_out);
    if (ctx->dev_ws) cudaFree(ctx->dev_ws);
    if (ctx->pin) cudaFreeHost(ctx->pin);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int lfd_ctx_dft2_host(lfd_ctx *ctx, const void *f_host, int64_t ldf, int32_t m, int32_t n,
                                 double alpha_r, double alpha_c, int32_t M, int32_t N,
                                 double shift_r, double shift_c, double off_r, double off_c,
                                 int32_t unitary, int32_t inverse, void *out_host, int64_t ldo) {
    LFD_REQUIRE(ctx && f_host && out_host, "lfd_ctx_dft2_host: NULL argument");
    LFD_REQUIRE(m > 0 && n > 0 && M > 0 && N > 0 && ldf >= n && ldo >= N,
                "lfd_ctx_dft2_host: bad shapes m=%d n=%d M=%d N=%d", m, n, M, N);
    LFD_CUDA_OK(cudaSetDevice(ctx->device));
    const size_t in_bytes = (size_t)m * n * 16, out_bytes = (size_t)M * N * 16;
    if (grow_dev(&ctx->dev_in, &ctx->dev_in_bytes, in_bytes)) return 1;
    if (grow_dev(&ctx->dev_out, &ctx->dev_out_bytes, out_bytes)) return 1;
    if (grow_pin(&ctx->pin, &ctx->pin_bytes, in_bytes > out_bytes ? in_bytes : out_bytes)) return 1;

    lfd_mft_desc d;
    memset(&d, 0, sizeof(d));
    d.f = ctx->dev_in; d.ldf = n; d.out = ctx->dev_out; d.ldo = N;
    d.m = m; d.n = n; d.M = M; d.N = N;
    d.alpha_r = alpha_r; d.alpha_c = alpha_c;
    d.shift_r = shift_r; d.shift_c = shift_c; d.off_r = off_r; d.off_c = off_c;
    d.unitary = unitary; d.inverse = inverse;
    size_t ws = lfd_mft_workspace_bytes(&d, 1);
    if (grow_dev(&ctx->dev_ws, &ctx->dev_ws_bytes, ws)) return 1;

    // host rows -> pinned (packs ldf away) -> device
    for (int i = 0; i < m; ++i)
        memcpy((char *)ctx->pin + (size_t)i * n * 16, (const char *)f_host + (size_t)i * ldf * 16,
               (size_t)n * 16);
    LFD_CUDA_OK(cudaMemcpyAsync(ctx->dev_in, ctx->pin, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (lfd_mft_c128(&d, ctx->dev_ws, ctx->dev_ws_bytes, ctx->stream)) return 1;
    LFD_CUDA_OK(cudaMemcpyAsync(ctx->pin, ctx->dev_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    LFD_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < M; ++i)
        memcpy((char *)out_host + (size_t)i * ldo * 16, (const char *)ctx->pin + (size_t)i * N * 16,
               (size_t)N * 16);
    return 0;
}

// ---- FP64 issue-rate probe ---------------------------------------------------------------------
namespace lfd {

__global__ void __launch_bounds__(256) probe_dmma_kernel(double *out, int iters) {
    double acc[16][2];
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;  // keep the loop alive
}

__global__ void __launch_bounds__(256) probe_dfma_kernel(double *out, int iters) {
    double acc[16];
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (double)i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

}  // namespace lfd

extern "C" int lfd_probe_fp64(double *out3, int iters) {
    LFD_REQUIRE(out3 && iters > 0, "lfd_probe_fp64: bad arguments");
    int dev = 0;
    LFD_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    LFD_CUDA_OK(cudaGetDeviceProperties(&p, dev));
    double *d = nullptr;
    LFD_CUDA_OK(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    LFD_CUDA_OK(cudaEventCreate(&e0));
    LFD_CUDA_OK(cudaEventCreate(&e1));
    const int blocks = p.multiProcessorCount * 2, threads = 256;
    float ms = 0.f;

    probe_dmma_kernel<<<blocks, threads>>>(d, iters / 8 + 1);  // warm-up
    LFD_CUDA_OK(cudaEventRecord(e0));
    probe_dmma_kernel<<<blocks, threads>>>(d, iters);
    LFD_CUDA_OK(cudaEventRecord(e1));
    LFD_CUDA_OK(cudaEventSynchronize(e1));
    LFD_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    double warps = (double)blocks * threads / 32.0;
    out3[0] = warps * iters * 16.0 * (8.0 * 8.0 * 4.0 * 2.0) / (ms * 1e-3) / 1e12;

    probe_dfma_kernel<<<blocks, threads>>>(d, iters / 8 + 1);
    LFD_CUDA_OK(cudaEventRecord(e0));
    probe_dfma_kernel<<<blocks, threads>>>(d, iters);
    LFD_CUDA_OK(cudaEventRecord(e1));
    LFD_CUDA_OK(cudaEventSynchronize(e1));
    LFD_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    out3[1] = (double)blocks * threads * iters * 16.0 * 2.0 / (ms * 1e-3) / 1e12;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    out3[2] = khz / 1000.0;
    count_launch(4);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return 0;
}
