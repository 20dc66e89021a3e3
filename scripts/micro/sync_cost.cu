// Cost, for ONE warp, of the synchronisation primitives an MMA-issuing warp executes between tcgen05.mma groups:
// mbarrier.try_wait / test_wait on an already-completed phase, ld.acquire.shared of a flag, elect.sync,
// tcgen05.fence::after_thread_sync, __syncwarp.  Development aid:  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(long long *out, int reps) {
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ volatile int flag[4];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i]))); flag[i] = 1; }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) for (int i = 0; i < 4; ++i) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[i])) : "memory");
    __syncthreads();
    if (threadIdx.x >= 32) return;
    long long t0, t1;
    uint32_t acc = 0;
    // 1. try_wait on a completed phase (parity 0 completed)
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[i & 3])), "r"(0u) : "memory");
        acc += done;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0);
    // 2. test_wait
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar[i & 3])), "r"(0u) : "memory");
        acc += done;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = (t1 - t0);
    // 3. ld.acquire.shared flag spin
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        int v = 0;
        while (!v) asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void *)&flag[i & 3])) : "memory");
        acc += v;
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = (t1 - t0);
    // 4. elect.sync + syncwarp
    t0 = clock64();
    for (int i = 0; i < reps; ++i) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) acc += i;
        __syncwarp();
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = (t1 - t0);
    // 5. tcgen05.fence::after_thread_sync
    t0 = clock64();
    for (int i = 0; i < reps; ++i) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0);
    // 6. clock64 itself
    t0 = clock64();
    long long s = 0;
    for (int i = 0; i < reps; ++i) s += clock64();
    t1 = clock64();
    if (threadIdx.x == 0) { out[5] = (t1 - t0); out[6] = s + acc; }
}

int main() {
    long long *d, h[8];
    cudaMalloc(&d, 64);
    const int reps = 4096;
    k<<<1, 64>>>(d, reps);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
    const char *n[] = {"mbarrier.try_wait (phase complete)", "mbarrier.test_wait (phase complete)", "ld.acquire.shared flag", "elect.sync + __syncwarp",
                       "tcgen05.fence::after_thread_sync", "clock64"};
    printf("%s\n", cudaGetErrorString(e));
    for (int i = 0; i < 6; ++i) printf("%-40s %7.1f cycles\n", n[i], (double)h[i] / reps);
    return 0;
}
