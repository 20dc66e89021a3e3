// Wake-up latency of a warp blocked in mbarrier.try_wait (vs spinning on test_wait): warp 1 arrives on the barrier at a
// recorded clock, warp 0 records the clock right after its wait returns.  Development aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(long long *out, int reps) {
    __shared__ __align__(8) uint64_t bar;
    __shared__ long long t_arrive;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    long long sum = 0, mx = 0;
    for (int i = 0; i < reps; ++i) {
        if (threadIdx.x < 32) {
            uint32_t done = 0;
            while (!done) {
                if (MODE == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)(i & 1)) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)(i & 1)) : "memory");
            }
            const long long t = clock64();
            __syncwarp();
            if (threadIdx.x == 0) { const long long d = t - *(volatile long long *)&t_arrive; sum += d; mx = d > mx ? d : mx; }
        } else if (threadIdx.x == 32) {
            // let the waiter block for a while (different delays probe different suspend windows)
            const long long t0 = clock64();
            while (clock64() - t0 < 300 + (i % 7) * 500) { }
            *(volatile long long *)&t_arrive = clock64();
            __threadfence_block();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = sum; out[1] = mx; }
}
// several warps blocked in try_wait on the SAME mbarrier: when does each resume?
__global__ void kmulti(long long *out, int reps) {
    __shared__ __align__(8) uint64_t bar;
    __shared__ long long t_arrive;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long sum = 0, mx = 0;
    for (int i = 0; i < reps; ++i) {
        if (warp < 5) {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&bar)), "r"((uint32_t)(i & 1)) : "memory");
            const long long t = clock64();
            __syncwarp();
            if (lane == 0) { const long long d = t - *(volatile long long *)&t_arrive; sum += d; mx = d > mx ? d : mx; }
        } else if (threadIdx.x == 160) {
            const long long t0 = clock64();
            while (clock64() - t0 < 300 + (i % 7) * 500) { }
            *(volatile long long *)&t_arrive = clock64();
            __threadfence_block();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncthreads();
    }
    if (warp < 5 && lane == 0) { out[2 * warp] = sum / reps; out[2 * warp + 1] = mx; }
}
int main() {
    {
        long long *d, h[10];
        cudaMalloc(&d, sizeof(h));
        kmulti<<<1, 192>>>(d, 2000); cudaDeviceSynchronize(); cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("5 warps in try_wait on one mbarrier: arrive -> resume, mean/max:");
        for (int w = 0; w < 5; ++w) printf("  w%d %lld/%lld", w, h[2 * w], h[2 * w + 1]);
        printf("\n");
    }
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const int reps = 2000;
    k<0><<<1, 64>>>(d, reps); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("try_wait : arrive -> waiter resumes: mean %.1f cycles, max %lld\n", (double)h[0] / reps, h[1]);
    k<1><<<1, 64>>>(d, reps); cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("test_wait: arrive -> waiter resumes: mean %.1f cycles, max %lld  (%s)\n", (double)h[0] / reps, h[1], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
