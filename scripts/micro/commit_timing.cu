// When does a tcgen05.commit signal its mbarrier while MORE MMAs are queued behind it?  Warp 0 issues NBLK k-blocks
// (6 x 128x96x8 tf32 TS MMAs + one commit each, to distinct mbarriers) back to back; warp 1 waits on the mbarriers in
// order and records the clock of each completion.  Development aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NBLK = 16, NR = 96, SBO = 256;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)1 << 16; d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46; d |= (uint64_t)6 << 61;
    return d;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__global__ void __launch_bounds__(64) k(long long *out, int gap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar[NBLK];
    __shared__ uint32_t tmem_base_s;
    __shared__ long long t_issue[NBLK];
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 12288; i += 64) ((float *)smem)[i] = 0.001f * (i % 7);
    if (tid == 0) { for (int i = 0; i < NBLK; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i]))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long t0 = clock64();
    if (warp == 0) {
        const uint64_t db = umma_desc(smem_u32(smem));
#pragma unroll
        for (int b = 0; b < NBLK; ++b) {
            if (tid == 0) t_issue[b] = clock64() - t0;
            uint32_t pred;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            if (pred) {
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tmem + (q & 3) * NR), "r"(tmem + 384 + (q & 3) * 8), "l"(db + (uint64_t)(q * 192)), "r"(idesc), "r"(1u));
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[b])));
            }
            __syncwarp();
            if (gap) { const long long t = clock64(); while (clock64() - t < gap) { } }
        }
    } else {
        for (int b = 0; b < NBLK; ++b) {
            mbar_wait(&bar[b], 0);
            if (tid == 32) out[b] = clock64() - t0;
        }
    }
    __syncthreads();
    if (tid == 0) for (int b = 0; b < NBLK; ++b) out[NBLK + b] = t_issue[b];
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
    long long *d, h[2 * NBLK];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int gap : {0, 200, 400}) {
        k<<<1, 64, 64 * 1024>>>(d, gap);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("gap %d (%s)\n  issued at :", gap, cudaGetErrorString(e));
        for (int b = 0; b < NBLK; ++b) printf(" %6lld", h[NBLK + b]);
        printf("\n  commit at :");
        for (int b = 0; b < NBLK; ++b) printf(" %6lld", h[b]);
        printf("\n");
    }
    return 0;
}
