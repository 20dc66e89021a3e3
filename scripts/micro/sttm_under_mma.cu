// Latency of tcgen05.st (32 columns) + tcgen05.wait::st for 4 warps, with the tensor pipe idle and with a continuous
// stream of 128x96x8 tf32 TS MMAs running (issued by a fifth warp).  Development aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NR = 96, SBO = 256;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)1 << 16; d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46; d |= (uint64_t)6 << 61;
    return d;
}
__global__ void __launch_bounds__(160) k(long long *out, int with_mma, int reps) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 12288; i += 160) ((float *)smem)[i] = 0.001f * (i % 7);
    if (tid == 0) stop = 0;
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 4) {
        if (with_mma) {
            const uint64_t db = umma_desc(smem_u32(smem));
            while (!stop) {
                uint32_t pred;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
                if (pred) {
#pragma unroll
                    for (int q = 0; q < 6; ++q)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(tmem + (q & 3) * NR), "r"(tmem + 384 + (q & 3) * 8), "l"(db + (uint64_t)(q * 192)), "r"(idesc), "r"(1u));
                }
                __syncwarp();
            }
        }
    } else {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 448;     // columns the MMAs do not read
        uint32_t v = lane;
        long long sum = 0, mx = 0;
        for (int i = 0; i < reps; ++i) {
            const long long t = clock64();
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};"
                         ::"r"(taddr), "r"(v) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            const long long dt = clock64() - t;
            sum += dt; mx = dt > mx ? dt : mx;
            // idle a little so that the four warps do not run in lock step
            const long long t2 = clock64(); while (clock64() - t2 < 150 + 37 * warp) { }
        }
        if (lane == 0) { out[warp * 2] = sum / reps; out[warp * 2 + 1] = mx; }
        __syncwarp();
        if (tid == 0) stop = 1;
    }
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// the generator body (fp64 rotation, fp32 factors, hi/lo split, tcgen05.st x32, wait::st) on 8 warps while warp 8 streams MMAs
__global__ void __launch_bounds__(288) kgen(long long *out, int with_mma, int reps) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 12288; i += 288) ((float *)smem)[i] = 0.001f * (i % 7);
    if (tid == 0) stop = 0;
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 8) {
        if (with_mma) {
            const uint64_t db = umma_desc(smem_u32(smem));
            while (!stop) {
                uint32_t pred;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
                if (pred) {
#pragma unroll
                    for (int q = 0; q < 6; ++q)
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                                     ::"r"(tmem + (q & 3) * NR), "r"(tmem + 480 + (q & 3) * 8), "l"(db + (uint64_t)(q * 192)), "r"(idesc), "r"(1u));
                }
                __syncwarp();
            }
        }
    } else {
        float2 S[8];
        for (int j = 0; j < 8; ++j) S[j] = make_float2(0.9f - 0.01f * j, 0.1f + 0.01f * j + 1e-3f * lane);
        double Wc = 0.8 + 1e-3 * lane, Ws = sqrt(1.0 - Wc * Wc), Rc = cos(1e-3 * (lane + 1)), Rs = sin(1e-3 * (lane + 1));
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384 + (warp >> 2) * 32;
        const long long t0 = clock64();
        for (int it = 0; it < reps; ++it) {
            float v[32];
            const float wc = (float)Wc, ws = (float)Ws;
            v[0] = tf32_hi(wc); v[8] = wc - v[0]; v[16] = tf32_hi(ws); v[24] = ws - v[16];
#pragma unroll
            for (int j = 1; j < 8; ++j) {
                const float c = wc * S[j].x - ws * S[j].y, sn = wc * S[j].y + ws * S[j].x;
                v[j] = tf32_hi(c); v[8 + j] = c - v[j];
                v[16 + j] = tf32_hi(sn); v[24 + j] = sn - v[16 + j];
            }
            const double nc = Wc * Rc - Ws * Rs, ns = Wc * Rs + Ws * Rc;
            Wc = nc; Ws = ns;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                         ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
                           "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
                           "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
                           "f"(v[30]), "f"(v[31]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;");
        }
        const long long dt = clock64() - t0;
        if (lane == 0) out[warp] = dt / reps;
        __syncwarp();
        asm volatile("bar.sync 1, 256;");
        if (tid == 0) stop = 1;
    }
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
    {
        long long *d, h[8];
        cudaMalloc(&d, sizeof(h));
        cudaFuncSetAttribute(kgen, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        for (int m = 0; m < 2; ++m) {
            kgen<<<1, 288, 64 * 1024>>>(d, m, 2000);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("generator body on 8 warps, MMA stream %s (%s): cycles/iteration per warp:", m ? "ON " : "off", cudaGetErrorString(e));
            for (int w = 0; w < 8; ++w) printf(" %lld", h[w]);
            printf("\n");
        }
    }
    long long *d, h[8];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int m = 0; m < 2; ++m) {
        k<<<1, 160, 64 * 1024>>>(d, m, 2000);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("MMA stream %s (%s): tcgen05.st x32 + wait::st, mean/max cycles per warp:", m ? "ON " : "off", cudaGetErrorString(e));
        for (int w = 0; w < 4; ++w) printf("  w%d %lld/%lld", w, h[2 * w], h[2 * w + 1]);
        printf("\n");
    }
    return 0;
}
