// The twiddle generator's per-k-block body (fp64 rotation, 8 fp32 in-block factors, tf32 hi/lo split, tcgen05.st of 32
// columns, wait::st) in isolation: cycles per iteration for 1, 2 and 3 warps per SM sub-partition.  Development aid.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
template <int MODE>   // 0: compute + st + wait, 1: compute only (results kept alive), 2: st + wait only
__global__ void k(long long *out, const float2 *Sg, int reps) {
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    float2 S[8];
    for (int j = 0; j < 8; ++j) S[j] = Sg[(threadIdx.x * 8 + j) % 1024];
    double Wc = 0.8 + 1e-3 * lane, Ws = sqrt(1.0 - Wc * Wc), Rc = cos(1e-3 * (lane + 1)), Rs = sin(1e-3 * (lane + 1));
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384 + ((warp >> 2) & 3) * 32;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < reps; ++it) {
        float v[32];
        if (MODE != 2) {
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
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = acc + j;
        }
        if (MODE != 1) {
            tmem_st32(taddr, v);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;");
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += v[j];
        }
    }
    const long long dt = clock64() - t0;
    if (lane == 0) out[warp] = dt / reps;
    if (acc == 123.f) out[100] = 1;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}
int main() {
    long long *d, h[16];
    float2 *S;
    cudaMalloc(&d, 1024); cudaMalloc(&S, 1024 * 8); cudaMemset(S, 0x3c, 1024 * 8);
    for (int warps : {4, 8, 12}) {
        k<0><<<1, warps * 32>>>(d, S, 2000); cudaDeviceSynchronize(); cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%2d warps, compute + st + wait : %lld cycles/iteration (warp 0), %lld (last warp)   %s\n", warps, h[0], h[warps - 1], cudaGetErrorString(cudaGetLastError()));
        k<1><<<1, warps * 32>>>(d, S, 2000); cudaDeviceSynchronize(); cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%2d warps, compute only        : %lld cycles/iteration\n", warps, h[0]);
        k<2><<<1, warps * 32>>>(d, S, 2000); cudaDeviceSynchronize(); cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%2d warps, st + wait only      : %lld cycles/iteration\n", warps, h[0]);
    }
    return 0;
}
