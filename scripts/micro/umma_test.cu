// Minimal tcgen05 (kind::tf32, cta_group::1) GEMM: D[128 x N] = A[128 x K] * B[N x K]^T with both
// operands K-major in the no-swizzle ("interleave") canonical shared-memory layout, accumulators
// in TMEM, read back with tcgen05.ld.  Validates the descriptor encodings used by K2b.
// Development aid:  nvcc -gencode arch=compute_100a,code=sm_100a -o umma_test umma_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 128, KB = 8;   // one MMA = 128 x 128 x 8 (tf32)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B; LBO between the two K chunks of
// an MMA (K=8 tf32 = 2 x 16 B), SBO between 8-row groups.
__host__ __device__ inline int canon_off_bytes(int row, int k, int lbo, int sbo) {
    return (row / 8) * sbo + (k / 4) * lbo + (row % 8) * 16 + (k % 4) * 4;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;            // version = 1 (Blackwell)
    // base_offset = 0, lbo_mode = 0, layout_type = 0 (SWIZZLE_NONE)
    return d;
}

__global__ void __launch_bounds__(128) umma_kernel(const float *A, const float *B, float *D, int K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int nkb = K / KB;
    const int LBO = 128, SBO = 256;               // per k-block tile: 128 rows x 8 k = 4 KB
    unsigned char *sA = smem, *sB = smem + (size_t)nkb * 4096;

    // fill operands (generic proxy)
    for (int i = tid; i < M * K; i += 128) {
        int r = i / K, k = i % K;
        *(float *)(sA + (k / KB) * 4096 + canon_off_bytes(r, k % KB, LBO, SBO)) = A[r * K + k];
    }
    for (int i = tid; i < N * K; i += 128) {
        int r = i / K, k = i % K;
        *(float *)(sB + (k / KB) * 4096 + canon_off_bytes(r, k % KB, LBO, SBO)) = B[r * K + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");     // make generic-proxy smem writes visible to the MMA
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        // instruction descriptor: c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int kb = 0; kb < nkb; ++kb) {
            uint64_t da = make_desc(smem_u32(sA + kb * 4096), LBO, SBO);
            uint64_t db = make_desc(smem_u32(sB + kb * 4096), LBO, SBO);
            uint32_t acc = kb > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)));
    }
    // wait for the MMAs
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    // each warp reads its 32 lanes x 128 columns, 32 columns at a time
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

static float tf32_round(float x) {
    uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y;
}

int main() {
    const int K = 64;
    std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N);
    srand(1);
    for (auto &v : A) v = tf32_round((rand() / (float)RAND_MAX) - 0.5f);
    for (auto &v : B) v = tf32_round((rand() / (float)RAND_MAX) - 0.5f);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[j * K + k];
            R[i * N + j] = (float)s;
        }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    size_t smem = (size_t)2 * (K / KB) * 4096;
    cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_kernel<<<1, 128, smem>>>(dA, dB, dD, K);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs((double)D[i] - R[i])); maxref = fmax(maxref, fabs((double)R[i])); }
    printf("max |D - ref| = %.3e (max ref %.3e)  D[0]=%f ref=%f  D[129]=%f ref=%f\n", maxerr, maxref, D[0], R[0], D[129], R[129]);
    return 0;
}
