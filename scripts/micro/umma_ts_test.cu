// tcgen05 "TS" form: D[128 x N] += A[128 x 8] * B[N x 8]^T with the A operand read from TENSOR MEMORY
// (written there by tcgen05.st, one thread per row = one TMEM lane, K along columns) and B K-major in
// shared memory (SWIZZLE_32B canonical tiles).  Validates what K2b's twiddle path relies on:
//   1. the A layout in TMEM (lane = row, column = k for 32-bit elements),
//   2. N = 48 (any multiple of 16 is legal for M = 128),
//   3. the st -> wait::st -> fence -> mbarrier -> mma ordering,
// and measures the issue rate of the 12-MMA k-block pattern of K2b (2 row tiles x 6 products).
// Development aid:  nvcc -gencode arch=compute_100a,code=sm_100a -o umma_ts_test umma_ts_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 48, KB = 8;
constexpr int SBO = 256;
constexpr int B_TILE = N * 32;   // bytes of one K-block tile of B

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline int tile_off(int r, int kc) { return (r >> 3) * SBO + (r & 7) * 32 + ((kc ^ ((r >> 2) & 1)) << 4); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;     // SWIZZLE_32B
    return d;
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}

// acc columns [0, N); A staging at column 64 + 8*kb
__global__ void __launch_bounds__(128) ts_kernel(const float *A, const float *B, float *D, int K, int reps, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t a_full, mma_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = K / KB;
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        *(float *)(smem + (k / KB) * B_TILE + tile_off(r, (k % KB) / 4) + (k % 4) * 4) = B[r * K + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(smem_u32(&a_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_done)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;

    // every thread writes its row of A: lane = tid, columns 448 + k  (K <= 64)
    {
        const uint32_t lane_addr = ((uint32_t)(warp * 32) << 16);
        for (int kb = 0; kb < nkb; ++kb) {
            float v[8];
            for (int j = 0; j < 8; ++j) v[j] = A[tid * K + kb * KB + j];
            tmem_st8(tmem + lane_addr + 448 + kb * 8, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full)) : "memory");
    }
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) {
        mbar_wait(&a_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        for (int kb = 0; kb < nkb; ++kb)
            umma_ts(tmem, tmem + 448 + kb * 8, umma_desc(smem_u32(smem + kb * B_TILE)), idesc, kb > 0 ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done)));
    }
    mbar_wait(&mma_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- rate: K2b's k-block pattern, 12 MMAs (2 row tiles x 6 products) on fixed operands ----
    if (tid == 0) {
        const long long t0 = clock64();
        for (int it = 0; it < reps; ++it) {
            const uint64_t b0 = umma_desc(smem_u32(smem + (it % nkb) * B_TILE));
#pragma unroll
            for (int rt = 0; rt < 2; ++rt) {
                const uint32_t t = tmem + rt * 4 * N, a = tmem + 448 + (it & 1) * 32;
                umma_ts(t, a, b0, idesc, 1u);
                umma_ts(t + N, a, b0, idesc, 1u);
                umma_ts(t + N, a + 8, b0, idesc, 1u);
                umma_ts(t + 2 * N, a + 16, b0, idesc, 1u);
                umma_ts(t + 3 * N, a + 16, b0, idesc, 1u);
                umma_ts(t + 3 * N, a + 24, b0, idesc, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done)));
        mbar_wait(&mma_done, 1);
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- issue-rate sweep: straight-line groups of 16 MMAs of 128 x NN x 8 cycling over NACC accumulators; A from TMEM
// (TS) or shared memory (SS).  Converged warp, one elected lane issues.
template <int NN, int TS, int NACC>
__global__ void __launch_bounds__(128) rate_kernel(int reps, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t mma_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16384; i += 128) ((float *)smem)[i] = 0.001f * (i % 7);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_done)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (warp == 0) {
        const uint64_t db = umma_desc(smem_u32(smem)), da = umma_desc(smem_u32(smem + 32768));
        const long long t0 = clock64();
        for (int it = 0; it < reps; it += 16) {
            if (elect_one()) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const uint32_t t = tmem + (q % NACC) * NN;
                    if (TS) umma_ts(t, tmem + 480 + (q & 3) * 8, db + (uint64_t)((q & 3) * 96), idesc, 1u);
                    else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                      ::"r"(t), "l"(da + (uint64_t)((q & 3) * 256)), "l"(db + (uint64_t)((q & 3) * 96)), "r"(idesc), "r"(1u));
                }
            }
            __syncwarp();
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done)));
        __syncwarp();
        mbar_wait(&mma_done, 0);
        if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- K2b's k-block exactly: 12 TS MMAs (2 row tiles x 6 products) over a 16-slot data ring, alternating twiddle
// stages, optional commits to two mbarriers per k-block (nobody waits on them)
template <int COMMITS>
__global__ void __launch_bounds__(128) kblock_kernel(int reps, long long *cycles) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t mma_done, dummy[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 16 * 6144 / 4; i += 128) ((float *)smem)[i] = 0.001f * (i % 7);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mma_done)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy[1])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    constexpr int NR = 48, B_ARR = 1536, B_BYTES = 6144;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NR >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    if (warp == 0) {
        const uint64_t desc0 = umma_desc(smem_u32(smem));
        const long long t0 = clock64();
        for (int jb = 0; jb < reps; ++jb) {
            const int s = jb & 1, sl = jb & 15;
            const uint64_t eh = desc0 + (uint64_t)((sl * B_BYTES) >> 4), el = eh + (B_ARR >> 4);
            const uint64_t oh = eh + 2 * (B_ARR >> 4), ol = eh + 3 * (B_ARR >> 4);
            const uint32_t acc = jb > 0 ? 1u : 0u;
            const uint32_t tw = tmem + 384 + s * 64;
            if (elect_one()) {
#pragma unroll
                for (int rt = 0; rt < 2; ++rt) {
                    const uint32_t t = tmem + rt * 4 * NR, a0 = tw + rt * 32;
                    umma_ts(t, a0, eh, idesc, acc);
                    umma_ts(t + NR, a0, el, idesc, acc);
                    umma_ts(t + NR, a0 + 8, eh, idesc, 1u);
                    umma_ts(t + 2 * NR, a0 + 16, oh, idesc, acc);
                    umma_ts(t + 3 * NR, a0 + 16, ol, idesc, acc);
                    umma_ts(t + 3 * NR, a0 + 24, oh, idesc, 1u);
                }
                if (COMMITS) {
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[0])));
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy[1])));
                }
            }
            __syncwarp();
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mma_done)));
        __syncwarp();
        mbar_wait(&mma_done, 0);
        if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int COMMITS>
static void run_kblock(int grid, int reps, long long *dC) {
    cudaFuncSetAttribute(kblock_kernel<COMMITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    kblock_kernel<COMMITS><<<grid, 128, 100 * 1024>>>(reps, dC);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> C(grid);
    cudaMemcpy(C.data(), dC, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : C) mx = v > mx ? v : mx;
    printf("k-block pattern, commits=%d: %.1f cycles per k-block (floor 288)  %s\n", COMMITS, (double)mx / reps, cudaGetErrorString(e));
}

template <int NN, int TS, int NACC>
static void run_rate(int grid, int reps, long long *dC) {
    cudaFuncSetAttribute(rate_kernel<NN, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    rate_kernel<NN, TS, NACC><<<grid, 128, 80 * 1024>>>(reps, dC);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> C(grid);
    cudaMemcpy(C.data(), dC, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto v : C) mx = v > mx ? v : mx;
    printf("rate N=%3d %s nacc=%d: %.1f cycles/MMA (floor %d)  %s\n", NN, TS ? "TS" : "SS", NACC, (double)mx / reps, 128 * NN / 256,
           cudaGetErrorString(e));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; float y; memcpy(&y, &u, 4); return y; }

int main() {
    const int K = 64, reps = 4000;
    std::vector<float> A(M * K), B(N * K), D(M * N), R(M * N);
    srand(1);
    for (auto &v : A) v = tf32_trunc((rand() / (float)RAND_MAX) - 0.5f);
    for (auto &v : B) v = tf32_trunc((rand() / (float)RAND_MAX) - 0.5f);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[j * K + k];
            R[i * N + j] = (float)s;
        }
    float *dA, *dB, *dD; long long *dC;
    const int grid = 148;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, grid * 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const size_t smem = (size_t)(K / KB) * B_TILE + 1024;
    cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ts_kernel<<<grid, 128, smem>>>(dA, dB, dD, K, reps, dC);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    std::vector<long long> C(grid);
    cudaMemcpy(C.data(), dC, grid * 8, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs((double)D[i] - R[i])); maxref = fmax(maxref, fabs((double)R[i])); }
    printf("TS N=%d: max |D - ref| = %.3e (max ref %.3e)  D[0]=%f ref=%f  D[%d]=%f ref=%f\n", N, maxerr, maxref, D[0], R[0],
           N + 1, D[N + 1], R[N + 1]);
    long long cmin = C[0], cmax = C[0];
    for (auto c : C) { cmin = c < cmin ? c : cmin; cmax = c > cmax ? c : cmax; }
    printf("12-MMA k-block (2 x 128 x %d x 8 x 6): %.1f .. %.1f cycles per k-block (floor 12 * 128*%d/256 = %d)\n", N,
           (double)cmin / reps, (double)cmax / reps, N, 12 * 128 * N / 256);
    run_kblock<0>(grid, reps, dC); run_kblock<1>(grid, reps, dC);
    run_rate<48, 1, 8>(grid, reps, dC); run_rate<48, 1, 1>(grid, reps, dC); run_rate<48, 0, 8>(grid, reps, dC);
    run_rate<64, 1, 4>(grid, reps, dC); run_rate<64, 0, 4>(grid, reps, dC); run_rate<96, 1, 4>(grid, reps, dC);
    run_rate<128, 1, 3>(grid, reps, dC); run_rate<128, 0, 3>(grid, reps, dC); run_rate<256, 1, 1>(grid, reps, dC);
    run_rate<256, 0, 1>(grid, reps, dC); run_rate<16, 1, 8>(grid, reps, dC); run_rate<32, 1, 8>(grid, reps, dC);
    return 0;
}
