// DMMA issue microbenchmark: how many warps per SMSP does it take to saturate the FP64 tensor
// pipe, and what do interleaved FP64 ops / shared loads cost?  Development aid.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NF64, int NLDS>
__global__ void k(double *out, int iters) {
    __shared__ double2 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double2(1.0 + i * 1e-9, 1.0 - i * 1e-9);
    __syncthreads();
    double acc[32][2];
    double a[2] = {1.0 + 1e-9 * threadIdx.x, 1.0 - 1e-9 * threadIdx.x}, s[2] = {1e-3, 2e-3};
    double r0 = 0.999999, r1 = 1e-3;
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i][0] = acc[i][1] = 0.0;
    const double2 *p = sm + (threadIdx.x & 31);
    for (int it = 0; it < iters; ++it) {
        double2 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = NLDS ? p[((it + q) & 7) * 64 + q * 8] : make_double2(a[0], a[1]);
        double na[2], ns[2];
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (NF64) { na[m] = a[m] * r0 - s[m] * r1; ns[m] = a[m] * r1 + s[m] * r0; }
        }
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                dmma884(acc[m * 16 + q * 4 + 0][0], acc[m * 16 + q * 4 + 0][1], a[m], v[q].x);
                dmma884(acc[m * 16 + q * 4 + 1][0], acc[m * 16 + q * 4 + 1][1], a[m], v[q].y);
                dmma884(acc[m * 16 + q * 4 + 2][0], acc[m * 16 + q * 4 + 2][1], s[m], v[q].x);
                dmma884(acc[m * 16 + q * 4 + 3][0], acc[m * 16 + q * 4 + 3][1], s[m], v[q].y);
            }
        if (NF64) {
#pragma unroll
            for (int m = 0; m < 2; ++m) { a[m] = na[m]; s[m] = ns[m]; }
        }
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += acc[i][0] + acc[i][1];
    if (t == 1234.5) out[0] = t;
}
template <int NF64, int NLDS>
void run(const char *name, int warps_per_sm, int sms) {
    double *d; cudaMalloc(&d, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4000;
    int threads = 128, blocks = sms * (warps_per_sm / 4);
    k<NF64, NLDS><<<blocks, threads>>>(d, 100);
    cudaEventRecord(e0);
    k<NF64, NLDS><<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double tf = (double)blocks * 4 * iters * 32 * 512.0 / (ms * 1e-3) / 1e12;
    printf("%-28s warps/SM=%2d  %7.2f TFLOP/s  (%.1f%% of 37.2)\n", name, warps_per_sm, tf, tf / 37.2 * 100);
    cudaFree(d);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    for (int w : {4, 8, 12, 16}) {
        run<0, 0>("pure DMMA", w, sms);
        run<1, 0>("DMMA + 8 FP64 rot", w, sms);
        run<0, 1>("DMMA + 4 LDS.128", w, sms);
        run<1, 1>("DMMA + rot + LDS", w, sms);
    }
    return 0;
}
