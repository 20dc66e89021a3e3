// K3 — coherent merge of overlapping output windows + |E|^2 (or complex) accumulate.
//
// Replaces lentil/field.py:231-305 (insert: place a window at out//2 - shape//2 + offset, clip,
// `out[slc] += |E**2| * weight`), lentil/field.py:308-346 + :413-461 (reduce/merge: overlapping
// windows of one wavefront are summed as complex fields first) and lentil/wavefront.py:114-165
// (intensity / insert / field).
//
// Owner-computes: one thread per output pixel walks the window list in order, summing the
// complex fields of a group (= one wavefront: all segments at one wavelength / field point),
// squaring at group boundaries and accumulating weight * |sum|^2.  A pixel no window covers
// gets += 0, which is what merging into a zero-filled bounding box produces in the reference.
// No atomics, so the result is independent of scheduling.  HBM-bound: 16 B read per covered
// (pixel, window), one 8 B read-modify-write per pixel.
#include "lfd_common.cuh"

namespace lfd {

constexpr int WIN_CHUNK = 64;

template <bool INTENSITY>
__global__ void __launch_bounds__(256)
accum_kernel(const lfd_window *__restrict__ wins, int nwin, double *__restrict__ out, int H, int W,
             long long ldo) {
    __shared__ lfd_window sw[WIN_CHUNK];
    const int c = blockIdx.x * 64 + (threadIdx.x & 63);
    const int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    const bool live = (r < H) && (c < W);

    double acc_r = 0.0, acc_i = 0.0;   // INTENSITY: acc_r only
    double sr = 0.0, si = 0.0, wgt = 0.0;
    int group = INT_MIN;

    for (int v0 = 0; v0 < nwin; v0 += WIN_CHUNK) {
        int nv = min(WIN_CHUNK, nwin - v0);
        __syncthreads();
        // cooperative copy of the descriptors (48 B each) as 8-byte words
        for (int i = threadIdx.x; i < nv * (int)(sizeof(lfd_window) / 8); i += blockDim.x)
            reinterpret_cast<unsigned long long *>(sw)[i] =
                reinterpret_cast<const unsigned long long *>(wins + v0)[i];
        __syncthreads();
        if (!live) continue;
        for (int v = 0; v < nv; ++v) {
            const lfd_window &w = sw[v];
            if (INTENSITY && w.group != group) {
                acc_r += wgt * (sr * sr + si * si);
                sr = si = 0.0;
                group = w.group;
                wgt = w.weight;
            }
            int rr = r - w.r0, cc = c - w.c0;
            if (rr >= 0 && rr < w.h && cc >= 0 && cc < w.w) {
                double2 e;
                if (w.c64 == 2) {
                    // float64 intensity window: a wavefront with a single Field, already squared by the MFT epilogue
                    acc_r += w.weight * reinterpret_cast<const double *>(w.E)[(long long)rr * w.ld + cc];
                    continue;
                } else if (w.c64) {
                    const float2 f = reinterpret_cast<const float2 *>(w.E)[(long long)rr * w.ld + cc];
                    e = make_double2((double)f.x, (double)f.y);
                } else {
                    e = reinterpret_cast<const double2 *>(w.E)[(long long)rr * w.ld + cc];
                }
                if (INTENSITY) {
                    sr += e.x;
                    si += e.y;
                } else {
                    acc_r += w.weight * e.x;
                    acc_i += w.weight * e.y;
                }
            }
        }
    }
    if (!live) return;
    if (INTENSITY) {
        acc_r += wgt * (sr * sr + si * si);
        out[(long long)r * ldo + c] += acc_r;
    } else {
        double2 *o = reinterpret_cast<double2 *>(out) + (long long)r * ldo + c;
        double2 cur = *o;
        *o = make_double2(cur.x + acc_r, cur.y + acc_i);
    }
}

// Fast path of the polychromatic sum: every window is a dense, full-frame float64 intensity plane (what the fused K2a
// epilogue writes for a single-Field wavefront), so the merge is a weighted sum of nwin planes, out += sum_v w_v I_v.
// Pure HBM streaming: two pixels per thread (16-byte loads), eight planes in flight per thread, window pointers and
// weights in shared memory.  The sum runs over v in the same order and with the same contraction as accum_kernel, so
// both paths give bit-identical images.
// FIELDS: the windows are dense complex128 fields that all cover the SAME rectangle of the output, one per wavefront (nothing
// to merge coherently): out[rect] += sum_v w_v |E_v|^2, one pixel per thread — what the per-wavelength drop-in loop produces
// (Wavefront.insert of a single-Field wavefront) and what a field point of a multi-field-point batch accumulates per chunk.
constexpr int FULL_CHUNK = 512;      // windows per launch of the fast path (one accumulation per pixel, as in accum_kernel); more -> generic path
template <bool FIELDS>
__global__ void __launch_bounds__(256)
accum_full_kernel(const lfd_window *__restrict__ wins, int nwin, double *__restrict__ out, long long npairs,
                  int r0 = 0, int c0 = 0, int w = 0, long long ldo = 0) {
    __shared__ const double2 *sE[FULL_CHUNK];
    __shared__ double sw[FULL_CHUNK];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int v0 = 0; v0 < nwin; v0 += FULL_CHUNK) {
        const int nv = min(FULL_CHUNK, nwin - v0);
        __syncthreads();
        for (int i = threadIdx.x; i < nv; i += blockDim.x) {
            sE[i] = reinterpret_cast<const double2 *>(wins[v0 + i].E);
            sw[i] = wins[v0 + i].weight;
        }
        __syncthreads();
        for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < npairs; e += stride) {
            if constexpr (FIELDS) {
                double a = 0.0;
                int v = 0;
                for (; v + 8 <= nv; v += 8) {
                    double2 x[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) x[k] = __ldcs(sE[v + k] + e);
#pragma unroll
                    for (int k = 0; k < 8; ++k) a += sw[v + k] * (x[k].x * x[k].x + x[k].y * x[k].y);
                }
                for (; v < nv; ++v) {
                    const double2 x = __ldcs(sE[v] + e);
                    a += sw[v] * (x.x * x.x + x.y * x.y);
                }
                const long long rr = e / w;                    // pixel e of the (dense, ld = w) windows -> output (r0 + rr, c0 + cc)
                out[(r0 + rr) * ldo + c0 + (e - rr * w)] += a;
                continue;
            }
            double a0 = 0.0, a1 = 0.0;
            int v = 0;
            for (; v + 8 <= nv; v += 8) {
                double2 x[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) x[k] = __ldcs(sE[v + k] + e);          // streamed once: evict first
#pragma unroll
                for (int k = 0; k < 8; ++k) { a0 += sw[v + k] * x[k].x; a1 += sw[v + k] * x[k].y; }
            }
            for (; v < nv; ++v) {
                const double2 x = __ldcs(sE[v] + e);
                a0 += sw[v] * x.x; a1 += sw[v] * x.y;
            }
            double2 *o = reinterpret_cast<double2 *>(out) + e;
            const double2 cur = *o;
            *o = make_double2(cur.x + a0, cur.y + a1);
        }
    }
}

// kind 2: dense full-frame float64 intensity planes; kind 0: dense complex128 fields on one common rectangle inside the
// output, every window its own wavefront (strictly increasing groups)
static bool all_full_frame(const lfd_window *wins, int nwin, const void *out, int H, int W, int64_t ldo, int kind) {
    if (nwin > FULL_CHUNK || ((uintptr_t)out & 15)) return false;
    if (kind == 2 && (ldo != W || (((long long)H * W) & 1))) return false;
    for (int v = 0; v < nwin; ++v) {
        const lfd_window &w = wins[v];
        if (w.c64 != kind || w.ld != w.w || ((uintptr_t)w.E & 15)) return false;
        if (kind == 2 && (w.r0 != 0 || w.c0 != 0 || w.h != H || w.w != W)) return false;
        if (kind == 0) {
            if (w.r0 < 0 || w.c0 < 0 || w.r0 + w.h > H || w.c0 + w.w > W) return false;
            if (w.r0 != wins[0].r0 || w.c0 != wins[0].c0 || w.h != wins[0].h || w.w != wins[0].w) return false;
            if (v > 0 && w.group <= wins[v - 1].group) return false;
        }
    }
    return true;
}

static int launch_accum(bool intensity, const lfd_window *wins, int32_t nwin, void *out, int32_t H,
                        int32_t W, int64_t ldo, void *scratch, size_t scratch_bytes,
                        cudaStream_t stream) {
    if (nwin == 0) return 0;
    LFD_REQUIRE(wins && out && scratch, "lfd_accum: NULL argument");
    LFD_REQUIRE(H > 0 && W > 0 && ldo >= W, "lfd_accum: bad output shape");
    LFD_REQUIRE(scratch_bytes >= (size_t)nwin * sizeof(lfd_window),
                "lfd_accum: scratch too small (%zu < %zu)", scratch_bytes,
                (size_t)nwin * sizeof(lfd_window));
    for (int v = 0; v < nwin; ++v) {
        LFD_REQUIRE(wins[v].E && wins[v].h > 0 && wins[v].w > 0 && wins[v].ld >= wins[v].w,
                    "lfd_accum: window %d malformed", v);
        LFD_REQUIRE(v == 0 || wins[v].group >= wins[v - 1].group,
                    "lfd_accum: window groups must be non-decreasing");
    }
    LFD_CUDA_OK(cudaMemcpyAsync(scratch, wins, (size_t)nwin * sizeof(lfd_window),
                                cudaMemcpyHostToDevice, stream));
    const bool planes = intensity && all_full_frame(wins, nwin, out, H, W, ldo, 2);
    const bool fields = intensity && !planes && all_full_frame(wins, nwin, out, H, W, ldo, 0);
    if (planes || fields) {
        const long long nelem = planes ? (long long)H * W / 2 : (long long)wins[0].h * wins[0].w;
        long long blocks = (nelem + 255) / 256;
        const long long cap = (long long)sm_or_default() * 8;
        if (blocks > cap) blocks = cap;
        if (planes) accum_full_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>((const lfd_window *)scratch, nwin, (double *)out, nelem);
        else accum_full_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>((const lfd_window *)scratch, nwin, (double *)out, nelem,
                                                                           wins[0].r0, wins[0].c0, wins[0].w, (long long)ldo);
        LFD_CUDA_OK(cudaGetLastError());
        count_launch();
        return 0;
    }
    dim3 grid((W + 63) / 64, (H + 3) / 4);
    if (intensity)
        accum_kernel<true><<<grid, 256, 0, stream>>>((const lfd_window *)scratch, nwin, (double *)out, H, W, ldo);
    else
        accum_kernel<false><<<grid, 256, 0, stream>>>((const lfd_window *)scratch, nwin, (double *)out, H, W, ldo);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

__global__ void __launch_bounds__(256)
field_mul_kernel(const double2 *__restrict__ a, long long lda, const double2 *__restrict__ b,
                 long long ldb, double sr, double si, double2 *__restrict__ out, long long ldo, int h,
                 int w) {
    const long long n = (long long)h * w;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        int r = (int)(e / w), c = (int)(e % w);
        double2 x = a[r * lda + c];
        if (b != nullptr) {
            double2 y = b[r * ldb + c];
            x = make_double2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
        }
        out[r * ldo + c] = make_double2(x.x * sr - x.y * si, x.x * si + x.y * sr);
    }
}

}  // namespace lfd

extern "C" int lfd_field_mul(const void *a, int64_t lda, const void *b, int64_t ldb, double s_re,
                             double s_im, void *out, int64_t ldo, int32_t h, int32_t w, void *stream) {
    LFD_REQUIRE(a && out && h > 0 && w > 0 && lda >= w && ldo >= w && (!b || ldb >= w),
                "lfd_field_mul: bad arguments");
    long long n = (long long)h * w, blocks = (n + 255) / 256;
    if (blocks > lfd::sm_or_default() * 8) blocks = lfd::sm_or_default() * 8;
    lfd::field_mul_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const double2 *)a, lda, (const double2 *)b, ldb, s_re, s_im, (double2 *)out, ldo, h, w);
    LFD_CUDA_OK(cudaGetLastError());
    lfd::count_launch();
    return 0;
}

extern "C" int lfd_accum_intensity(const lfd_window *wins, int32_t nwin, double *I, int32_t H,
                                   int32_t W, int64_t ldI, void *scratch, size_t scratch_bytes,
                                   void *stream) {
    return lfd::launch_accum(true, wins, nwin, I, H, W, ldI, scratch, scratch_bytes,
                             (cudaStream_t)stream);
}

extern "C" int lfd_accum_field(const lfd_window *wins, int32_t nwin, void *out, int32_t H, int32_t W,
                               int64_t ldo, void *scratch, size_t scratch_bytes, void *stream) {
    return lfd::launch_accum(false, wins, nwin, out, H, W, ldo, scratch, scratch_bytes,
                             (cudaStream_t)stream);
}
