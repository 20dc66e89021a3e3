// fit_tilt on the device: per-segment least-squares piston/tip/tilt
// fit of the OPD over the segment mask and removal of the tilt terms.
//
// Replaces lentil/plane.py:564-611 (Plane.fit_tilt) and :522-562 (ptt_vector).  The reference
// builds a dense (npix x 3) design matrix per segment, [1, r*dx0, -c*dx1] * mask, and calls
// np.linalg.lstsq; rows outside the mask are zero, so the solution is that of the 3x3 normal
// equations over the masked pixels.  Here one kernel reduces the 9 moments per segment
// (warp shuffles -> shared memory -> one atomicAdd per block and moment) with coordinates
// centred on the segment's own bounding box (keeps the normal matrix well conditioned; the shift
// is undone on the host when the 3x3 system is solved), a second kernel rewrites the OPD:
//     opd' = sum_s (opd - t1_s * r*dx0 * m_s - t2_s * (-c*dx1) * m_s) * m_s      (plane.py:602-607)
#include "lfd_common.cuh"

namespace lfd {

constexpr int FT_MOMENTS = 9;   // S1, Sx, Sy, Sxx, Sxy, Syy, Sz, Sxz, Syz

struct FitSeg {
    int r0, c0, h, w, mask_index, pad_;
    double xc, yc;      // centre of the bbox in basis coordinates (x = r*dx0, y = -c*dx1)
};

__global__ void __launch_bounds__(256)
fit_tilt_moments_kernel(const double *__restrict__ opd, const uint8_t *__restrict__ mask,
                        const double *__restrict__ amp_for_mask, int n_r, int n_c, double dx0, double dx1,
                        const FitSeg *__restrict__ segs, double *__restrict__ moments) {
    const FitSeg sg = segs[blockIdx.y];
    const long long nelem = (long long)sg.h * sg.w;
    const int hr = n_r / 2, hc = n_c / 2;
    double acc[FT_MOMENTS];
#pragma unroll
    for (int k = 0; k < FT_MOMENTS; ++k) acc[k] = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nelem;
         e += (long long)gridDim.x * blockDim.x) {
        const int rr = sg.r0 + (int)(e / sg.w), cc = sg.c0 + (int)(e % sg.w);
        const long long pix = (long long)rr * n_c + cc;
        const bool in = mask ? (mask[(long long)sg.mask_index * n_r * n_c + pix] != 0) : (amp_for_mask[pix] != 0.0);
        if (!in) continue;
        const double x = (double)(rr - hr) * dx0 - sg.xc, y = -(double)(cc - hc) * dx1 - sg.yc, z = opd[pix];
        acc[0] += 1.0; acc[1] += x; acc[2] += y; acc[3] += x * x; acc[4] += x * y; acc[5] += y * y;
        acc[6] += z; acc[7] += x * z; acc[8] += y * z;
    }
    __shared__ double red[8][FT_MOMENTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < FT_MOMENTS; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < FT_MOMENTS) {
        double v = 0.0;
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        atomicAdd(moments + (size_t)blockIdx.y * FT_MOMENTS + threadIdx.x, v);
    }
}

__global__ void __launch_bounds__(256)
remove_tilt_kernel(const double *__restrict__ opd, const uint8_t *__restrict__ mask,
                   const double *__restrict__ amp_for_mask, int n_r, int n_c, int nseg, double dx0, double dx1,
                   const double *__restrict__ coef /* nseg x 3: piston, t1, t2 */, double *__restrict__ out) {
    const long long npix = (long long)n_r * n_c;
    const int hr = n_r / 2, hc = n_c / 2;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
         pix += (long long)gridDim.x * blockDim.x) {
        const int rr = (int)(pix / n_c), cc = (int)(pix % n_c);
        const double x = (double)(rr - hr) * dx0, y = -(double)(cc - hc) * dx1, z = opd[pix];
        if (nseg == 1) {
            // plane.py:590-593: opd -= t1*x*m + t2*y*m   (pixels outside the mask keep their OPD)
            const bool in = mask ? (mask[pix] != 0) : (amp_for_mask[pix] != 0.0);
            out[pix] = in ? z - (coef[1] * x + coef[2] * y) : z;
        } else {
            double v = 0.0;
            for (int s = 0; s < nseg; ++s)
                if (mask[(long long)s * npix + pix] != 0) v += z - (coef[3 * s + 1] * x + coef[3 * s + 2] * y);
            out[pix] = v;
        }
    }
}

// ---- bounding box of the support of each mask plane (lentil/util.py:190-218 boundary) -----------------
// out[4p .. 4p+3] = rmin, rmax, cmin, cmax of plane p (initialised to n_r, -1, n_c, -1 by the launcher)
template <typename T>
__global__ void __launch_bounds__(256)
bbox_kernel(const T *__restrict__ x, int n_r, int n_c, int nonzero, int *__restrict__ out) {
    const long long npix = (long long)n_r * n_c;
    const T *plane = x + (long long)blockIdx.y * npix;
    int rmin = n_r, rmax = -1, cmin = n_c, cmax = -1;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < npix;
         pix += (long long)gridDim.x * blockDim.x) {
        const T v = plane[pix];
        const bool hit = nonzero ? (v != (T)0) : (v > (T)0);
        if (hit) {
            const int r = (int)(pix / n_c), c = (int)(pix % n_c);
            rmin = min(rmin, r); rmax = max(rmax, r); cmin = min(cmin, c); cmax = max(cmax, c);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
        rmax = max(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
        cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    }
    if ((threadIdx.x & 31) == 0 && rmax >= 0) {
        int *o4 = out + 4 * blockIdx.y;
        atomicMin(o4 + 0, rmin); atomicMax(o4 + 1, rmax); atomicMin(o4 + 2, cmin); atomicMax(o4 + 3, cmax);
    }
}

__global__ void bbox_init_kernel(int *out, int nplanes, int n_r, int n_c) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < nplanes) { out[4 * p] = n_r; out[4 * p + 1] = -1; out[4 * p + 2] = n_c; out[4 * p + 3] = -1; }
}

}  // namespace lfd

using namespace lfd;

extern "C" int lfd_mask_bbox(const void *x, int32_t is_f64, int32_t nonzero, int32_t n_r, int32_t n_c,
                             int32_t nplanes, int32_t *out_dev, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFD_REQUIRE(x && out_dev && n_r > 0 && n_c > 0 && nplanes > 0, "lfd_mask_bbox: bad arguments");
    bbox_init_kernel<<<(nplanes + 127) / 128, 128, 0, stream>>>(out_dev, nplanes, n_r, n_c);
    long long npix = (long long)n_r * n_c, bx = (npix + 255) / 256;
    if (bx > sm_or_default() * 8) bx = sm_or_default() * 8;
    dim3 grid((unsigned)bx, (unsigned)nplanes);
    if (is_f64) bbox_kernel<double><<<grid, 256, 0, stream>>>((const double *)x, n_r, n_c, nonzero, out_dev);
    else bbox_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t *)x, n_r, n_c, nonzero, out_dev);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch(2);
    return 0;
}

extern "C" int lfd_fit_tilt_moments(const double *opd, const uint8_t *mask, const double *amp_for_mask,
                                    int32_t n_r, int32_t n_c, double dx0, double dx1,
                                    const lfd_segment *segs, int32_t nseg, double *moments_dev,
                                    void *scratch_dev, size_t scratch_bytes, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFD_REQUIRE(opd && (mask || amp_for_mask) && segs && moments_dev && scratch_dev, "lfd_fit_tilt_moments: NULL argument");
    LFD_REQUIRE(nseg > 0 && scratch_bytes >= (size_t)nseg * sizeof(FitSeg), "lfd_fit_tilt_moments: scratch too small");
    FitSeg *h = (FitSeg *)malloc((size_t)nseg * sizeof(FitSeg));
    LFD_REQUIRE(h != nullptr, "out of host memory");
    long long max_elem = 0;
    for (int s = 0; s < nseg; ++s) {
        h[s].r0 = segs[s].r0; h[s].c0 = segs[s].c0; h[s].h = segs[s].h; h[s].w = segs[s].w;
        h[s].mask_index = segs[s].mask_index; h[s].pad_ = 0;
        h[s].xc = ((double)segs[s].r0 + 0.5 * (segs[s].h - 1) - (double)(n_r / 2)) * dx0;
        h[s].yc = -((double)segs[s].c0 + 0.5 * (segs[s].w - 1) - (double)(n_c / 2)) * dx1;
        long long ne = (long long)segs[s].h * segs[s].w;
        if (ne > max_elem) max_elem = ne;
    }
    cudaError_t e = cudaMemcpyAsync(scratch_dev, h, (size_t)nseg * sizeof(FitSeg), cudaMemcpyHostToDevice, stream);
    free(h);
    LFD_CUDA_OK(e);
    LFD_CUDA_OK(cudaMemsetAsync(moments_dev, 0, (size_t)nseg * FT_MOMENTS * sizeof(double), stream));
    long long bx = (max_elem + 255) / 256;
    if (bx > sm_or_default() * 4) bx = sm_or_default() * 4;
    fit_tilt_moments_kernel<<<dim3((unsigned)bx, (unsigned)nseg), 256, 0, stream>>>(
        opd, mask, amp_for_mask, n_r, n_c, dx0, dx1, (const FitSeg *)scratch_dev, moments_dev);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}

extern "C" int lfd_remove_tilt(const double *opd, const uint8_t *mask, const double *amp_for_mask,
                               int32_t n_r, int32_t n_c, int32_t nseg, double dx0, double dx1,
                               const double *coef_dev, double *out, void *stream) {
    LFD_REQUIRE(opd && (mask || amp_for_mask) && coef_dev && out && nseg > 0, "lfd_remove_tilt: bad arguments");
    LFD_REQUIRE(nseg == 1 || mask, "lfd_remove_tilt: segmented planes need a mask cube");
    long long npix = (long long)n_r * n_c, bx = (npix + 255) / 256;
    if (bx > sm_or_default() * 16) bx = sm_or_default() * 16;
    remove_tilt_kernel<<<(unsigned)bx, 256, 0, (cudaStream_t)stream>>>(opd, mask, amp_for_mask, n_r, n_c, nseg, dx0, dx1,
                                                                      coef_dev, out);
    LFD_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
}
