// K1 — pupil prep: phasor = amp * mask_s * exp(+2 pi i opd / lambda) on each segment's bounding
// box, for a whole list of wavelengths in one launch.
//
// Replaces the per-wavelength numpy temporaries of lentil/plane.py:494-510 (Plane.__mul__):
//     amp = self.amplitude[s] * mask[s];  opd = self.opd[s]
//     phasor = amp * np.exp(2*np.pi*1j*opd/wavelength)
// HBM-bound by design: amp/opd/mask are read once per pixel and reused for every wavelength in
// the chunk; the only traffic that scales with the wavelength count is the 16 B/pixel store.
#include "lfd_common.cuh"

namespace lfd {

constexpr int PP_MAX_SEG = 32;
constexpr int PP_MAX_LAM = 128;

struct PupilPrepParams {
    lfd_segment seg[PP_MAX_SEG];
    double inv_lam[PP_MAX_LAM];      // 1 / wavelength, rounded once on the host
    int nseg, nlam;
};

template <typename CT> __device__ __forceinline__ CT make_cplx(double re, double im);
template <> __device__ __forceinline__ double2 make_cplx<double2>(double re, double im) { return make_double2(re, im); }
template <> __device__ __forceinline__ float2 make_cplx<float2>(double re, double im) { return make_float2((float)re, (float)im); }

// CT = double2 (complex128, the reference's type) or float2 (complex64, input of the K2b path)
template <typename CT>
__global__ void __launch_bounds__(256)
pupil_prep_kernel(const double *__restrict__ amp, const double *__restrict__ opd,
                  const uint8_t *__restrict__ mask, int n_r, int n_c,
                  const __grid_constant__ PupilPrepParams P, CT *__restrict__ out,
                  long long lam_stride) {
    const lfd_segment &sg = P.seg[blockIdx.y];
    const long long nelem = (long long)sg.h * sg.w;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nelem;
         e += (long long)gridDim.x * blockDim.x) {
        int rr = (int)(e / sg.w), cc = (int)(e % sg.w);
        long long pix = (long long)(sg.r0 + rr) * n_c + (sg.c0 + cc);
        double a = amp[pix];
        if (mask != nullptr && mask[(long long)sg.mask_index * n_r * n_c + pix] == 0) a = 0.0;
        double o = opd[pix];
        CT *dst = out + sg.out_offset + e;
#pragma unroll 4
        for (int l = 0; l < P.nlam; ++l) {
            double tcyc = o * P.inv_lam[l];       // phase in cycles (a multiplication by the rounded reciprocal instead of a division:
                                                  // at most one ulp of the phase, ~1e-16 rad for physical OPDs; the fused loads of K2a do the same)
            double r = tcyc - rint(tcyc);          // exact: |r| <= 0.5
            if constexpr (sizeof(CT) == sizeof(float2)) {
                // complex64 output: fp32 sine/cosine of the fp64-reduced phase (as in the fused fold of mft_c64.cu)
                float s, c;
                sincospif(2.0f * (float)r, &s, &c);
                const float af = (float)a;
                dst[(long long)l * lam_stride] = make_float2(af * c, af * s);
            } else {
                double s, c;
                cis_unit(r, c, s);
                dst[(long long)l * lam_stride] = make_cplx<CT>(a * c, a * s);
            }
        }
    }
}

}  // namespace lfd

using namespace lfd;

static int pupil_prep_impl(bool c64, const double *amp, const double *opd, const uint8_t *mask,
                           int32_t n_r, int32_t n_c, const lfd_segment *segs, int32_t nseg,
                           const double *wavelengths, int32_t nlam, void *out,
                           int64_t out_lam_stride, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LFD_REQUIRE(amp && opd && out && segs && wavelengths, "lfd_pupil_prep: NULL argument");
    LFD_REQUIRE(n_r > 0 && n_c > 0 && nseg > 0 && nlam > 0, "lfd_pupil_prep: empty problem");
    for (int s = 0; s < nseg; ++s)
        LFD_REQUIRE(segs[s].r0 >= 0 && segs[s].c0 >= 0 && segs[s].h > 0 && segs[s].w > 0 &&
                        segs[s].r0 + segs[s].h <= n_r && segs[s].c0 + segs[s].w <= n_c,
                    "lfd_pupil_prep: segment %d bbox outside the %d x %d pupil", s, n_r, n_c);
    for (int l = 0; l < nlam; ++l)
        LFD_REQUIRE(wavelengths[l] != 0.0, "lfd_pupil_prep: wavelength %d is zero", l);

    for (int s0 = 0; s0 < nseg; s0 += PP_MAX_SEG) {
        for (int l0 = 0; l0 < nlam; l0 += PP_MAX_LAM) {
            PupilPrepParams P;
            P.nseg = (nseg - s0 < PP_MAX_SEG) ? nseg - s0 : PP_MAX_SEG;
            P.nlam = (nlam - l0 < PP_MAX_LAM) ? nlam - l0 : PP_MAX_LAM;
            long long max_elem = 0;
            for (int s = 0; s < P.nseg; ++s) {
                P.seg[s] = segs[s0 + s];
                long long ne = (long long)P.seg[s].h * P.seg[s].w;
                if (ne > max_elem) max_elem = ne;
            }
            for (int l = 0; l < P.nlam; ++l) P.inv_lam[l] = 1.0 / wavelengths[l0 + l];
            long long bx = (max_elem + 255) / 256;
            if (bx > sm_or_default() * 16) bx = sm_or_default() * 16;  // grid-stride beyond 16 CTAs per SM
            dim3 grid((unsigned)bx, (unsigned)P.nseg);
            if (c64)
                pupil_prep_kernel<float2><<<grid, 256, 0, stream>>>(amp, opd, mask, n_r, n_c, P,
                                                                    (float2 *)out + (long long)l0 * out_lam_stride,
                                                                    out_lam_stride);
            else
                pupil_prep_kernel<double2><<<grid, 256, 0, stream>>>(amp, opd, mask, n_r, n_c, P,
                                                                     (double2 *)out + (long long)l0 * out_lam_stride,
                                                                     out_lam_stride);
            LFD_CUDA_OK(cudaGetLastError());
            count_launch();
        }
    }
    return 0;
}

extern "C" int lfd_pupil_prep(const double *amp, const double *opd, const uint8_t *mask,
                              int32_t n_r, int32_t n_c, const lfd_segment *segs, int32_t nseg,
                              const double *wavelengths, int32_t nlam, void *out,
                              int64_t out_lam_stride, void *stream) {
    return pupil_prep_impl(false, amp, opd, mask, n_r, n_c, segs, nseg, wavelengths, nlam, out, out_lam_stride, stream);
}

extern "C" int lfd_pupil_prep_c64(const double *amp, const double *opd, const uint8_t *mask,
                                  int32_t n_r, int32_t n_c, const lfd_segment *segs, int32_t nseg,
                                  const double *wavelengths, int32_t nlam, void *out,
                                  int64_t out_lam_stride, void *stream) {
    return pupil_prep_impl(true, amp, opd, mask, n_r, n_c, segs, nseg, wavelengths, nlam, out, out_lam_stride, stream);
}
