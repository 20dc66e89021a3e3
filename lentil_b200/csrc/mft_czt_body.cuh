// Device code of the chirp-z row transform, compiled once per precision: mft_czt.cu includes this file inside
// namespace czt::f64 (RL = double, V2 = double2) and inside namespace czt::f32 (RL = float, V2 = float2).  The including
// namespace provides RL, V2, mk2(), tw_table(), phasor() and REG_THREADS (threads per SM the register allocation must allow).
// No include guard on purpose.

__device__ __forceinline__ V2 cmul(V2 a, V2 b) { return mk2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ V2 cadd(V2 a, V2 b) { return mk2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ V2 csub(V2 a, V2 b) { return mk2(a.x - b.x, a.y - b.y); }
// multiply by -i (S = +1: forward transform) or by +i (S = -1: inverse)
template <int S> __device__ __forceinline__ V2 mul_mi(V2 a) { return S > 0 ? mk2(a.y, -a.x) : mk2(-a.y, a.x); }
// multiply by the constant (wr, -S wi): a root of unity of the forward (S = +1) or inverse transform
template <int S> __device__ __forceinline__ V2 mul_root(V2 a, RL wr, RL wi) {
    return S > 0 ? mk2(a.x * wr + a.y * wi, a.y * wr - a.x * wi) : mk2(a.x * wr - a.y * wi, a.y * wr + a.x * wi);
}

template <int S> __device__ __forceinline__ void dft2p(V2 &x0, V2 &x1) {
    const V2 s = cadd(x0, x1), d = csub(x0, x1);
    x0 = s; x1 = d;
}
template <int S> __device__ __forceinline__ void dft4(V2 &x0, V2 &x1, V2 &x2, V2 &x3) {
    const V2 s0 = cadd(x0, x2), s1 = csub(x0, x2), s2 = cadd(x1, x3), s3 = mul_mi<S>(csub(x1, x3));
    x0 = cadd(s0, s2); x2 = csub(s0, s2); x1 = cadd(s1, s3); x3 = csub(s1, s3);
}
template <int S> __device__ __forceinline__ void dft8(V2 (&v)[8]) {
    const RL h = (RL)0.70710678118654752440;
    V2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    V2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
    // b_r *= w8^r, w8 = exp(-+ i pi / 4)
    b1 = S > 0 ? mk2(h * (b1.x + b1.y), h * (b1.y - b1.x)) : mk2(h * (b1.x - b1.y), h * (b1.y + b1.x));
    b2 = mul_mi<S>(b2);
    b3 = S > 0 ? mk2(h * (b3.y - b3.x), -h * (b3.x + b3.y)) : mk2(-h * (b3.x + b3.y), h * (b3.x - b3.y));
    dft4<S>(a0, a1, a2, a3);
    dft4<S>(b0, b1, b2, b3);
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}
// 16-point DFT, natural order in and out, as 4 x 4: n = 4 n1 + n2, k = k1 + 4 k2,
//   X[k1 + 4 k2] = sum_n2 w4^(n2 k2) [ w16^(n2 k1) sum_n1 x[4 n1 + n2] w4^(n1 k1) ]
// IN8: the inputs v[8..15] are zero (a zero-padded row: n_in <= L / 2) and are not read; OUT8: only the outputs v[0..7] are
// wanted (n_out <= L / 2), v[8..15] are left undefined.  Both save the additions that would handle zeros / unused sums.
template <int S, bool IN8 = false, bool OUT8 = false> __device__ __forceinline__ void dft16(V2 (&v)[16]) {
    const RL h = (RL)0.70710678118654752440, c1 = (RL)0.92387953251128675613, s1 = (RL)0.38268343236508977173;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {                                                        // -> y[n2][k1] at v[4 k1 + n2]
        if constexpr (IN8) {
            const V2 x0 = v[n2], x1 = v[4 + n2], m = mul_mi<S>(x1);
            v[n2] = cadd(x0, x1); v[4 + n2] = cadd(x0, m); v[8 + n2] = csub(x0, x1); v[12 + n2] = csub(x0, m);
        } else {
            dft4<S>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
        }
    }
    v[5]  = mul_root<S>(v[5], c1, s1);        // k1 = 1: w16^1, w16^2, w16^3
    v[6]  = mul_root<S>(v[6], h, h);
    v[7]  = mul_root<S>(v[7], s1, c1);
    v[9]  = mul_root<S>(v[9], h, h);          // k1 = 2: w16^2, w16^4, w16^6
    v[10] = mul_mi<S>(v[10]);
    v[11] = mul_root<S>(v[11], -h, h);
    v[13] = mul_root<S>(v[13], s1, c1);       // k1 = 3: w16^3, w16^6, w16^9
    v[14] = mul_root<S>(v[14], -h, h);
    v[15] = mul_root<S>(v[15], -c1, -s1);
    if constexpr (OUT8) {                     // X[k1] (k2 = 0) and X[k1 + 4] (k2 = 1) only
        V2 lo[4], hi[4];
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            const V2 a0 = v[4 * k1], a1 = v[4 * k1 + 1], a2 = v[4 * k1 + 2], a3 = v[4 * k1 + 3];
            lo[k1] = cadd(cadd(a0, a2), cadd(a1, a3));
            hi[k1] = cadd(csub(a0, a2), mul_mi<S>(csub(a1, a3)));
        }
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) { v[k1] = lo[k1]; v[4 + k1] = hi[k1]; }
    } else {
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) dft4<S>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // X[k1 + 4 k2] at v[4 k1 + k2]
        // 4 x 4 transpose of the register names
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a + 1; b < 4; ++b) { const V2 x = v[4 * a + b]; v[4 * a + b] = v[4 * b + a]; v[4 * b + a] = x; }
    }
}
template <int R, int S> __device__ __forceinline__ void dftR(V2 (&v)[R]) {
    if constexpr (R == 16) dft16<S>(v);
    else if constexpr (R == 8) dft8<S>(v);
    else if constexpr (R == 4) dft4<S>(v[0], v[1], v[2], v[3]);
    else dft2p<S>(v[0], v[1]);
}

// v[r] *= w^r (CONJ: conj(w)^r), r = 1 .. R - 1: four interleaved chains of powers, each stepping by w^4
template <int R, bool CONJ> __device__ __forceinline__ void twiddle_powers(V2 (&v)[R], V2 w1) {
    if (CONJ) w1.y = -w1.y;
    v[1] = cmul(v[1], w1);
    if constexpr (R >= 4) {
        const V2 w2 = cmul(w1, w1), w3 = cmul(w2, w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        if constexpr (R >= 8) {
            const V2 w4 = cmul(w2, w2);
            V2 q0 = w4, q1 = w1, q2 = w2, q3 = w3;
#pragma unroll
            for (int a = 4; a < R; a += 4) {
                q1 = cmul(q1, w4); q2 = cmul(q2, w4); q3 = cmul(q3, w4);
                v[a] = cmul(v[a], q0); v[a + 1] = cmul(v[a + 1], q1); v[a + 2] = cmul(v[a + 2], q2); v[a + 3] = cmul(v[a + 3], q3);
                if (a + 4 < R) q0 = cmul(q0, w4);
            }
        }
    }
}

// shared-memory slot of element i (per row; rows are interleaved with stride ROWS)
template <int ROWS> __device__ __forceinline__ int slot(int i) { return (i + (i >> 4)) * ROWS; }

// last forward pass, product with H and first adjoint pass of one butterfly of the turn, in registers and in place;
// hv[r] = H[j + r L/RT] for this butterfly (j = t + q L/16), fetched by the caller before the barrier
template <int LOG2L, int ROWS>
__device__ __forceinline__ void turn(V2 *X, const V2 *hv, int t, int q, V2 w1) {
    constexpr int L = 1 << LOG2L, T = L / 16, RT = turn_radix(LOG2L), NB = 16 / RT;
    V2 v[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) v[r] = X[slot<ROWS>(t + (q + r * NB) * T)];
    twiddle_powers<RT, false>(v, w1);
    dftR<RT, 1>(v);
#pragma unroll
    for (int r = 0; r < RT; ++r) v[r] = cmul(v[r], hv[r]);
    dftR<RT, -1>(v);
    twiddle_powers<RT, true>(v, w1);
#pragma unroll
    for (int r = 0; r < RT; ++r) X[slot<ROWS>(t + (q + r * NB) * T)] = v[r];
}

// One row (per thread: its 16 elements of one row) of the chirp-z convolution y = IFFT(FFT(x) * H).  `load16(v)` fills
// v[s] with x[t + s L/16] (global memory), y leaves through `store16(v)` (v[s] = y[t + s L/16]).  X / Y are this
// thread's row base pointers in the two buffers (the same buffer when NBUF == 1); with two buffers the roles alternate
// from row to row, so the next row's first scatter never meets this row's last reads.  Everything a pass needs from
// global memory — its twiddle, the 16 values of H for the turn, the first eight post-chirp factors for the last pass — is
// fetched BEFORE the barrier that precedes the pass: the data registers are dead there (the row lives in shared memory),
// and the load latency overlaps the barrier wait instead of the arithmetic.
// HALF: n_in <= L / 2 and n_out <= L / 2 (the usual case: L is the power of two above n_in + n_out): the loader fills only
// v[0..7], the first butterfly skips the zero half, the last one computes only the wanted half, the store writes v[0..7].
// `after_first_barrier()` runs once, right after the first barrier of the row (every thread has consumed its inputs by then):
// the kernel uses it to start the asynchronous copy of the NEXT unit's input rows into the staging buffer.
template <int LOG2L, int ROWS, int NBUF, bool HALF, class Load16, class Store16, class Hook>
__device__ __forceinline__ void czt_row(V2 *&X, V2 *&Y, int t, const Load16 &load16, const V2 *__restrict__ H,
                                        const V2 *__restrict__ post, int nout, const Store16 &store16,
                                        const Hook &after_first_barrier) {
    constexpr int L = 1 << LOG2L, T = L / 16, NREG = nreg(LOG2L), RT = turn_radix(LOG2L), NB = 16 / RT, NS = L / RT;
    const V2 *__restrict__ tw = tw_table(LOG2L);
    const std::integral_constant<bool, HALF> half_tag;
    V2 v[16];
    load16(v, half_tag);
    dft16<1, HALF, false>(v);
#pragma unroll
    for (int r = 0; r < 16; ++r) X[(17 * t + r) * ROWS] = v[r];          // = slot(16 t + r)
    int Ns = 16;
#pragma unroll
    for (int p = 1; p < NREG; ++p) {
        const int k = t & (Ns - 1), j0 = (t - k) * 16 + k;
        const V2 w1 = tw[tw_offset(p) + k];
        __syncthreads();
        if (p == 1) after_first_barrier();
#pragma unroll
        for (int s = 0; s < 16; ++s) v[s] = X[slot<ROWS>(t + s * T)];
        twiddle_powers<16, false>(v, w1);
        dft16<1>(v);
        if (NBUF == 1) __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) Y[slot<ROWS>(j0 + r * Ns)] = v[r];
        if (NBUF == 2) { V2 *x = X; X = Y; Y = x; }
        Ns *= 16;
    }
    {
        // the first HN of this thread's 16 values of H before the barrier, the others once the turn is under way
        constexpr int HN = LOG2L <= LFD_CZT_PRE_MAXLG ? LFD_CZT_HPRE : 0;
        V2 wt[NB], hv[16];
#pragma unroll
        for (int q = 0; q < NB; ++q) wt[q] = tw[tw_offset(NREG) + t + q * T];
#pragma unroll
        for (int e = 0; e < HN; ++e) hv[e] = H[t + (e / RT) * T + (e % RT) * NS];
        __syncthreads();
        if (NREG == 1) after_first_barrier();
#pragma unroll
        for (int q = 0; q < NB; ++q) {
#pragma unroll
            for (int r = 0; r < RT; ++r)
                if (q * RT + r >= HN) hv[q * RT + r] = H[t + q * T + r * NS];
            turn<LOG2L, ROWS>(X, hv + q * RT, t, q, wt[q]);
        }
    }
#pragma unroll
    for (int p = NREG - 1; p >= 1; --p) {
        Ns /= 16;
        const int k = t & (Ns - 1), j0 = (t - k) * 16 + k;
        const V2 w1 = tw[tw_offset(p) + k];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = X[slot<ROWS>(j0 + r * Ns)];
        dft16<-1>(v);
        twiddle_powers<16, true>(v, w1);
        if (NBUF == 1) __syncthreads();
#pragma unroll
        for (int s = 0; s < 16; ++s) Y[slot<ROWS>(t + s * T)] = v[s];
        if (NBUF == 2) { V2 *x = X; X = Y; Y = x; }
    }
    constexpr int PP = LOG2L <= LFD_CZT_PRE_MAXLG ? LFD_CZT_PPRE : 0;   // post-chirp factors fetched before the barrier (outputs t + s L/16, s < PP)
    V2 pv[PP > 0 ? PP : 1];
#pragma unroll
    for (int sI = 0; sI < PP; ++sI) pv[sI] = (t + sI * T < nout) ? post[t + sI * T] : mk2((RL)0, (RL)0);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = X[(17 * t + r) * ROWS];
    dft16<-1, false, HALF>(v);
#pragma unroll
    for (int sI = 0; sI < PP; ++sI) v[sI] = cmul(v[sI], pv[sI]);
#pragma unroll
    for (int sI = PP; sI < 8; ++sI) v[sI] = (t + sI * T < nout) ? cmul(v[sI], post[t + sI * T]) : v[sI];
    if (!HALF && 8 * T < nout) {                           // uniform; outputs beyond half the transform length are rare
#pragma unroll
        for (int sI = 8; sI < 16; ++sI) v[sI] = (t + sI * T < nout) ? cmul(v[sI], post[t + sI * T]) : v[sI];
    }
    store16(v, half_tag);
    if (NBUF == 1) __syncthreads();
    else { V2 *x = X; X = Y; Y = x; }
}

// ---- one stage: every row of every plane whose FFT length is L ----------------------------------------------
// STAGE_A: row i of f (n elements, or the fused phasor) -> Gt[:, i] (N outputs, transposed store)
// else   : row v of Gt (m elements)                     -> out[:, v] (M outputs; complex128 or |.|^2 float64)
// A CTA takes ROWS consecutive rows at a time (lane % ROWS = row).  Register budget: 128 per thread (512 threads per SM).
__host__ __device__ constexpr int min_ctas(int lg) { return cta_threads(lg) >= REG_THREADS ? 1 : REG_THREADS / cta_threads(lg); }

// Work units (ROWS rows of one plane) are numbered plane-major over the planes of THIS length; starts[p] = units in planes
// < p (count + 1 entries).  Units are dealt round-robin: at any moment the CTAs of the grid work on ADJACENT rows, so the
// 16-byte pieces they scatter into the transposed intermediate (stage A) or the output columns (stage B) complete their
// 128-byte lines in L2 within one unit time.  (A contiguous run of units per CTA keeps a plane's tables in L1 but leaves
// N x CTAs partially written lines in flight — 77 MB for 2048-point planes — which L2 evicts half filled: measured
// 1.6x slower at 2001^2 -> 2048^2, LFD_CZT_CONTIG=1.)
template <int LOG2L, bool STAGE_A>
__global__ void __launch_bounds__(cta_threads(LOG2L), min_ctas(LOG2L))
czt_stage_kernel(const Plane *__restrict__ descs, const int *__restrict__ starts, int count) {
    constexpr int L = 1 << LOG2L, T = L / 16, ROWS = rows_for(LOG2L), NBUF = nbuf_for(LOG2L), NT = T * ROWS;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    V2 *sm = reinterpret_cast<V2 *>(sm_raw);
    const int c = threadIdx.x % ROWS, t = threadIdx.x / ROWS;
    V2 *X = sm + c, *Y = sm + (NBUF - 1) * ROWS * (L + L / 16) + c;
    const int total = starts[count];
#if LFD_CZT_CONTIG
    const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x, stride = 1;
    const int w0 = (int)blockIdx.x * per, w1 = min(total, w0 + per);
#else
    const int stride = (int)gridDim.x, w0 = (int)blockIdx.x, w1 = total;
#endif
    if (w0 >= w1) return;
    __shared__ Plane sd;                 // the plane this CTA is working on
    // Input staging (STG): while a unit is being transformed, the copy engine brings the NEXT unit's input rows into shared
    // memory (cp.async.bulk, completion on an mbarrier), so the first pass of the next unit reads shared memory instead of
    // waiting for HBM / L2.  One staging buffer suffices: it is refilled right after the first barrier of a row (all reads
    // of it are done) and has the rest of the row (~10 us) to arrive.  Region of row c: STG_REGION bytes (one complex row, or
    // the amp row followed by the opd row); the copy starts at the 16-byte boundary below the row, s_sh holds the shifts.
    constexpr bool STG = LFD_CZT_STAGING && LOG2L >= LFD_CZT_STAGING_MINLG;
    constexpr int STG_REGION = (L / 2 + 4) * 16;
    unsigned char *stg = sm_raw + (size_t)NBUF * ROWS * (L + L / 16) * sizeof(V2);
    __shared__ Plane sdn;                // the plane after it (the successor of a unit usually lies there)
    __shared__ __align__(8) uint64_t stg_bar;
    __shared__ int s_staged, s_sh[2 * ROWS];
    uint32_t stg_phase = 0;
    int pend2 = 0;                       // first unit after the NEXT plane
    if (STG) {
        if (threadIdx.x == 0) { mbar_init(&stg_bar, 1); mbar_fence_init(); s_staged = 0; }
        __syncthreads();
    }
    int plane = 0, cur = -1;
    {   // last plane whose first unit is <= w0 (planes of another length own no units: their start equals the next one's)
        int lo = 0, hi = count;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (starts[mid] <= w0) lo = mid; else hi = mid; }
        plane = lo;
    }
    int pbeg = starts[plane], pend = starts[plane + 1];
    for (int w = w0; w < w1; w += stride) {
        while (w >= pend) { ++plane; pbeg = pend; pend = starts[plane + 1]; }
        const int row0 = (w - pbeg) * ROWS;
        if (plane != cur) {              // uniform over the CTA
            __syncthreads();
            const unsigned long long *g = reinterpret_cast<const unsigned long long *>(descs + plane);
            unsigned long long *sdw = reinterpret_cast<unsigned long long *>(&sd);
            for (int i = threadIdx.x; i < (int)(sizeof(Plane) / 8); i += NT) sdw[i] = g[i];
            if (STG && plane + 1 < count) {
                const unsigned long long *gn = reinterpret_cast<const unsigned long long *>(descs + plane + 1);
                unsigned long long *snw = reinterpret_cast<unsigned long long *>(&sdn);
                for (int i = threadIdx.x; i < (int)(sizeof(Plane) / 8); i += NT) snw[i] = gn[i];
            }
            __syncthreads();
            cur = plane;
            pend2 = (STG && plane + 1 < count) ? starts[plane + 2] : pend;
        }
        const Plane &d = sd;
        const int nrows = STAGE_A ? d.m : d.N;
        const int row = row0 + c;
        const bool rv = row < nrows;
        const int nin = STAGE_A ? d.n : d.m, nout = STAGE_A ? d.N : d.M;
        const V2 *__restrict__ pre = (const V2 *)(STAGE_A ? d.preA : d.preB);
        const V2 *__restrict__ post = (const V2 *)(STAGE_A ? d.postA : d.postB);
        const V2 *__restrict__ H = (const V2 *)(STAGE_A ? d.HA : d.HB);
        const Plane *dp = &d;
        // was this unit's input staged by the previous iteration?  (uniform: written by thread 0 before the last barrier)
        const bool staged = STG && s_staged != 0;
        int sh0 = 0, sh1 = 0;
        if (staged) {
            sh0 = s_sh[2 * c]; sh1 = s_sh[2 * c + 1];
            mbar_wait(&stg_bar, stg_phase);
            stg_phase ^= 1;
        }
        const unsigned char *stg_row = stg + c * STG_REGION;
        // where the next unit of this CTA lies (thread 0 issues its copies from the hook below)
        const int wn = w + stride;
        const bool next_here = wn < w1 && wn < pend, next_there = wn < w1 && wn >= pend && wn < pend2;
        const Plane *np_ = next_here ? &sd : &sdn;
        const int nrow0 = next_here ? (wn - pbeg) * ROWS : (wn - pend) * ROWS;
        auto hook = [=]() {
            if (!STG || threadIdx.x != 0) return;
            int ok = 0;
            const int nnin = STAGE_A ? np_->n : np_->m, nnrows = STAGE_A ? np_->m : np_->N;
            const int nnout = STAGE_A ? np_->N : np_->M;
            if ((next_here || next_there) && (STAGE_A ? np_->logLA : np_->logLB) == LOG2L && nnin <= L / 2 && nnout <= L / 2) {
                const bool fusedn = STAGE_A && np_->amp != nullptr;
                const int es = fusedn ? 8 : (int)sizeof(V2);
                uint32_t total = 0;
                bool safe = true;
                const unsigned char *src[2 * ROWS];
                uint32_t nb[2 * ROWS];
#pragma unroll
                for (int cc = 0; cc < ROWS; ++cc) {
                    const long long r = nrow0 + cc;
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        src[2 * cc + a] = nullptr; nb[2 * cc + a] = 0;
                        if (r >= nnrows || (a == 1 && !fusedn)) continue;
                        const unsigned char *b0 = fusedn ? (const unsigned char *)(a ? np_->opd : np_->amp) + ((r + np_->pr0) * np_->pld + np_->pc0) * 8
                                                : (STAGE_A ? (const unsigned char *)np_->f + r * np_->ldf * es
                                                           : (const unsigned char *)np_->Gt + r * np_->mpad * es);
                        const uint32_t sh = (uint32_t)((uintptr_t)b0 & 15);
                        const uint32_t bytes = (sh + (uint32_t)nnin * es + 15u) & ~15u;
                        // never read past the end of the plane's last row: the tail beyond a row is the next row (or padding)
                        if (r == nnrows - 1 && ((sh + (uint32_t)nnin * es) & 15u) != 0 && STAGE_A) safe = false;
                        src[2 * cc + a] = b0 - sh; nb[2 * cc + a] = bytes;
                        s_sh[2 * cc + a] = (int)sh;
                        total += bytes;
                    }
                }
                if (safe && total > 0) {
                    mbar_expect_tx(&stg_bar, total);
#pragma unroll
                    for (int cc = 0; cc < ROWS; ++cc)
#pragma unroll
                        for (int a = 0; a < 2; ++a)
                            if (nb[2 * cc + a]) bulk_g2s(stg + cc * STG_REGION + a * (STG_REGION / 2), src[2 * cc + a], nb[2 * cc + a], &stg_bar);
                    ok = 1;
                }
            }
            s_staged = ok;
        };
        if (!STG && LFD_CZT_L2PRE && LOG2L <= LFD_CZT_PRE_MAXLG && w + stride < w1 && w + stride < pend && row + stride * ROWS < nrows) {   // next unit in the same plane: pull its input rows into L2 now
            const long long nrow = row + stride * ROWS;
#pragma unroll
            for (int sI = 0; sI < 16; sI += 2) {                         // one prefetch per 32-byte sector pair of this thread's elements
                const int i = t + sI * T;
                if (i < nin) {
                    if (STAGE_A && dp->amp != nullptr) {
                        const long long e = (nrow + dp->pr0) * dp->pld + dp->pc0 + i;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(dp->amp + e));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(dp->opd + e));
                    } else {
                        const V2 *nsrc = STAGE_A ? (const V2 *)dp->f + nrow * dp->ldf : (const V2 *)dp->Gt + nrow * dp->mpad;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + i));
                    }
                }
            }
        }
        // first forward pass: the inputs of this thread straight from global memory (x pre-chirp; zero beyond the input
        // length), in halves of eight so that every load of a half is issued before its arithmetic.  Threads whose eight
        // elements are all inside the row (every thread but those at the row's end) take the unpredicated path.
        auto load = [=](V2 (&v)[16], auto half_tag) {
            constexpr bool HALF = decltype(half_tag)::value;
#pragma unroll
            for (int hf = 0; hf < (HALF ? 1 : 2); ++hf) {
                if (!HALF && hf * 8 * T >= nin) {                        // uniform: this half lies beyond the input
#pragma unroll
                    for (int s = 0; s < 8; ++s) v[hf * 8 + s] = mk2((RL)0, (RL)0);
                    continue;
                }
                const bool all_in = rv && t + (hf * 8 + 7) * T < nin;
                V2 pr[8];
                if (STAGE_A && dp->amp != nullptr) {
                    double am[8], op[8];
                    const long long base = (long long)(dp->pr0 + row) * dp->pld + dp->pc0;
                    const unsigned char *mk = dp->mask;
                    // the row's amplitude and OPD: staged in shared memory by the previous unit, or straight from global memory
                    const double *sa = staged ? (const double *)(stg_row + sh0) : dp->amp + base;
                    const double *so = staged ? (const double *)(stg_row + STG_REGION / 2 + sh1) : dp->opd + base;
                    if (all_in) {
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const int i = t + (hf * 8 + s) * T;
                            am[s] = sa[i];
                            op[s] = so[i];
                            pr[s] = pre[i];
                        }
                        if (mk != nullptr) {
#pragma unroll
                            for (int s = 0; s < 8; ++s)
                                if (mk[base + t + (hf * 8 + s) * T] == 0) am[s] = 0.0;
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const int i = t + (hf * 8 + s) * T;
                            const bool in = rv && i < nin;
                            am[s] = in ? sa[i] : 0.0;
                            op[s] = in ? so[i] : 0.0;
                            pr[s] = in ? pre[i] : mk2((RL)0, (RL)0);
                            if (in && mk != nullptr && mk[base + i] == 0) am[s] = 0.0;
                        }
                    }
                    const double lam = dp->wavelength, inv_lam = dp->inv_wavelength;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        V2 x = mk2((RL)0, (RL)0);
                        if (am[s] != 0.0) x = phasor(am[s], op[s], lam, inv_lam);   // same arithmetic as K1 (pupil_prep.cu)
                        v[hf * 8 + s] = cmul(x, pr[s]);
                    }
                } else {
                    const V2 *src = staged ? (const V2 *)(stg_row + sh0)
                                           : (STAGE_A ? (const V2 *)dp->f + (long long)row * dp->ldf : (const V2 *)dp->Gt + (long long)row * dp->mpad);
                    V2 x[8];
                    if (all_in) {
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const int i = t + (hf * 8 + s) * T;
                            x[s] = src[i];
                            pr[s] = pre[i];
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            const int i = t + (hf * 8 + s) * T;
                            const bool in = rv && i < nin;
                            x[s] = in ? src[i] : mk2((RL)0, (RL)0);
                            pr[s] = in ? pre[i] : mk2((RL)0, (RL)0);
                        }
                    }
#pragma unroll
                    for (int s = 0; s < 8; ++s) v[hf * 8 + s] = cmul(x[s], pr[s]);
                }
            }
        };
        // the usual case: both the input and the wanted output fit half the transform length (uniform over the plane)
        const bool half = nin <= 8 * T && nout <= 8 * T;
        const bool all_out = rv && t + 7 * T < nout;
        if (STAGE_A) {
            V2 *Gt = (V2 *)dp->Gt + row; const long long mpad = dp->mpad;
            auto store = [=](V2 (&v)[16], auto half_tag) {
                constexpr int NS_ = decltype(half_tag)::value ? 8 : 16;
                if (decltype(half_tag)::value && all_out) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) Gt[(long long)(t + s * T) * mpad] = v[s];
                    return;
                }
#pragma unroll
                for (int s = 0; s < NS_; ++s) {
                    const int i = t + s * T;
                    if (rv && i < nout) Gt[(long long)i * mpad] = v[s];
                }
            };
            if (half) czt_row<LOG2L, ROWS, NBUF, true>(X, Y, t, load, H, post, nout, store, hook);
            else czt_row<LOG2L, ROWS, NBUF, false>(X, Y, t, load, H, post, nout, store, hook);
        } else if (dp->intensity) {
            double *out = (double *)dp->out + row; const long long ldo = dp->ldo;
            auto store = [=](V2 (&v)[16], auto half_tag) {
                constexpr int NS_ = decltype(half_tag)::value ? 8 : 16;
                if (decltype(half_tag)::value && all_out) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) out[(long long)(t + s * T) * ldo] = (double)v[s].x * (double)v[s].x + (double)v[s].y * (double)v[s].y;
                    return;
                }
#pragma unroll
                for (int s = 0; s < NS_; ++s) {
                    const int i = t + s * T;
                    if (rv && i < nout) out[(long long)i * ldo] = (double)v[s].x * (double)v[s].x + (double)v[s].y * (double)v[s].y;
                }
            };
            if (half) czt_row<LOG2L, ROWS, NBUF, true>(X, Y, t, load, H, post, nout, store, hook);
            else czt_row<LOG2L, ROWS, NBUF, false>(X, Y, t, load, H, post, nout, store, hook);
        } else {
            V2 *out = (V2 *)dp->out + row; const long long ldo = dp->ldo;
            auto store = [=](V2 (&v)[16], auto half_tag) {
                constexpr int NS_ = decltype(half_tag)::value ? 8 : 16;
                if (decltype(half_tag)::value && all_out) {
#pragma unroll
                    for (int s = 0; s < 8; ++s) out[(long long)(t + s * T) * ldo] = v[s];
                    return;
                }
#pragma unroll
                for (int s = 0; s < NS_; ++s) {
                    const int i = t + s * T;
                    if (rv && i < nout) out[(long long)i * ldo] = v[s];
                }
            };
            if (half) czt_row<LOG2L, ROWS, NBUF, true>(X, Y, t, load, H, post, nout, store, hook);
            else czt_row<LOG2L, ROWS, NBUF, false>(X, Y, t, load, H, post, nout, store, hook);
        }
    }
}

