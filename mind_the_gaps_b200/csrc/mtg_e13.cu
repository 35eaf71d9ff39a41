// mtg_e13.cu -- K6: Emmanoulopoulos et al. (2013) PDF adjustment, batched on the device.
//
// Restates E13Simulator.adjust_lightcurve_pdf / adjust_pdf (mind_the_gaps/simulator.py:65-140):
//   A     = |rfft(countrate)| / (n//2 + 1)                                     (:96-98)
//   loop: phi = angle(rfft(xsim)); adj = irfft(A e^{i phi}); rank-match adj to the sorted draws
//         (xsim_adjust[argsort(-xsim_adjust)] = xsim[argsort(-xsim)])           (:101-105, :113-121)
//         until np.allclose(xsim_adjust, xsim, rtol=1e-4) or max_iter           (:111)
//   result = xsim (the LAST ACCEPTED iterate, not xsim_adjust)                  (:129)
// The multiset of values never changes, so the descending-sorted draws are computed once and every
// iteration only needs the descending argsort of adj (CUB radix sort) and a scatter.
// FFTs are cuFFT D2Z / Z2D (library calls, like pyfftw in the reference); the phase replacement
// A F/|F| avoids atan2 + sincos (equal to A e^{i angle(F)} to rounding; angle(0) = 0 -> A).
#include "mtg_sim.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_sort.cuh>

__device__ __forceinline__ uint4 e13_philox(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ double e13_u01(uint32_t a, uint32_t b)
{ return ((double)((((uint64_t)a << 32) | b) >> 11) + 0.5) * 1.1102230246251565e-16; }

// gather the segments of the lightcurves listed in `members` into a dense [cnt][n] batch, adding the mean
// start of the working window of lightcurve l: its segment start, moved back when a padded window would run off the series
__device__ __forceinline__ int e13_start(const SegInfo &s, int n_fft, int n) { return s.i0 + n <= n_fft ? s.i0 : n_fft - n; }

__global__ void k_e13_gather(const double *__restrict__ lc, int n_fft, const SegInfo *__restrict__ seg,
                             const int *__restrict__ members, int n, double mean, double *__restrict__ dst)
{
    const int m = blockIdx.y;
    const int l = members ? members[m] : m;
    const double *src = lc + (size_t)l * n_fft + e13_start(seg[l], n_fft, n);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[(size_t)m * n + i] = src[i] + mean;
}
__global__ void k_e13_scatter_back(double *__restrict__ lc, int n_fft, const SegInfo *__restrict__ seg,
                                   const int *__restrict__ members, int n, const double *__restrict__ src,
                                   const int32_t *__restrict__ iters, int32_t *__restrict__ iters_out)
{
    const int m = blockIdx.y;
    const int l = members ? members[m] : m;
    double *dst = lc + (size_t)l * n_fft + e13_start(seg[l], n_fft, n);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        dst[i] = src[(size_t)m * n + i];
    if (iters_out && blockIdx.x == 0 && threadIdx.x == 0) iters_out[l] = iters[m];
}

// per-lightcurve mean and population standard deviation (np.std), one block per lightcurve
__global__ void __launch_bounds__(256) k_e13_std(const double *__restrict__ x, int n, double *__restrict__ out_std)
{
    __shared__ double sh[256];
    const double *p = x + (size_t)blockIdx.x * n;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += p[i];
    sh[threadIdx.x] = s; __syncthreads();
    for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k]; __syncthreads(); }
    const double mean = sh[0] / n;
    __syncthreads();
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) { const double d = p[i] - mean; v += d * d; }
    sh[threadIdx.x] = v; __syncthreads();
    for (int k = 128; k > 0; k >>= 1) { if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k]; __syncthreads(); }
    if (threadIdx.x == 0) out_std[blockIdx.x] = sqrt(sh[0] / n);
}

// pdf.rvs(size=n) (simulator.py:90) from standard variates: lognormal scale*exp(sigma z)
// (stats.py:116-129), uniform loc + scale*u (stats.py:132-146)
__global__ void k_e13_draw(int pdf, double mean, const double *__restrict__ stds, int n, const double *__restrict__ variates,
                           size_t var_stride, const int *__restrict__ members, uint64_t seed, uint64_t lc_offset,
                           double *__restrict__ xsim)
{
    const int m = blockIdx.y;
    const double std_ = stds[m], var = std_ * std_;
    double p0, p1;
    if (pdf == MTG_PDF_LOGNORMAL) { p0 = sqrt(log(var / (mean * mean) + 1.0)); p1 = exp(log((mean * mean) / sqrt(var + mean * mean))); }
    else { const double b = sqrt(3.0 * var) + mean, a = 2.0 * mean - b; p0 = a; p1 = b - a; }
    const int l = members ? members[m] : m;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double v;
        if (variates) v = variates[(size_t)l * var_stride + i];
        else {
            const uint64_t item = (lc_offset + l) * (uint64_t)0x100000000ull + (uint64_t)i;
            const uint4 r = e13_philox(make_uint4((uint32_t)item, (uint32_t)(item >> 32), 0u, 2u),
                                       make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            const double u1 = e13_u01(r.x, r.y), u2 = e13_u01(r.z, r.w);
            if (pdf == MTG_PDF_LOGNORMAL) { double s, c; sincospi(2.0 * u2, &s, &c); v = sqrt(-2.0 * log(u1)) * c; }
            else v = u1;
        }
        xsim[(size_t)m * n + i] = (pdf == MTG_PDF_LOGNORMAL) ? p1 * exp(p0 * v) : p0 + p1 * v;
    }
}

__global__ void k_e13_amplitude(const double2 *__restrict__ F, size_t count, double inv, double *__restrict__ A)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        A[i] = hypot(F[i].x, F[i].y) * inv;
}
// F <- A F/|F| / n   (the 1/n of numpy's irfft folded in)
__global__ void k_e13_phasor(double2 *__restrict__ F, const double *__restrict__ A, size_t count, double inv_n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const double2 f = F[i];
        // A F/|F| with one reciprocal square root (1 ulp) instead of hypot and two divisions; |F|^2 of a sum of <= 2^22 flux
        // values neither overflows nor underflows unless F is exactly 0 (angle(0) = 0)
        const double r2 = fma(f.x, f.x, f.y * f.y), a = A[i] * inv_n;
        const double sc = a * rsqrt(r2);
        F[i] = (r2 > 0.0 && r2 < 1.0e300) ? make_double2(f.x * sc, f.y * sc)
                                          : (r2 > 0.0 ? make_double2(a * (f.x / hypot(f.x, f.y)), a * (f.y / hypot(f.x, f.y))) : make_double2(a, 0.0));
    }
}
__global__ void k_e13_iota(int32_t *__restrict__ v, int n, int cnt)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < (size_t)n * cnt; i += (size_t)gridDim.x * blockDim.x)
        v[i] = (int32_t)(i % n);
}
// new[perm[r]] = sorted_desc[r]
__global__ void k_e13_rank_match(const int32_t *__restrict__ perm, const double *__restrict__ sorted_desc, int n,
                                 double *__restrict__ out)
{
    const int m = blockIdx.y;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
        out[(size_t)m * n + perm[(size_t)m * n + r]] = sorted_desc[(size_t)m * n + r];
}
// ---- rank order from 32-bit keys ----------------------------------------------------------------------------------
// The descending argsort of 1.3e6 doubles per lightcurve and iteration was half of K6's time as a 64-bit-key radix sort
// (8 one-sweep passes over 12-byte pairs).  Rounding to float is monotone, so sorting (float(adj), index) pairs -- 4 passes
// over 8-byte pairs -- orders everything except runs of EQUAL float keys; k_e13_fix_ties then orders every such run by the
// double values (insertion sort; runs are 1-3 long for continuous data).  The result is exactly the descending order of the
// doubles, ties by ascending index.
__global__ void k_e13_keys32(const double *__restrict__ x, size_t count, float *__restrict__ k)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) k[i] = (float)x[i];
}
__global__ void k_e13_fix_ties(const float *__restrict__ k, const double *__restrict__ x, int n, int32_t *__restrict__ perm)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const float kr = k[r];
        if (r > 0 && k[r - 1] == kr) continue;               // inside a run: its first element does the work
        int e = r + 1;
        while (e < n && k[e] == kr) ++e;
        for (int i = r + 1; i < e; ++i) {
            const int32_t pi = perm[i];
            const double vi = x[pi];
            int j = i - 1;
            while (j >= r) {
                const int32_t pj = perm[j];
                const double vj = x[pj];
                if (vj > vi || (vj == vi && pj < pi)) break;
                perm[j + 1] = pj; --j;
            }
            perm[j + 1] = pi;
        }
    }
}

// np.allclose(a, b, rtol=1e-4, atol=1e-8): notclose[m] = any(|a-b| > atol + rtol |b|)
__global__ void k_e13_notclose(const double *__restrict__ a, const double *__restrict__ b, int n,
                               const int32_t *__restrict__ done, int32_t *__restrict__ notclose)
{
    const int m = blockIdx.y;
    if (done[m]) return;
    int bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double x = a[(size_t)m * n + i], y = b[(size_t)m * n + i];
        if (!(fabs(x - y) <= 1e-8 + 1e-4 * fabs(y))) bad = 1;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(&notclose[m], 1);
}
// accept / stop decision per lightcurve (simulator.py:111-125)
__global__ void k_e13_decide(int cnt, int max_iter, int32_t *__restrict__ notclose, int32_t *__restrict__ done,
                             int32_t *__restrict__ iters, int32_t *__restrict__ n_active)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= cnt) return;
    if (!done[m]) {
        if (!notclose[m] || iters[m] >= max_iter) done[m] = 1;
        else { iters[m] += 1; atomicAdd(n_active, 1); }
    }
    notclose[m] = 0;
}
// xsim <- new for lightcurves still iterating (done was updated by k_e13_decide BEFORE this copy)
__global__ void k_e13_accept(const double *__restrict__ src, double *__restrict__ xsim, int n, const int32_t *__restrict__ done)
{
    const int m = blockIdx.y;
    if (done[m]) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        xsim[(size_t)m * n + i] = src[(size_t)m * n + i];
}

struct E13Work {
    double *seg = nullptr, *xsim = nullptr, *adj = nullptr, *sorted = nullptr, *keys = nullptr, *A = nullptr, *stds = nullptr;
    double2 *F = nullptr; int32_t *iota = nullptr, *perm = nullptr, *done = nullptr, *notclose = nullptr, *iters = nullptr, *n_active = nullptr;
    int *offsets = nullptr; void *cub_tmp = nullptr; size_t cub_bytes = 0;
    float *key32 = nullptr, *key32_out = nullptr;
    int cap_cnt = 0, cap_n = 0, offsets_n = 0;
    void release()
    {
        cudaFree(key32); cudaFree(key32_out);
        cudaFree(seg); cudaFree(xsim); cudaFree(adj); cudaFree(sorted); cudaFree(keys); cudaFree(A); cudaFree(stds); cudaFree(F);
        cudaFree(iota); cudaFree(perm); cudaFree(done); cudaFree(notclose); cudaFree(iters); cudaFree(n_active); cudaFree(offsets);
        cudaFree(cub_tmp);
    }
};

static int e13_alloc(void **p, size_t bytes)
{
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 8);
    if (e != cudaSuccess) { mtg_set_error("E13 cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return MTG_ENOMEM; }
    return 0;
}

// descending sort of every length-n row of keys_in; optional payload
static int e13_sort_desc(E13Work &w, int cnt, int n, const double *keys_in, double *keys_out, const int32_t *vals_in,
                         int32_t *vals_out, cudaStream_t st)
{
    if (n >= 32768) {
        if (vals_in) {
            const unsigned g = (unsigned)std::min<long long>(((long long)cnt * n + 255) / 256, 148 * 32);
            k_e13_keys32<<<g, 256, 0, st>>>(keys_in, (size_t)cnt * n, w.key32);
            for (int m = 0; m < cnt; ++m) {
                size_t bytes = w.cub_bytes;
                cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(w.cub_tmp, bytes, w.key32 + (size_t)m * n, w.key32_out + (size_t)m * n,
                                                                          vals_in + (size_t)m * n, vals_out + (size_t)m * n, n, 0, 32, st);
                if (e != cudaSuccess) return mtg_cuda_check(e, "cub radix sort (32-bit keys)");
                k_e13_fix_ties<<<(unsigned)std::min((n + 255) / 256, 148 * 8), 256, 0, st>>>(w.key32_out + (size_t)m * n, keys_in + (size_t)m * n, n,
                                                                                            vals_out + (size_t)m * n);
            }
            mtg_count_launch((uint64_t)cnt * 5 + 1);
            return mtg_cuda_check(cudaGetLastError(), "E13 rank order");
        }
        for (int m = 0; m < cnt; ++m) {
            size_t bytes = w.cub_bytes;
            cudaError_t e = cub::DeviceRadixSort::SortKeysDescending(w.cub_tmp, bytes, keys_in + (size_t)m * n, keys_out + (size_t)m * n, n, 0, 64, st);
            if (e != cudaSuccess) return mtg_cuda_check(e, "cub radix sort");
        }
        mtg_count_launch((uint64_t)cnt * 4);
        return 0;
    }
    size_t bytes = w.cub_bytes;
    cudaError_t e = vals_in
        ? cub::DeviceSegmentedSort::SortPairsDescending(w.cub_tmp, bytes, keys_in, keys_out, vals_in, vals_out, (long long)cnt * n, cnt,
                                                        w.offsets, w.offsets + 1, st)
        : cub::DeviceSegmentedSort::SortKeysDescending(w.cub_tmp, bytes, keys_in, keys_out, (long long)cnt * n, cnt, w.offsets, w.offsets + 1, st);
    mtg_count_launch(3);
    return mtg_cuda_check(e, "cub segmented sort");
}

// E13 on a dense batch: w.seg[cnt][n] holds the TK95 segment count rates (mean included) on entry.
// On exit w.xsim holds the adjusted lightcurves and w.iters the iteration counts.
static int e13_run_batch(Sim *s, std::map<long long, cufftHandle> &pz2d, std::map<long long, cufftHandle> &pd2z, E13Work &w,
                         int cnt, int n, int pdf, double mean, int max_iter, const double *d_variates, size_t var_stride,
                         const int *d_members, uint64_t seed, uint64_t lc_offset, cudaStream_t st,
                         bool draws_ready = false)
{
    const int Mh = n / 2 + 1;
    int rc;
    cufftHandle fwd, inv;
    if ((rc = sim_get_plan(s, pd2z, CUFFT_D2Z, n, cnt, &fwd))) return rc;
    if ((rc = sim_get_plan(s, pz2d, CUFFT_Z2D, n, cnt, &inv))) return rc;
    cufftSetStream(fwd, st); cufftSetStream(inv, st);
    const dim3 g2((unsigned)std::min<long long>((n + 255) / 256, 1024), cnt);
    const unsigned g1 = (unsigned)std::min<long long>(((long long)cnt * Mh + 255) / 256, 148 * 16);

    if (!draws_ready) {
        k_e13_std<<<cnt, 256, 0, st>>>(w.seg, n, w.stds);
        k_e13_draw<<<g2, 256, 0, st>>>(pdf, mean, w.stds, n, d_variates, var_stride, d_members, seed, lc_offset, w.xsim);
    }
    if (cufftExecD2Z(fwd, w.seg, (cufftDoubleComplex *)w.F) != CUFFT_SUCCESS) { mtg_set_error("cufftExecD2Z failed"); return MTG_EINVAL; }
    k_e13_amplitude<<<g1, 256, 0, st>>>(w.F, (size_t)cnt * Mh, 1.0 / (double)(n / 2 + 1), w.A);
    k_e13_iota<<<g1, 256, 0, st>>>(w.iota, n, cnt);
    if ((rc = e13_sort_desc(w, cnt, n, w.xsim, w.sorted, nullptr, nullptr, st))) return rc;
    MTG_CUDA(cudaMemsetAsync(w.done, 0, sizeof(int32_t) * cnt, st));
    MTG_CUDA(cudaMemsetAsync(w.notclose, 0, sizeof(int32_t) * cnt, st));
    MTG_CUDA(cudaMemsetAsync(w.iters, 0, sizeof(int32_t) * cnt, st));
    mtg_count_launch(4);

    int32_t *h_active = nullptr;
    MTG_CUDA(cudaMallocHost((void **)&h_active, sizeof(int32_t)));
    *h_active = cnt;
    rc = 0;
    for (int it = 0; it <= max_iter + 1 && *h_active > 0 && !rc; ++it) {
        // adj = irfft(A e^{i angle(rfft(xsim))})
        if (cufftExecD2Z(fwd, w.xsim, (cufftDoubleComplex *)w.F) != CUFFT_SUCCESS) { mtg_set_error("cufftExecD2Z failed"); rc = MTG_EINVAL; break; }
        k_e13_phasor<<<g1, 256, 0, st>>>(w.F, w.A, (size_t)cnt * Mh, 1.0 / (double)n);
        if (cufftExecZ2D(inv, (cufftDoubleComplex *)w.F, w.adj) != CUFFT_SUCCESS) { mtg_set_error("cufftExecZ2D failed"); rc = MTG_EINVAL; break; }
        // rank matching
        if ((rc = e13_sort_desc(w, cnt, n, w.adj, w.keys, w.iota, w.perm, st))) break;
        k_e13_rank_match<<<g2, 256, 0, st>>>(w.perm, w.sorted, n, w.adj);
        // convergence / acceptance
        k_e13_notclose<<<g2, 256, 0, st>>>(w.adj, w.xsim, n, w.done, w.notclose);
        cudaMemsetAsync(w.n_active, 0, sizeof(int32_t), st);
        k_e13_decide<<<(cnt + 127) / 128, 128, 0, st>>>(cnt, max_iter, w.notclose, w.done, w.iters, w.n_active);
        k_e13_accept<<<g2, 256, 0, st>>>(w.adj, w.xsim, n, w.done);
        mtg_count_launch(5);
        if ((it & 3) == 3 || it < 2) {
            cudaMemcpyAsync(h_active, w.n_active, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
            rc = mtg_cuda_check(cudaStreamSynchronize(st), "E13 sync");
        }
    }
    cudaFreeHost(h_active);
    return rc;
}

// workspace for batches of up to `cnt` series of up to `n` points; `use_cnt` x `use_n` is the batch about to run (the rows
// of a batch are packed with stride use_n).  A cached workspace (one per simulator) is reused while it is large enough.
static int e13_work_alloc(E13Work &w, int cnt, int n, int use_cnt = 0, int use_n = 0)
{
    if (use_n <= 0) { use_cnt = cnt; use_n = n; }
    const bool big = use_n >= 32768;
    if (use_cnt <= w.cap_cnt && use_n <= w.cap_n && (w.cap_n >= 32768) == big) {
        if (!big && w.offsets_n != use_n) {
            std::vector<int> off(w.cap_cnt + 1);
            for (int m = 0; m <= w.cap_cnt; ++m) off[m] = m * use_n;
            MTG_CUDA(cudaMemcpy(w.offsets, off.data(), sizeof(int) * (w.cap_cnt + 1), cudaMemcpyHostToDevice));
            w.offsets_n = use_n;
        }
        return 0;
    }
    w.release();
    w = E13Work();
    if ((n >= 32768) != big) n = use_n;
    if (cnt < use_cnt) cnt = use_cnt;
    if (n < use_n) n = use_n;
    const size_t cn = (size_t)cnt * n, cm = (size_t)cnt * (n / 2 + 1);
    int rc = 0;
    if ((rc = e13_alloc((void **)&w.seg, cn * 8)) || (rc = e13_alloc((void **)&w.xsim, cn * 8)) || (rc = e13_alloc((void **)&w.adj, cn * 8)) ||
        (rc = e13_alloc((void **)&w.sorted, cn * 8)) || (rc = e13_alloc((void **)&w.keys, cn * 8)) || (rc = e13_alloc((void **)&w.A, cm * 8)) ||
        (rc = e13_alloc((void **)&w.F, cm * 16)) || (rc = e13_alloc((void **)&w.iota, cn * 4)) || (rc = e13_alloc((void **)&w.perm, cn * 4)) ||
        (rc = e13_alloc((void **)&w.key32, big ? cn * 4 : 8)) || (rc = e13_alloc((void **)&w.key32_out, big ? cn * 4 : 8)) ||
        (rc = e13_alloc((void **)&w.stds, (size_t)cnt * 8)) || (rc = e13_alloc((void **)&w.done, (size_t)cnt * 4)) ||
        (rc = e13_alloc((void **)&w.notclose, (size_t)cnt * 4)) || (rc = e13_alloc((void **)&w.iters, (size_t)cnt * 4)) ||
        (rc = e13_alloc((void **)&w.n_active, 4)) || (rc = e13_alloc((void **)&w.offsets, (size_t)(cnt + 1) * sizeof(int)))) return rc;
    std::vector<int> off(cnt + 1);
    for (int m = 0; m <= cnt; ++m) off[m] = m * use_n;
    MTG_CUDA(cudaMemcpy(w.offsets, off.data(), sizeof(int) * (cnt + 1), cudaMemcpyHostToDevice));
    size_t b1 = 0, b2 = 0;
    if (big) {
        cub::DeviceRadixSort::SortPairsDescending(nullptr, b1, (float *)nullptr, (float *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, n);
        cub::DeviceRadixSort::SortKeysDescending(nullptr, b2, (double *)nullptr, (double *)nullptr, n);
    } else {
        cub::DeviceSegmentedSort::SortPairsDescending(nullptr, b1, (double *)nullptr, (double *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr,
                                                      (long long)cn, cnt, w.offsets, w.offsets + 1);
        cub::DeviceSegmentedSort::SortKeysDescending(nullptr, b2, (double *)nullptr, (double *)nullptr, (long long)cn, cnt, w.offsets, w.offsets + 1);
    }
    w.cub_bytes = b1 > b2 ? b1 : b2;
    if ((rc = e13_alloc(&w.cub_tmp, w.cub_bytes))) return rc;
    w.cap_cnt = cnt; w.cap_n = n; w.offsets_n = use_n;
    return 0;
}

// ---- long series (n >= 32768): one lightcurve per slot ----------------------------------------------------------------
// The batched loop above carries EVERY lightcurve of a batch through the FFT pair and the sort until the last one has
// converged (max over the batch ~ 25 iterations where the mean is 19), and a batch of 1.3e6-point series streams every
// array through HBM.  Long series therefore run one at a time: a slot (own stream, cuFFT plans and ~60 n bytes of work
// arrays) carries ONE lightcurve to convergence, the host reads its allclose flag after every iteration and hands the slot
// the next lightcurve as soon as it is done; MTG_E13_SLOTS (default 4) slots keep the device busy while the host decides
// (measured 188 / 288-317 / 351-366 / 383-390 lightcurves/s with 1 / 2 / 3 / 4).  Accepting an iterate is a pointer swap, the rank matching and the
// allclose test are one kernel, and the series moments come from a grid-wide reduction instead of one CTA per series.
#define E13_NB 296
struct E13Slot {
    cudaStream_t st = nullptr; cudaEvent_t ev = nullptr;
    double *x[2] = {nullptr, nullptr};                 // current iterate / candidate (swapped on acceptance)
    double *A = nullptr, *sorted = nullptr, *part = nullptr, *std_ = nullptr; double2 *F = nullptr;
    float *k32 = nullptr, *k32o = nullptr; int32_t *perm = nullptr, *flag = nullptr, *h_flag = nullptr;
    void *tmp = nullptr; size_t tmp_bytes = 0;
    std::map<int, cufftHandle> fwd, inv;
    int cap = 0;
    int lc = -1, it = 0, n = 0, cur = 0;               // state: lightcurve in flight, iteration, length, index of the current iterate
    double *dst = nullptr;
    std::map<long long, cudaGraphExec_t> graphs;       // (n << 1 | cur) -> one captured iteration (tied to this slot's buffers)
    bool graphs_ok = true;
    void drop_graphs()
    {
        for (auto &kv : graphs) cudaGraphExecDestroy(kv.second);
        graphs.clear();
    }
    void free_buffers()
    {
        drop_graphs();
        cudaFree(x[0]); cudaFree(x[1]); cudaFree(A); cudaFree(sorted); cudaFree(part); cudaFree(std_); cudaFree(F);
        cudaFree(k32); cudaFree(k32o); cudaFree(perm); cudaFree(flag); cudaFree(tmp);
        x[0] = x[1] = A = sorted = part = std_ = nullptr; F = nullptr; k32 = k32o = nullptr; perm = flag = nullptr; tmp = nullptr; cap = 0;
    }
    void release()
    {
        free_buffers();
        for (auto &kv : fwd) cufftDestroy(kv.second);
        for (auto &kv : inv) cufftDestroy(kv.second);
        fwd.clear(); inv.clear();
        if (h_flag) cudaFreeHost(h_flag);
        if (ev) cudaEventDestroy(ev);
        if (st) cudaStreamDestroy(st);
        h_flag = nullptr; ev = nullptr; st = nullptr;
    }
};
struct E13Large {
    std::vector<E13Slot> slots; int32_t *iota = nullptr; int iota_n = 0; cudaEvent_t ev_in = nullptr;
    bool allow_graphs = false;                         // set by the caller: all series of the run have one length
    int32_t *h_iters = nullptr; int h_iters_n = 0;
    void release()
    {
        for (auto &q : slots) q.release();
        slots.clear();
        cudaFree(iota); iota = nullptr; iota_n = 0;
        if (h_iters) cudaFreeHost(h_iters);
        h_iters = nullptr; h_iters_n = 0;
        if (ev_in) cudaEventDestroy(ev_in);
        ev_in = nullptr;
    }
};

__global__ void k_e13_addmean(const double *__restrict__ src, int n, double mean, double *__restrict__ dst)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i] + mean;
}
// partial sums of d = x - x[0] and d^2 per block; k_e13_moments_fin turns them into the population standard deviation
__global__ void __launch_bounds__(256) k_e13_moments(const double *__restrict__ x, int n, double *__restrict__ part)
{
    __shared__ double sh[2][8];
    const double shift = x[0];
    double s1 = 0.0, s2 = 0.0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) { const double d = x[i] - shift; s1 += d; s2 = fma(d, d, s2); }
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); s2 += __shfl_down_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s1; sh[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < 8; ++w) { a += sh[0][w]; b += sh[1][w]; }
        part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
    }
}
__global__ void k_e13_moments_fin(const double *__restrict__ part, int nb, int n, double *__restrict__ out_std)
{
    if (threadIdx.x || blockIdx.x) return;
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nb; ++i) { a += part[2 * i]; b += part[2 * i + 1]; }
    const double m = a / n, v = b / n - m * m;
    out_std[0] = sqrt(v > 0.0 ? v : 0.0);
}
// out[perm[r]] = sorted_desc[r] and the np.allclose(out, xold, rtol=1e-4, atol=1e-8) test of the same elements
__global__ void __launch_bounds__(256) k_e13_match_check(const int32_t *__restrict__ perm, const double *__restrict__ sorted_desc,
                                                         const double *__restrict__ xold, double *__restrict__ out, int n,
                                                         int32_t *__restrict__ flag)
{
    int bad = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int i = perm[r];
        const double v = sorted_desc[r], y = xold[i];
        out[i] = v;
        if (!(fabs(v - y) <= 1e-8 + 1e-4 * fabs(y))) bad = 1;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) *flag = 1;
}

static int e13_slot_reserve(E13Slot &q, int n)
{
    if (!q.st) {
        MTG_CUDA(cudaStreamCreateWithFlags(&q.st, cudaStreamNonBlocking));
        MTG_CUDA(cudaEventCreateWithFlags(&q.ev, cudaEventDisableTiming));
        MTG_CUDA(cudaMallocHost((void **)&q.h_flag, sizeof(int32_t)));
    }
    if (n <= q.cap) return 0;
    q.free_buffers();
    const size_t N = (size_t)n, Mh = N / 2 + 1;
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, b1, (float *)nullptr, (float *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, n);
    cub::DeviceRadixSort::SortKeysDescending(nullptr, b2, (double *)nullptr, (double *)nullptr, n);
    q.tmp_bytes = b1 > b2 ? b1 : b2;
    int rc = 0;
    if ((rc = e13_alloc((void **)&q.x[0], N * 8)) || (rc = e13_alloc((void **)&q.x[1], N * 8)) || (rc = e13_alloc((void **)&q.A, Mh * 8)) ||
        (rc = e13_alloc((void **)&q.sorted, N * 8)) || (rc = e13_alloc((void **)&q.F, Mh * 16)) || (rc = e13_alloc((void **)&q.k32, N * 4)) ||
        (rc = e13_alloc((void **)&q.k32o, N * 4)) || (rc = e13_alloc((void **)&q.perm, N * 4)) || (rc = e13_alloc((void **)&q.flag, 4)) ||
        (rc = e13_alloc((void **)&q.part, sizeof(double) * 2 * E13_NB)) || (rc = e13_alloc((void **)&q.std_, 8)) ||
        (rc = e13_alloc(&q.tmp, q.tmp_bytes))) return rc;
    q.cap = n;
    return 0;
}

static int e13_slot_plan(E13Slot &q, int n, bool forward, cufftHandle *out)
{
    std::map<int, cufftHandle> &cache = forward ? q.fwd : q.inv;
    auto it = cache.find(n);
    if (it != cache.end()) { *out = it->second; return 0; }
    cufftHandle h;
    cufftResult r = cufftPlan1d(&h, n, forward ? CUFFT_D2Z : CUFFT_Z2D, 1);
    if (r != CUFFT_SUCCESS) { mtg_set_error("cufftPlan1d(n=%d) failed: %d", n, (int)r); return MTG_ENOMEM; }
    cufftSetStream(h, q.st);
    cache[n] = h; *out = h;
    return 0;
}

// enqueue the set-up of one lightcurve on slot q: q.x[1] holds the TK95 segment (mean included) when src == nullptr
static int e13_slot_begin(E13Slot &q, const double *src, int n, double mean, int pdf, const double *variates, uint64_t seed, uint64_t lc_global,
                          bool draws_ready)
{
    int rc;
    cufftHandle fwd;
    if ((rc = e13_slot_plan(q, n, true, &fwd))) return rc;
    const unsigned g = (unsigned)std::min((n + 255) / 256, 148 * 8);
    const int Mh = n / 2 + 1;
    if (src) k_e13_addmean<<<g, 256, 0, q.st>>>(src, n, mean, q.x[1]);
    if (!draws_ready) {
        k_e13_moments<<<E13_NB, 256, 0, q.st>>>(q.x[1], n, q.part);
        k_e13_moments_fin<<<1, 32, 0, q.st>>>(q.part, E13_NB, n, q.std_);
        k_e13_draw<<<dim3(g, 1), 256, 0, q.st>>>(pdf, mean, q.std_, n, variates, 0, nullptr, seed, lc_global, q.x[0]);
    }
    if (cufftExecD2Z(fwd, q.x[1], (cufftDoubleComplex *)q.F) != CUFFT_SUCCESS) { mtg_set_error("cufftExecD2Z failed"); return MTG_EINVAL; }
    k_e13_amplitude<<<g, 256, 0, q.st>>>(q.F, (size_t)Mh, 1.0 / (double)(n / 2 + 1), q.A);
    size_t bytes = q.tmp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortKeysDescending(q.tmp, bytes, q.x[0], q.sorted, n, 0, 64, q.st);
    if (e != cudaSuccess) return mtg_cuda_check(e, "cub radix sort");
    mtg_count_launch(16);
    q.n = n; q.it = 0; q.cur = 0;
    return mtg_cuda_check(cudaGetLastError(), "E13 set-up");
}

// the work of one iteration on slot q's stream; its allclose flag lands in q.h_flag
static int e13_slot_enqueue(E13Slot &q, const int32_t *iota)
{
    int rc;
    const int n = q.n, Mh = n / 2 + 1;
    cufftHandle fwd, inv;
    if ((rc = e13_slot_plan(q, n, true, &fwd)) || (rc = e13_slot_plan(q, n, false, &inv))) return rc;
    const unsigned g = (unsigned)std::min((n + 255) / 256, 148 * 8);
    double *x = q.x[q.cur], *adj = q.x[q.cur ^ 1];
    if (cufftExecD2Z(fwd, x, (cufftDoubleComplex *)q.F) != CUFFT_SUCCESS) { mtg_set_error("cufftExecD2Z failed"); return MTG_EINVAL; }
    k_e13_phasor<<<g, 256, 0, q.st>>>(q.F, q.A, (size_t)Mh, 1.0 / (double)n);
    if (cufftExecZ2D(inv, (cufftDoubleComplex *)q.F, adj) != CUFFT_SUCCESS) { mtg_set_error("cufftExecZ2D failed"); return MTG_EINVAL; }
    k_e13_keys32<<<g, 256, 0, q.st>>>(adj, (size_t)n, q.k32);
    size_t bytes = q.tmp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairsDescending(q.tmp, bytes, q.k32, q.k32o, iota, q.perm, n, 0, 32, q.st);
    if (e != cudaSuccess) return mtg_cuda_check(e, "cub radix sort (32-bit keys)");
    k_e13_fix_ties<<<g, 256, 0, q.st>>>(q.k32o, adj, n, q.perm);
    cudaMemsetAsync(q.flag, 0, sizeof(int32_t), q.st);
    k_e13_match_check<<<g, 256, 0, q.st>>>(q.perm, q.sorted, x, adj, n, q.flag);
    cudaMemcpyAsync(q.h_flag, q.flag, sizeof(int32_t), cudaMemcpyDeviceToHost, q.st);
    return 0;
}

// MTG_E13_GRAPH=0: enqueue every iteration call by call instead of replaying its captured CUDA graph
static bool e13_use_graphs()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("MTG_E13_GRAPH"); v = e ? (atoi(e) != 0) : 1; }
    return v != 0;
}

// enqueue one iteration on slot q; q.ev completes when its flag is in q.h_flag.  An iteration is ~21 launches (two cuFFT
// transforms of four passes each, the radix sort's six kernels, five of our own) for ~150 us of device work, and one host
// thread feeds every slot: the sequence is captured once per (length, buffer parity) and replayed as ONE graph launch, so a
// slow host does not become the bottleneck (measured on a fast one: 288 -> 317 lightcurves/s with two slots, 351 -> 366 with
// three).  Only for the lengthened grid, where every series has the same length: on the reference grid the two segment
// lengths, two buffer parities and Bluestein plans make the captures themselves cost more than they save (120 -> 95 /s).
static int e13_slot_iterate(E13Slot &q, const int32_t *iota, bool allow_graph)
{
    int rc;
    if (allow_graph && e13_use_graphs() && q.graphs_ok) {
        const long long key = ((long long)q.n << 1) | (long long)q.cur;
        cudaGraphExec_t ge = nullptr;
        auto it = q.graphs.find(key);
        if (it != q.graphs.end()) ge = it->second;
        else {
            cufftHandle f, b;                     // plan creation allocates: not inside a capture
            if ((rc = e13_slot_plan(q, q.n, true, &f)) || (rc = e13_slot_plan(q, q.n, false, &b))) return rc;
            cudaGraph_t g = nullptr;
            if (cudaStreamBeginCapture(q.st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                rc = e13_slot_enqueue(q, iota);
                const cudaError_t e = cudaStreamEndCapture(q.st, &g);
                if (!rc && e == cudaSuccess && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess) q.graphs[key] = ge;
                else { ge = nullptr; q.graphs_ok = false; }
                if (g) cudaGraphDestroy(g);
            } else q.graphs_ok = false;
            cudaGetLastError();                   // a failed capture falls back to plain launches below
        }
        if (ge) {
            MTG_CUDA(cudaGraphLaunch(ge, q.st));
            cudaEventRecord(q.ev, q.st);
            mtg_count_launch(20);
            return mtg_cuda_check(cudaGetLastError(), "E13 iteration (graph)");
        }
    }
    if ((rc = e13_slot_enqueue(q, iota))) return rc;
    cudaEventRecord(q.ev, q.st);
    mtg_count_launch(20);
    return mtg_cuda_check(cudaGetLastError(), "E13 iteration");
}

static int e13_slots()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("MTG_E13_SLOTS"); v = e ? atoi(e) : 4; if (v < 1) v = 1; if (v > 8) v = 8; }
    return v;
}

static int e13_large_prepare(E13Large &g, int n_max, int count)
{
    if ((int)g.slots.size() != e13_slots()) { g.release(); g.slots.resize(e13_slots()); }
    if (!g.ev_in) MTG_CUDA(cudaEventCreateWithFlags(&g.ev_in, cudaEventDisableTiming));
    if (g.iota_n < n_max) {
        for (auto &q : g.slots) q.drop_graphs();             // the captured iterations point at the old index array
        cudaFree(g.iota); g.iota = nullptr; g.iota_n = 0;
        int rc = e13_alloc((void **)&g.iota, sizeof(int32_t) * (size_t)n_max);
        if (rc) return rc;
        k_e13_iota<<<148 * 8, 256>>>(g.iota, n_max, 1);
        MTG_CUDA(cudaDeviceSynchronize());
        g.iota_n = n_max;
    }
    if (g.h_iters_n < count) {
        if (g.h_iters) cudaFreeHost(g.h_iters);
        g.h_iters = nullptr; g.h_iters_n = 0;
        MTG_CUDA(cudaMallocHost((void **)&g.h_iters, sizeof(int32_t) * (size_t)count));
        g.h_iters_n = count;
    }
    const int used = std::min<int>((int)g.slots.size(), count);
    for (int k = 0; k < used; ++k) { int rc = e13_slot_reserve(g.slots[k], n_max); if (rc) return rc; }
    return 0;
}

// Adjust `count` long series.  Series k: `n[k]` points at src[k] (device; the TK95 segment WITHOUT the mean), result written to
// dst[k] (may alias src[k]); `stream` orders the inputs before and the results after.  draws0: injected initial draws (device,
// one series, mean already in src) -- the test entry.
static int e13_run_large(E13Large &g, int count, const int *n, double *const *src, double *const *dst, double mean, int pdf, int max_iter,
                         const double *const *variates, uint64_t seed, uint64_t lc_offset, int32_t *d_iters, cudaStream_t stream,
                         const double *draws0 = nullptr)
{
    int n_max = 0;
    for (int k = 0; k < count; ++k) n_max = std::max(n_max, n[k]);
    int rc = e13_large_prepare(g, n_max, count);
    if (rc) return rc;
    MTG_CUDA(cudaEventRecord(g.ev_in, stream));
    const int used = std::min<int>((int)g.slots.size(), count);
    int next = 0, active = 0;
    auto start = [&](E13Slot &q) -> int {
        const int k = next++;
        q.lc = k; q.dst = dst[k];
        int r;
        if (draws0) {
            MTG_CUDA(cudaMemcpyAsync(q.x[0], draws0, sizeof(double) * n[k], cudaMemcpyDeviceToDevice, q.st));
            MTG_CUDA(cudaMemcpyAsync(q.x[1], src[k], sizeof(double) * n[k], cudaMemcpyDeviceToDevice, q.st));
            r = e13_slot_begin(q, nullptr, n[k], mean, pdf, nullptr, seed, lc_offset + k, true);
        } else r = e13_slot_begin(q, src[k], n[k], mean, pdf, variates ? variates[k] : nullptr, seed, lc_offset + k, false);
        if (!r) r = e13_slot_iterate(q, g.iota, g.allow_graphs);
        if (!r) ++active;
        return r;
    };
    for (int k = 0; k < used && !rc; ++k) {
        g.slots[k].lc = -1;
        MTG_CUDA(cudaStreamWaitEvent(g.slots[k].st, g.ev_in, 0));
        rc = start(g.slots[k]);
    }
    while (active > 0 && !rc) {
        for (int k = 0; k < used && !rc; ++k) {
            E13Slot &q = g.slots[k];
            if (q.lc < 0) continue;
            if ((rc = mtg_cuda_check(cudaEventSynchronize(q.ev), "E13 iteration"))) break;
            const bool notclose = *q.h_flag != 0;
            if (notclose && q.it < max_iter) { q.cur ^= 1; q.it += 1; rc = e13_slot_iterate(q, g.iota, g.allow_graphs); continue; }
            // converged (or out of iterations): the result is the CURRENT iterate, not the candidate (simulator.py:129)
            MTG_CUDA(cudaMemcpyAsync(q.dst, q.x[q.cur], sizeof(double) * q.n, cudaMemcpyDeviceToDevice, q.st));
            g.h_iters[q.lc] = q.it;
            q.lc = -1; --active;
            if (next < count) rc = start(q);
        }
    }
    for (int k = 0; k < used; ++k) {
        cudaEventRecord(g.slots[k].ev, g.slots[k].st);
        cudaStreamWaitEvent(stream, g.slots[k].ev, 0);
    }
    if (!rc && d_iters) rc = mtg_cuda_check(cudaMemcpyAsync(d_iters, g.h_iters, sizeof(int32_t) * count, cudaMemcpyHostToDevice, stream), "E13 iterations");
    if (rc) cudaDeviceSynchronize();
    return rc;
}

void sim_e13_release(Sim *s)
{
    if (s->e13_large) {
        E13Large *g = (E13Large *)s->e13_large;
        g->release();
        delete g;
        s->e13_large = nullptr;
    }
    if (!s->e13_cache) return;
    E13Work *w = (E13Work *)s->e13_cache;
    w->release();
    delete w;
    s->e13_cache = nullptr;
}

// smallest 7-smooth integer >= v
static int e13_next_smooth(int v)
{
    for (;; ++v) {
        int r = v;
        for (int f : {2, 3, 5, 7}) while (r % f == 0) r /= f;
        if (r == 1) return v;
    }
}

// E13 on the chunk held in s->d_lc.  Reference grid: the lightcurves are grouped by segment length (it varies by one
// sample with the random shift) and each group is adjusted as a batch.  Lengthened grid (MTG_SIM_E13_PADDED): every
// lightcurve is adjusted on a window of the same 7-smooth length starting at its segment, one batch, no host round trip.
int sim_e13_chunk(Sim *s, int Lc, int l0, uint64_t seed, const double *d_pdf_variates, int32_t *d_iters, cudaStream_t stream)
{
    const double *var = d_pdf_variates ? d_pdf_variates + (size_t)l0 * s->cfg.n_fft : nullptr;      // injected variates: indexed by the GLOBAL lightcurve
    const int max_iter = s->cfg.max_iter > 0 ? s->cfg.max_iter : 400;
    int n_pad = 0;
    if (s->cfg.flags & MTG_SIM_E13_PADDED) {
        const int need = (int)ceil(s->cfg.sim_duration / s->cfg.sim_dt) + 2;
        n_pad = e13_next_smooth(need < 2 ? 2 : need);
        if (n_pad > s->cfg.n_fft) n_pad = 0;
    }
    if (n_pad >= 32768 || (!n_pad && s->cfg.sim_duration / s->cfg.sim_dt >= 32770.0)) {
        // long series: one lightcurve per slot (e13_run_large)
        std::vector<SegInfo> seg(Lc);
        MTG_CUDA(cudaMemcpyAsync(seg.data(), s->d_seg, sizeof(SegInfo) * Lc, cudaMemcpyDeviceToHost, stream));
        MTG_CUDA(cudaStreamSynchronize(stream));
        std::vector<int> nn(Lc);
        std::vector<double *> ptr(Lc);
        std::vector<const double *> vptr(Lc);
        bool all_long = true;
        for (int l = 0; l < Lc; ++l) {
            const int n = n_pad ? n_pad : seg[l].i1 - seg[l].i0;
            const int start = n_pad ? (seg[l].i0 + n <= s->cfg.n_fft ? seg[l].i0 : s->cfg.n_fft - n) : seg[l].i0;
            nn[l] = n; ptr[l] = s->d_lc + (size_t)l * s->cfg.n_fft + start;
            vptr[l] = var ? var + (size_t)l * s->cfg.n_fft : nullptr;
            if (n < 32768) all_long = false;
        }
        if (all_long) {
            if (!s->e13_large) s->e13_large = new E13Large();
            ((E13Large *)s->e13_large)->allow_graphs = n_pad != 0;
            return e13_run_large(*(E13Large *)s->e13_large, Lc, nn.data(), ptr.data(), ptr.data(), s->cfg.mean, s->cfg.pdf, max_iter,
                                 var ? vptr.data() : nullptr, seed, (uint64_t)l0, d_iters, stream);
        }
    }
    if (!s->e13_cache) s->e13_cache = new E13Work();
    E13Work &w = *(E13Work *)s->e13_cache;
    if (n_pad) {
        const int n = n_pad;
        {
            int rc = e13_work_alloc(w, Lc, n);
            if (rc) return rc;
            const dim3 g2((unsigned)std::min<long long>((n + 255) / 256, 1024), Lc);
            k_e13_gather<<<g2, 256, 0, stream>>>(s->d_lc, s->cfg.n_fft, s->d_seg, nullptr, n, s->cfg.mean, w.seg);
            mtg_count_launch();
            rc = e13_run_batch(s, s->plans, s->plans_d2z, w, Lc, n, s->cfg.pdf, s->cfg.mean, max_iter, var, (size_t)s->cfg.n_fft, nullptr, seed,
                               (uint64_t)l0, stream);
            if (rc) return rc;
            k_e13_scatter_back<<<g2, 256, 0, stream>>>(s->d_lc, s->cfg.n_fft, s->d_seg, nullptr, n, w.xsim, w.iters, d_iters);
            mtg_count_launch();
            return mtg_cuda_check(cudaGetLastError(), "E13 (lengthened window)");
        }
    }
    std::vector<SegInfo> seg(Lc);
    MTG_CUDA(cudaMemcpyAsync(seg.data(), s->d_seg, sizeof(SegInfo) * Lc, cudaMemcpyDeviceToHost, stream));
    MTG_CUDA(cudaStreamSynchronize(stream));
    std::map<int, std::vector<int>> groups;
    int n_max = 0;
    for (int l = 0; l < Lc; ++l) { const int n = seg[l].i1 - seg[l].i0; groups[n].push_back(l); n_max = std::max(n_max, n); }
    int *d_members = nullptr;
    int rc = e13_alloc((void **)&d_members, sizeof(int) * Lc);
    for (auto &kv : groups) {
        if (rc) break;
        const int n = kv.first, cnt = (int)kv.second.size();
        if (n < 2) { mtg_set_error("E13: segment of %d points", n); rc = MTG_EINVAL; break; }
        // one workspace for all groups of the chunk: sized for the longest segment and the whole chunk
        rc = e13_work_alloc(w, Lc, (n >= 32768) == (n_max >= 32768) ? n_max : n, cnt, n);
        if (!rc) rc = mtg_cuda_check(cudaMemcpyAsync(d_members, kv.second.data(), sizeof(int) * cnt, cudaMemcpyHostToDevice, stream), "members");
        if (!rc) {
            const dim3 g2((unsigned)std::min<long long>((n + 255) / 256, 1024), cnt);
            k_e13_gather<<<g2, 256, 0, stream>>>(s->d_lc, s->cfg.n_fft, s->d_seg, d_members, n, s->cfg.mean, w.seg);
            mtg_count_launch();
            rc = e13_run_batch(s, s->plans, s->plans_d2z, w, cnt, n, s->cfg.pdf, s->cfg.mean, max_iter, var, (size_t)s->cfg.n_fft, d_members, seed,
                               (uint64_t)l0, stream);
            if (!rc) {
                k_e13_scatter_back<<<g2, 256, 0, stream>>>(s->d_lc, s->cfg.n_fft, s->d_seg, d_members, n, w.xsim, w.iters, d_iters);
                mtg_count_launch();
                rc = mtg_cuda_check(cudaStreamSynchronize(stream), "E13 group sync");      // d_members is reused by the next group
            }
        }
    }
    cudaFree(d_members);
    return rc;
}

extern "C" int mtg_e13_adjust_host(int n, const double *countrate, const double *xsim0, int max_iter,
                                   double *out_x, int32_t *out_iters)
{
    if (n < 2 || !countrate || !xsim0 || !out_x || max_iter < 0) { mtg_set_error("mtg_e13_adjust_host: bad argument"); return MTG_EINVAL; }
    if (n >= 32768) {              // the long-series path (one slot)
        E13Large g;
        double *d_in = nullptr, *d_x0 = nullptr; int32_t *d_it = nullptr;
        int rc = e13_alloc((void **)&d_in, sizeof(double) * n);
        if (!rc) rc = e13_alloc((void **)&d_x0, sizeof(double) * n);
        if (!rc) rc = e13_alloc((void **)&d_it, sizeof(int32_t));
        if (!rc) rc = mtg_cuda_check(cudaMemcpy(d_in, countrate, sizeof(double) * n, cudaMemcpyHostToDevice), "h2d");
        if (!rc) rc = mtg_cuda_check(cudaMemcpy(d_x0, xsim0, sizeof(double) * n, cudaMemcpyHostToDevice), "h2d");
        double *srcs[1] = {d_in};
        if (!rc) rc = e13_run_large(g, 1, &n, srcs, srcs, 0.0, MTG_PDF_UNIFORM, max_iter, nullptr, 0, 0, d_it, (cudaStream_t)0, d_x0);
        if (!rc) rc = mtg_cuda_check(cudaDeviceSynchronize(), "E13");
        if (!rc) rc = mtg_cuda_check(cudaMemcpy(out_x, d_in, sizeof(double) * n, cudaMemcpyDeviceToHost), "d2h");
        if (!rc && out_iters) rc = mtg_cuda_check(cudaMemcpy(out_iters, d_it, sizeof(int32_t), cudaMemcpyDeviceToHost), "d2h");
        cudaFree(d_in); cudaFree(d_x0); cudaFree(d_it);
        g.release();
        return rc;
    }
    Sim tmp;                      // only used as a cuFFT plan cache
    E13Work w;
    int rc = e13_work_alloc(w, 1, n);
    if (!rc) rc = mtg_cuda_check(cudaMemcpy(w.seg, countrate, sizeof(double) * n, cudaMemcpyHostToDevice), "h2d");
    if (!rc) rc = mtg_cuda_check(cudaMemcpy(w.xsim, xsim0, sizeof(double) * n, cudaMemcpyHostToDevice), "h2d");
    if (!rc) rc = e13_run_batch(&tmp, tmp.plans, tmp.plans_d2z, w, 1, n, MTG_PDF_UNIFORM, 0.0, max_iter, nullptr, 0, nullptr, 0, 0,
                                (cudaStream_t)0, /*draws_ready=*/true);
    if (!rc) rc = mtg_cuda_check(cudaMemcpy(out_x, w.xsim, sizeof(double) * n, cudaMemcpyDeviceToHost), "d2h");
    if (!rc && out_iters) rc = mtg_cuda_check(cudaMemcpy(out_iters, w.iters, sizeof(int32_t), cudaMemcpyDeviceToHost), "d2h");
    for (auto &kv : tmp.plans) cufftDestroy(kv.second);
    for (auto &kv : tmp.plans_d2z) cufftDestroy(kv.second);
    w.release();
    return rc;
}
