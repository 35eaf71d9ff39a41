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
        const double r = hypot(f.x, f.y), a = A[i] * inv_n;
        F[i] = r > 0.0 ? make_double2(a * (f.x / r), a * (f.y / r)) : make_double2(a, 0.0);
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

void sim_e13_release(Sim *s)
{
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
    if (!s->e13_cache) s->e13_cache = new E13Work();
    E13Work &w = *(E13Work *)s->e13_cache;
    const double *var = d_pdf_variates ? d_pdf_variates + (size_t)l0 * s->cfg.n_fft : nullptr;      // injected variates: indexed by the GLOBAL lightcurve
    const int max_iter = s->cfg.max_iter > 0 ? s->cfg.max_iter : 400;
    if (s->cfg.flags & MTG_SIM_E13_PADDED) {
        const int need = (int)ceil(s->cfg.sim_duration / s->cfg.sim_dt) + 2;
        const int n = e13_next_smooth(need < 2 ? 2 : need);
        if (n <= s->cfg.n_fft) {
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
