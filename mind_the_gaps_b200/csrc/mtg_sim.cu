// mtg_sim.cu -- the lightcurve simulator on the device (compiled with -fmad=false so that every
// add / multiply is rounded separately, exactly like the NumPy expressions it restates: the index
// work -- segment cut, bin membership, count thresholds -- must agree bit for bit).
//
//   K3  k_spectrum   Timmer & Koenig half-spectrum: Philox normals x sqrt(S(w)/2), DC / Nyquist fix-ups
//                    (get_fft, mind_the_gaps/simulator.py:468-501), inverse-FFT scale folded in
//       cuFFT Z2D    batched inverse real FFT (library call; simulator.py:386)
//   K5  k_segment    random segment: uniform offset + two searchsorted (simulator.py:536-539, :414)
//       k_bin        exposure binning: exact bin membership + numpy-order pairwise mean (simulator.py:340-367)
//   K7  k_noise      Poisson / Kraft / Gaussian noise and uncertainties (noise_models.py:49-184)
// K6 (Emmanoulopoulos PDF adjustment) lives in mtg_e13.cu.
#include "mtg_sim.cuh"
#include <mutex>
#include <vector>

#include "mtg_sim_dev.cuh"

// ---- celerite PSD coefficients for one parameter vector (runtime J) ---------------------------------
__global__ void k_psd_coeffs(const __grid_constant__ Plan pl, const double *__restrict__ params, int L, PsdCoef *out)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const double *x = params + (size_t)l * pl.n_free;
    PsdCoef c; c.nr = 0; c.nc = 0;
    for (int j = 0; j < pl.jr; ++j) {
        c.ar[c.nr] = plan_real_a(pl, x, j); c.cr[c.nr] = plan_real_c(pl, x, j); ++c.nr;
    }
    for (int j = 0; j < pl.jc; ++j) {
        const double p0 = plan_param(pl, x, pl.c_idx[j][0]), p1 = plan_param(pl, x, pl.c_idx[j][1]);
        const double p2 = plan_param(pl, x, pl.c_idx[j][2]), p3 = plan_param(pl, x, pl.c_idx[j][3]);
        double a, b = 0.0, cc, d;
        switch (pl.c_kind[j]) {
        case CS_COMPLEX3: a = exp(p0); cc = exp(p1); d = exp(p2); break;
        case CS_COMPLEX4: a = exp(p0); b = exp(p1); cc = exp(p2); d = exp(p3); break;
        case CS_RAW: a = p0; b = p1; cc = p2; d = p3; break;
        case CS_LORENTZIAN: a = exp(p0); cc = 0.5 * exp(p2) / exp(p1); d = exp(p2); break;
        case CS_COSINUS: a = exp(p0); cc = 0.0; d = exp(p1); break;
        case CS_BENDINGPL: a = exp(p0); b = exp(p1); cc = exp(p2); d = cc; break;
        case CS_MATERN32: { const double w0 = sqrt(3.0) * exp(-p1), S0 = exp(2.0 * p0) / w0;
                            a = w0 * S0; b = w0 * w0 * S0 / pl.c_extra[j]; cc = w0; d = pl.c_extra[j]; } break;
        default: { const double S0 = exp(p0), Q = exp(p1), w0 = exp(p2);
                   if (Q >= 0.5) { const double f = sqrt(4.0 * Q * Q - 1.0); a = S0 * w0 * Q; b = a / f; cc = 0.5 * w0 / Q; d = cc * f; }
                   else { const double f = sqrt(1.0 - 4.0 * Q * Q);
                          c.ar[c.nr] = 0.5 * S0 * w0 * Q * (1.0 + 1.0 / f); c.cr[c.nr] = 0.5 * w0 / Q * (1.0 - f); ++c.nr;
                          c.ar[c.nr] = 0.5 * S0 * w0 * Q * (1.0 - 1.0 / f); c.cr[c.nr] = 0.5 * w0 / Q * (1.0 + f); ++c.nr;
                          continue; } } break;
        }
        c.ac[c.nc] = a; c.bc[c.nc] = b; c.cc[c.nc] = cc; c.dc[c.nc] = d; ++c.nc;
    }
    out[l] = c;
}

static inline void spectrum_scales(int n_fft, double sim_dt, double *val, double *fold)
{
    volatile double nd = (double)n_fft * sim_dt;             // volatile: no host-side contraction / reassociation
    *val = 1.0 / nd;
    volatile double arg = nd * 2.5066282746310002;
    *fold = sqrt(arg) / nd;
}

// K3: one thread per (frequency k, lightcurve l).  spec[l][k], k = 0..M-1.
__global__ void __launch_bounds__(256)
k_spectrum(int L, int n_fft, int M, double val, double fold, const PsdCoef *__restrict__ coefs,
           const double *__restrict__ psd_table, const double *__restrict__ normals, uint64_t seed,
           uint64_t lc_offset, int pair_stride, double2 *__restrict__ spec)
{
    __shared__ PsdCoef sc;
    const int l = blockIdx.y;
    if (coefs && threadIdx.x == 0) sc = coefs[l];
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= M) return;
    double2 out = make_double2(0.0, 0.0);
    if (k > 0) {
        // np.fft.rfftfreq(N, dt) * 2 * np.pi  ==  ((k * (1/(N dt))) * 2) * pi ; val = 1/(N dt) and the scale
        // fold = sqrt(N dt sqrt(2 pi)) / (N dt)  [rate = irfft(X) sqrt(N dt sqrt(2 pi)) / dt, numpy's irfft = Z2D / N,
        // simulator.py:386-391] are IEEE double results computed once on the host (spectrum_scales)
        const double w = (((double)k * val) * 2.0) * 3.141592653589793;
        if (normals) {
            // parity path: the reference's operations in IEEE arithmetic on the injected normals
            const double psd = coefs ? psd_eval(sc, w) : psd_table[k];
            const double re = normals[((size_t)l * 2) * M + k], im = normals[((size_t)l * 2 + 1) * M + k];
            const double amp = sqrt(0.5 * psd) * fold;
            out.x = re * amp; out.y = im * amp;
        } else {
            const double psd = coefs ? psd_eval_fast(sc, w) : psd_table[k];
            const unsigned row = (unsigned)k / (unsigned)pair_stride, col = (unsigned)k - row * (unsigned)pair_stride;
            float2 ze, zo;
            spectrum_normals(seed, spectrum_block(lc_offset + l, M, pair_stride, row, col), &ze, &zo);
            const float2 z = (row & 1u) ? zo : ze;
            const double amp = fast_sqrt(0.5 * psd) * fold;
            out.x = (double)z.x * amp; out.y = (double)z.y * amp;
        }
        if ((n_fft % 2 == 0) && k == M - 1) out.y = 0.0;          // Nyquist bin is real (simulator.py:499-500)
    }
    // k == 0: the reference sets an arbitrary DC (1e6) and removes the mean afterwards
    // (simulator.py:497, :393); DC = 0 gives the zero-mean series directly
    spec[(size_t)l * M + k] = out;
}

// first index i in [0, n) with grid[i] >= x   (np.searchsorted, side="left")
__device__ __forceinline__ int lower_bound_grid(const double *__restrict__ g, int n, double x)
{
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (g[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

// K5a: cut_random_segment + shift (simulator.py:536-539, :412-414)
__global__ void k_segment(int L, int n_fft, const double *__restrict__ grid, double sim_dt, double duration,
                          double first_bin_start, const double *__restrict__ u_in, uint64_t seed, uint64_t lc_offset,
                          SegInfo *__restrict__ seg)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const double u = u_in ? u_in[l] : philox_u2(seed, 1u, lc_offset + l, 0u).x;
    const double low = grid[0], high = grid[n_fft - 1] - duration;
    const double shift = low + (high - low) * u;                   // np.random.uniform(low, high)
    SegInfo s;
    s.i0 = lower_bound_grid(grid, n_fft, shift);
    s.i1 = lower_bound_grid(grid, n_fft, shift + duration);
    if (s.i0 >= n_fft) s.i0 = n_fft - 1;
    const double tstart = grid[s.i0] - 0.5 * sim_dt;               // stingray: tstart = time[0] - dt/2
    s.time_shift = -tstart + first_bin_start;
    seg[l] = s;
}

// first j in [0, len) with (time(j) >= x), time(j) = base[j] + shift, starting from a guess
template <class TimeFn>
__device__ __forceinline__ int first_ge(TimeFn time, int len, double x, int guess)
{
    int j = min(max(guess, 0), len);
    while (j > 0 && time(j - 1) >= x) --j;
    while (j < len && time(j) < x) ++j;
    return j;
}

// K5b: exposure binning.  thread -> (l fastest, n); out_tm[n][L_total] at column l_out0 + l.
__global__ void __launch_bounds__(256)
k_bin(int L, int n_obs, int n_fft, const double *__restrict__ grid, const double *__restrict__ lc, double sim_dt,
      const SegInfo *__restrict__ seg, const double *__restrict__ bin_lo, const double *__restrict__ bin_hi,
      double mean_add, double *__restrict__ out_tm, int L_total, int l_out0)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)L * n_obs) return;
    const int l = (int)(idx % L), n = (int)(idx / L);
    const SegInfo s = seg[l];
    const int len = s.i1 - s.i0;
    const double *g = grid + s.i0;
    const double sh = s.time_shift;
    auto time = [&](int j) { return g[j] + sh; };                  // segment.shift(): time + shift
    const double lo = bin_lo[n], hi = bin_hi[n];
    const double t0 = time(0);
    const int j_lo = first_ge(time, len, lo, (int)((lo - t0) / sim_dt));
    const int j_hi = first_ge(time, len, hi, (int)((hi - t0) / sim_dt) + 1);
    const int cnt = j_hi - j_lo;
    const double *src = lc + (size_t)l * n_fft + s.i0 + j_lo;
    double r;
    if (cnt <= 0) r = nan("");
    else if (mean_add == 0.0) r = np_mean(src, cnt);
    else {
        // countrate - meanrate + mean is applied per sample before averaging (simulator.py:393)
        double buf[32];
        if (cnt <= 32) { for (int i = 0; i < cnt; ++i) buf[i] = src[i] + mean_add; r = np_mean(buf, cnt); }
        else {
            // long bins: mean(x + m) evaluated as mean(x) + m (differs from the reference only in rounding)
            r = np_mean(src, cnt) + mean_add;
        }
    }
    out_tm[(size_t)n * L_total + l_out0 + l] = r;
}

// K5c: which samples does every bin of every lightcurve need?  Same membership arithmetic as k_bin; feeds the fused
// TK95 path (mtg_tk95.cu).  thread -> (l, n) with n fastest.  mslot[l][q_off[n] + i] = absolute sample index or -1.
__global__ void __launch_bounds__(256)
k_need(int L, int n_obs, const double *__restrict__ grid, double sim_dt, const SegInfo *__restrict__ seg,
       const double *__restrict__ bin_lo, const double *__restrict__ bin_hi, const int *__restrict__ q_off, int Qmax,
       int *__restrict__ mslot, int *__restrict__ cnt_out, int *__restrict__ flag)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)L * n_obs) return;
    const int l = (int)(idx / n_obs), n = (int)(idx % n_obs);
    const SegInfo s = seg[l];
    const int len = s.i1 - s.i0;
    const double *g = grid + s.i0;
    const double sh = s.time_shift;
    auto time = [&](int j) { return g[j] + sh; };
    const double lo = bin_lo[n], hi = bin_hi[n];
    const double t0 = time(0);
    const int j_lo = first_ge(time, len, lo, (int)((lo - t0) / sim_dt));
    const int j_hi = first_ge(time, len, hi, (int)((hi - t0) / sim_dt) + 1);
    int cnt = j_hi - j_lo;
    const int q0 = q_off[n], cap = q_off[n + 1] - q0;
    if (cnt > cap) { atomicExch(flag, 1); cnt = cap; }
    if (cnt < 0) cnt = 0;
    cnt_out[(size_t)l * n_obs + n] = cnt;
    int *ms = mslot + (size_t)l * Qmax + q0;
    for (int i = 0; i < cap; ++i) ms[i] = i < cnt ? s.i0 + j_lo + i : -1;
}

// ---- K7: noise ----------------------------------------------------------------------------------------
struct PhiloxStream {
    uint64_t seed, item; uint32_t purpose, draw; double spare; bool has;
    __device__ double next()
    {
        if (has) { has = false; return spare; }
        const double2 u = philox_u2(seed, purpose, item, draw++);
        spare = u.y; has = true; return u.x;
    }
};

// Poisson variate: multiplication method below 10, Hoermann's PTRS transformed rejection above
__device__ double poisson_draw(double lam, PhiloxStream &rs)
{
    if (lam == 0.0) return 0.0;
    if (lam < 10.0) {
        const double enlam = exp(-lam);
        double prod = rs.next(); long long k = 0;
        while (prod > enlam) { prod *= rs.next(); ++k; }
        return (double)k;
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (;;) {
        const double U = rs.next() - 0.5, V = rs.next();
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + k * loglam - lgamma(k + 1.0))) return k;
    }
}

// regularised upper incomplete gamma Q(a, x) for integer a >= 1: e^-x sum_{k<a} x^k / k!
__device__ double gammaq_int(int a, double x)
{
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < a; ++k) { term *= x / (double)k; sum += term; }
    return exp(-x) * sum;
}
// solve Q(a, x) = q for x >= 0 (Q decreasing): safeguarded Newton, dQ/dx = -e^-x x^(a-1)/(a-1)!
__device__ double gammaq_int_inv(int a, double q)
{
    if (q >= 1.0) return 0.0;
    double lo = 0.0, hi = (double)a + 40.0 + 10.0 * sqrt((double)a);
    double x = (double)a;
    for (int it = 0; it < 200; ++it) {
        const double f = gammaq_int(a, x) - q;
        if (f > 0.0) lo = x; else hi = x;
        const double d = -exp(-x + ((double)(a - 1)) * log(x) - lgamma((double)a));
        double xn = x - f / d;
        if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
        if (fabs(xn - x) <= 4e-16 * fmax(xn, 1.0)) { x = xn; break; }
        x = xn;
    }
    return x;
}
// Kraft, Burrows & Nousek (1991) posterior f(s) = C e^-(s+B) (s+B)^N / N!   (stats.py:15-18)
__device__ __forceinline__ double kraft_logpdf_unnorm(double s, int N, double B)
{ return -(s + B) + (N > 0 ? (double)N * log(s + B) : 0.0); }
// minimal-length credible interval (astropy poisson_conf_interval "kraft-burrows-nousek")
__device__ void kbn_interval(int N, double B, double CL, double *smin, double *smax)
{
    const double C = 1.0 / gammaq_int(N + 1, B);
    auto s_min_of = [&](double s) {
        const double y = kraft_logpdf_unnorm(s, N, B);
        if (kraft_logpdf_unnorm(0.0, N, B) >= y) return 0.0;
        double lo = 0.0, hi = (double)N - B;                       // f increasing on [0, N - B]
        for (int it = 0; it < 100 && hi - lo > 1e-15 * fmax(hi, 1.0); ++it) {
            const double mid = 0.5 * (lo + hi);
            if (kraft_logpdf_unnorm(mid, N, B) < y) lo = mid; else hi = mid;
        }
        return 0.5 * (lo + hi);
    };
    auto func = [&](double s) {
        const double a = s_min_of(s);
        return C * (gammaq_int(N + 1, a + B) - gammaq_int(N + 1, s + B)) - CL;
    };
    double lo = fmax((double)N - B, 0.0), hi = 100.0;
    for (int it = 0; it < 100 && hi - lo > 1e-14 * fmax(hi, 1.0); ++it) {
        const double mid = 0.5 * (lo + hi);
        if (func(mid) < 0.0) lo = mid; else hi = mid;
    }
    *smax = 0.5 * (lo + hi);
    *smin = s_min_of(*smax);
}

// thread -> (l fastest, n).  rates / outputs: element (n, l) at n*ld_n + l*ld_l  (time-major: ld_n = L, ld_l = 1)
__global__ void __launch_bounds__(256)
k_noise(int kind, int L, int n_obs, size_t ld_n, size_t ld_l, const double *__restrict__ rates,
        const double *__restrict__ exposures, const double *__restrict__ bkg_counts,
        const double *__restrict__ bkg_rate_err, double sigma, double kraft_counts, uint64_t seed,
        const long long *__restrict__ counts_in, const double *__restrict__ normals_in,
        double *__restrict__ out_rates, double *__restrict__ out_dy, int32_t *__restrict__ status)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)L * n_obs) return;
    const int l = (int)(idx % L), n = (int)(idx / L);
    const size_t at = (size_t)n * ld_n + (size_t)l * ld_l;
    const double rate = rates[at];
    int st = 0;
    double noisy, dy;
    if (kind == MTG_NOISE_GAUSSIAN) {
        const double z = normals_in ? normals_in[at] : philox_normal2(seed, 3u, (uint64_t)l * n_obs + n, 0u).x;
        noisy = rate + sigma * z;                                  // noise_models.py:181
        dy = sigma;
    } else {
        const double ex = exposures[n], bkg = bkg_counts ? bkg_counts[n] : 0.0, err = bkg_rate_err ? bkg_rate_err[n] : 0.0;
        const double total = rate * ex + bkg;                      // noise_models.py:70
        double k;
        if (counts_in) k = (double)counts_in[at];
        else if (!(total >= 0.0)) { k = nan(""); st = 1; }         // numpy: ValueError("lam < 0")
        else { PhiloxStream rs{seed, (uint64_t)l * n_obs + n, 3u, 0u, 0.0, false}; k = poisson_draw(total, rs); }
        const double net = k - bkg;
        noisy = net / ex;                                          // :74-78
        const double q = sqrt(k) / ex;
        dy = sqrt(q * q + err * err);                              // :76
        if (kind == MTG_NOISE_KRAFT && st == 0) {
            const double tc = noisy * ex + bkg;                    // :129
            if (tc < kraft_counts) {                               // :136
                const int Nr = (int)rint(tc);                      // np.round(...).astype(int)  :140
                const int Nt = (int)tc;                            // .astype(int)               :144
                const double med = gammaq_int_inv(Nr + 1, 0.5 * gammaq_int(Nr + 1, bkg)) - bkg;
                noisy = med / ex;                                  // :141
                double a, b;
                kbn_interval(Nt, bkg, 0.68, &a, &b);
                dy = (b - a) / 2.0 / ex;                           // :146
            }
        }
    }
    out_rates[at] = noisy; out_dy[at] = dy;
    if (status) status[at] = st;
}

// ---- standalone exact down-sampling on explicit arrays ---------------------------------------------
__global__ void k_downsample_explicit(int n_seg, const double *__restrict__ time, const double *__restrict__ rate,
                                      int n_obs, const double *__restrict__ bin_lo, const double *__restrict__ bin_hi,
                                      double *__restrict__ out, int32_t *__restrict__ first, int32_t *__restrict__ count)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_obs) return;
    const int j_lo = lower_bound_grid(time, n_seg, bin_lo[n]);
    const int j_hi = lower_bound_grid(time, n_seg, bin_hi[n]);
    const int cnt = j_hi - j_lo;
    out[n] = cnt > 0 ? np_mean(rate + j_lo, cnt) : nan("");
    if (first) first[n] = j_lo;
    if (count) count[n] = cnt > 0 ? cnt : 0;
}

// =====================================================================================================
// host side
// =====================================================================================================
static int sim_alloc(void **p, size_t bytes)
{
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 8);
    if (e != cudaSuccess) { mtg_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return MTG_ENOMEM; }
    return 0;
}
static int upload(double **dst, const double *src, size_t n, cudaStream_t st)
{
    int rc = sim_alloc((void **)dst, n * sizeof(double));
    if (rc) return rc;
    return mtg_cuda_check(cudaMemcpyAsync(*dst, src, n * sizeof(double), cudaMemcpyHostToDevice, st), "upload");
}

extern "C" int mtg_sim_create(const mtg_sim_config_t *cfg, void **out)
{
    if (!cfg || !out || cfg->n_obs < 1 || cfg->n_fft < 2 || !(cfg->sim_dt > 0) || !cfg->grid || !cfg->bin_lo ||
        !cfg->bin_hi || !cfg->exposures) { mtg_set_error("mtg_sim_create: bad configuration"); return MTG_EINVAL; }
    if (cfg->pdf < MTG_PDF_GAUSSIAN || cfg->pdf > MTG_PDF_UNIFORM || cfg->noise < MTG_NOISE_POISSON || cfg->noise > MTG_NOISE_GAUSSIAN) {
        mtg_set_error("mtg_sim_create: unknown pdf / noise kind"); return MTG_EINVAL;
    }
    Sim *s = new Sim();
    s->cfg = *cfg;
    s->M = cfg->n_fft / 2 + 1;
    int rc = mtg_cuda_check(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking), "stream");
    if (!rc) rc = upload(&s->d_grid, cfg->grid, cfg->n_fft, s->stream);
    if (!rc) rc = upload(&s->d_lo, cfg->bin_lo, cfg->n_obs, s->stream);
    if (!rc) rc = upload(&s->d_hi, cfg->bin_hi, cfg->n_obs, s->stream);
    if (!rc) rc = upload(&s->d_exp, cfg->exposures, cfg->n_obs, s->stream);
    if (!rc && cfg->bkg_counts) rc = upload(&s->d_bkg, cfg->bkg_counts, cfg->n_obs, s->stream);
    if (!rc && cfg->bkg_rate_err) rc = upload(&s->d_err, cfg->bkg_rate_err, cfg->n_obs, s->stream);
    if (!rc) rc = mtg_cuda_check(cudaStreamSynchronize(s->stream), "sim create sync");
    s->first_bin_start = cfg->bin_lo[0];
    if (!rc && cfg->pdf == MTG_PDF_GAUSSIAN) {
        std::vector<int> q_off;
        tk_geometry(cfg->n_fft, cfg->n_obs, cfg->bin_lo, cfg->bin_hi, cfg->sim_dt, &s->tk, &q_off);
        if (s->tk.ok) {
            std::vector<double2> tw;
            tk_twiddles(s->tk, &tw);
            rc = sim_alloc((void **)&s->d_tw, tw.size() * sizeof(double2));
            if (!rc) rc = mtg_cuda_check(cudaMemcpy(s->d_tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice), "twiddles");
            if (!rc) rc = sim_alloc((void **)&s->d_qoff, q_off.size() * sizeof(int));
            if (!rc) rc = mtg_cuda_check(cudaMemcpy(s->d_qoff, q_off.data(), q_off.size() * sizeof(int), cudaMemcpyHostToDevice), "slot offsets");
            if (!rc) rc = sim_alloc((void **)&s->d_flag, sizeof(int));
            if (!rc) rc = mtg_cuda_check(cudaMemset(s->d_flag, 0, sizeof(int)), "flag");
        }
    }
    s->cfg.grid = s->cfg.bin_lo = s->cfg.bin_hi = s->cfg.exposures = s->cfg.bkg_counts = s->cfg.bkg_rate_err = nullptr;
    if (rc) { mtg_sim_destroy(s); return rc; }
    *out = s;
    return 0;
}

// ---- workspace pool -------------------------------------------------------------------------------
// A simulator's chunk buffers are gigabytes (n_fft ~ 2.7e6: 43 MB per lightcurve in flight) and its cuFFT plans
// (Bluestein for the reference's un-padded lengths) take longer to build than a chunk takes to run.  The LRT driver
// creates one simulator per pipeline slot per call, so destroyed simulators park their buffers and plans here and
// the next simulator of the same FFT length on the same device adopts them instead of paying cudaMalloc / cudaFree
// (which synchronises the device) / cufftPlanMany again.  At most MTG_SIM_POOL (default 4) workspaces are kept;
// mtg_sim_trim() releases them.
struct SimScratch {
    int device = -1, n_fft = 0, cap = 0;
    double2 *d_spec = nullptr; double *d_lc = nullptr; PsdCoef *d_coef = nullptr; SegInfo *d_seg = nullptr;
    std::map<long long, cufftHandle> plans, plans_d2z;
    int capF = 0, Qmax = 0, n_obs = 0;                      // fused TK95 chunk buffers
    double2 *d_Y = nullptr; int *d_mslot = nullptr, *d_cnt = nullptr; PsdCoef *d_coefF = nullptr; SegInfo *d_segF = nullptr;
};
static std::mutex g_pool_mu;
static std::vector<SimScratch> g_pool;

static void scratch_release(SimScratch &w)
{
    for (auto &kv : w.plans) cufftDestroy(kv.second);
    for (auto &kv : w.plans_d2z) cufftDestroy(kv.second);
    cudaFree(w.d_spec); cudaFree(w.d_lc); cudaFree(w.d_coef); cudaFree(w.d_seg);
    cudaFree(w.d_Y); cudaFree(w.d_mslot); cudaFree(w.d_cnt); cudaFree(w.d_coefF); cudaFree(w.d_segF);
}
static int pool_limit()
{
    if (const char *e = getenv("MTG_SIM_POOL")) return atoi(e);
    return 4;
}

extern "C" int mtg_sim_is_fused(void *sim)
{
    const Sim *s = (const Sim *)sim;
    return s && s->tk.ok && s->cfg.pdf == MTG_PDF_GAUSSIAN ? 1 : 0;
}

extern "C" int mtg_sim_trim(void)
{
    std::lock_guard<std::mutex> lock(g_pool_mu);
    int dev = 0; cudaGetDevice(&dev);
    for (auto &w : g_pool) { cudaSetDevice(w.device); scratch_release(w); }
    g_pool.clear();
    cudaSetDevice(dev);
    return 0;
}

extern "C" int mtg_sim_destroy(void *p)
{
    Sim *s = (Sim *)p;
    if (!s) return 0;
    SimScratch w;
    cudaGetDevice(&w.device);
    w.n_fft = s->cfg.n_fft; w.cap = s->cap;
    w.d_spec = s->d_spec; w.d_lc = s->d_lc; w.d_coef = s->d_coef; w.d_seg = s->d_seg;
    w.plans.swap(s->plans); w.plans_d2z.swap(s->plans_d2z);
    w.capF = s->capF; w.Qmax = s->tk.Qmax; w.n_obs = s->cfg.n_obs;
    w.d_Y = s->d_Y; w.d_mslot = s->d_mslot; w.d_cnt = s->d_cnt; w.d_coefF = s->d_coefF; w.d_segF = s->d_segF;
    if (s->stream) cudaStreamSynchronize(s->stream);
    bool parked = false;
    if ((w.cap > 0 || w.capF > 0) && pool_limit() > 0) {
        std::lock_guard<std::mutex> lock(g_pool_mu);
        if ((int)g_pool.size() >= pool_limit()) {           // drop the oldest
            int dev = 0; cudaGetDevice(&dev);
            cudaSetDevice(g_pool.front().device); scratch_release(g_pool.front()); cudaSetDevice(dev);
            g_pool.erase(g_pool.begin());
        }
        g_pool.push_back(std::move(w));
        parked = true;
    }
    if (!parked) scratch_release(w);
    cudaFree(s->d_grid); cudaFree(s->d_lo); cudaFree(s->d_hi); cudaFree(s->d_exp); cudaFree(s->d_bkg); cudaFree(s->d_err);
    cudaFree(s->d_tw); cudaFree(s->d_qoff); cudaFree(s->d_flag); cudaFree(s->d_ampt);
    sim_e13_release(s);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
    return 0;
}

int sim_get_plan(Sim *s, std::map<long long, cufftHandle> &cache, cufftType type, int n, int batch, cufftHandle *out)
{
    const long long key = ((long long)n << 20) | batch;
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
    cufftHandle h;
    const int M = n / 2 + 1;
    int nn[1] = {n};
    cufftResult r = (type == CUFFT_Z2D) ? cufftPlanMany(&h, 1, nn, nullptr, 1, M, nullptr, 1, n, CUFFT_Z2D, batch)
                                         : cufftPlanMany(&h, 1, nn, nullptr, 1, n, nullptr, 1, M, CUFFT_D2Z, batch);
    if (r != CUFFT_SUCCESS) { mtg_set_error("cufftPlanMany(n=%d, batch=%d) failed: %d", n, batch, (int)r); return MTG_ENOMEM; }
    cache[key] = h; *out = h;
    return 0;
}

// lightcurves per chunk for the scratch budget (MTG_SIM_MB, default 12 GiB)
static int sim_chunk(const Sim *s, int L)
{
    size_t budget = (size_t)12 << 30;
    if (const char *e = getenv("MTG_SIM_MB")) { long v = atol(e); if (v > 0) budget = (size_t)v << 20; }
    size_t per = (size_t)s->M * 16 + (size_t)s->cfg.n_fft * 8;
    // E13 work arrays: per lightcurve of the chunk on the batched path (short series); long series run in a fixed number of
    // slots (mtg_e13.cu, e13_run_large), whatever the chunk
    if (s->cfg.pdf != MTG_PDF_GAUSSIAN && s->cfg.sim_duration / s->cfg.sim_dt < 32770.0) per += (size_t)s->cfg.n_fft * 8 * 8;
    size_t n = budget / per / 2;                                                  // leave room for cuFFT work areas
    if (n < 1) n = 1;
    return (int)(n < (size_t)L ? n : (size_t)L);
}

// adopt a parked workspace of the same FFT length (and window, for the fused buffers) that is large enough
static bool sim_adopt(Sim *s, int chunk, int chunkF)
{
    if (s->cap != 0 || s->capF != 0) return false;
    std::lock_guard<std::mutex> lock(g_pool_mu);
    int dev = 0; cudaGetDevice(&dev);
    for (size_t i = 0; i < g_pool.size(); ++i) {
        SimScratch &w = g_pool[i];
        if (w.device != dev || w.n_fft != s->cfg.n_fft || w.cap < chunk || w.capF < chunkF) continue;
        if (w.capF > 0 && (w.Qmax != s->tk.Qmax || w.n_obs != s->cfg.n_obs)) continue;
        s->d_spec = w.d_spec; s->d_lc = w.d_lc; s->d_coef = w.d_coef; s->d_seg = w.d_seg; s->cap = w.cap;
        s->plans.swap(w.plans); s->plans_d2z.swap(w.plans_d2z);
        s->d_Y = w.d_Y; s->d_mslot = w.d_mslot; s->d_cnt = w.d_cnt; s->d_coefF = w.d_coefF; s->d_segF = w.d_segF; s->capF = w.capF;
        g_pool.erase(g_pool.begin() + i);
        return true;
    }
    return false;
}

// lightcurves per chunk of the fused TK95 path: Y[l][K2][Qmax] (MTG_SIM_FUSED_MB, default 3 GiB)
static int sim_chunk_fused(const Sim *s, int L)
{
    size_t budget = (size_t)3 << 30;
    if (const char *e = getenv("MTG_SIM_FUSED_MB")) { long v = atol(e); if (v > 0) budget = (size_t)v << 20; }
    const size_t per = (size_t)s->tk.Qmax * s->tk.K2 * sizeof(double2);
    size_t n = budget / per;
    if (n < 1) n = 1;
    return (int)(n < (size_t)L ? n : (size_t)L);
}

static int sim_reserve_fused(Sim *s, int chunk)
{
    if (chunk <= s->capF) return 0;
    if (sim_adopt(s, 0, chunk)) return 0;
    cudaFree(s->d_Y); cudaFree(s->d_mslot); cudaFree(s->d_cnt);
    s->d_Y = nullptr; s->d_mslot = nullptr; s->d_cnt = nullptr; s->capF = 0;
    cudaFree(s->d_coefF); cudaFree(s->d_segF); s->d_coefF = nullptr; s->d_segF = nullptr;
    int rc;
    if ((rc = sim_alloc((void **)&s->d_Y, (size_t)chunk * s->tk.K2 * s->tk.Qmax * sizeof(double2)))) return rc;
    if ((rc = sim_alloc((void **)&s->d_mslot, (size_t)chunk * s->tk.Qmax * sizeof(int)))) return rc;
    if ((rc = sim_alloc((void **)&s->d_cnt, (size_t)chunk * s->cfg.n_obs * sizeof(int)))) return rc;
    if ((rc = sim_alloc((void **)&s->d_coefF, (size_t)chunk * sizeof(PsdCoef)))) return rc;
    if ((rc = sim_alloc((void **)&s->d_segF, (size_t)chunk * sizeof(SegInfo)))) return rc;
    s->capF = chunk;
    return 0;
}

static int sim_reserve(Sim *s, int chunk)
{
    if (chunk <= s->cap) return 0;
    if (sim_adopt(s, chunk, 0)) return 0;
    cudaFree(s->d_spec); cudaFree(s->d_lc); cudaFree(s->d_coef); cudaFree(s->d_seg);
    s->d_spec = nullptr; s->d_lc = nullptr; s->d_coef = nullptr; s->d_seg = nullptr; s->cap = 0;
    int rc;
    if ((rc = sim_alloc((void **)&s->d_spec, (size_t)chunk * s->M * sizeof(double2)))) return rc;
    if ((rc = sim_alloc((void **)&s->d_lc, (size_t)chunk * s->cfg.n_fft * sizeof(double)))) return rc;
    if ((rc = sim_alloc((void **)&s->d_coef, (size_t)chunk * sizeof(PsdCoef)))) return rc;
    if ((rc = sim_alloc((void **)&s->d_seg, (size_t)chunk * sizeof(SegInfo)))) return rc;
    s->cap = chunk;
    return 0;
}

// core: all pointers are device pointers; runs on `stream`; out_tm is [n_obs][L] time-major
int sim_generate_core(Sim *s, int L, const mtg_model_t *model, const double *d_params, const double *d_psd_table,
                      uint64_t seed, const double *d_normals, const double *d_useg, const double *d_pdfvar,
                      double *d_out_tm, SegInfo *d_seg_out, int32_t *d_e13_iters, cudaStream_t stream)
{
    const int n_fft = s->cfg.n_fft, M = s->M, N = s->cfg.n_obs;
    if (!model == !d_psd_table) { mtg_set_error("mtg_sim_generate: give exactly one of model+params / psd_table"); return MTG_EINVAL; }
    Plan pl;
    if (model) {
        int rc = mtg_build_plan(model, &pl);
        if (rc) return rc;
        if (!d_params && pl.n_free > 0) { mtg_set_error("mtg_sim_generate: params missing"); return MTG_EINVAL; }
    }
    if (s->tk.ok && s->cfg.pdf == MTG_PDF_GAUSSIAN) {
        // fused path (mtg_tk95.cu): segment -> needed samples -> column FFTs in shared memory -> row sums + binning
        const int chunkF = sim_chunk_fused(s, L);
        int rc = sim_reserve_fused(s, chunkF);
        if (rc) return rc;
        double sp_val, sp_fold;
        spectrum_scales(n_fft, s->cfg.sim_dt, &sp_val, &sp_fold);
        const double *d_ampt = nullptr;
        if (d_psd_table && !d_normals) {
            if (!s->d_ampt && (rc = sim_alloc((void **)&s->d_ampt, sizeof(double) * (size_t)s->tk.K2 * s->tk.N1))) return rc;
            if ((rc = tk_launch_amp_table(s->tk, sp_fold, d_psd_table, s->d_ampt, stream))) return rc;
            d_ampt = s->d_ampt;
        }
        for (int l0 = 0; l0 < L; l0 += chunkF) {
            const int Lc = L - l0 < chunkF ? L - l0 : chunkF;
            if (model) {
                k_psd_coeffs<<<(Lc + 127) / 128, 128, 0, stream>>>(pl, d_params + (size_t)l0 * pl.n_free, Lc, s->d_coefF);
                mtg_count_launch();
            }
            k_segment<<<(Lc + 127) / 128, 128, 0, stream>>>(Lc, n_fft, s->d_grid, s->cfg.sim_dt, s->cfg.sim_duration,
                                                           s->first_bin_start, d_useg ? d_useg + l0 : nullptr, seed,
                                                           (uint64_t)l0, s->d_segF);
            const long long tot = (long long)Lc * N;
            k_need<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(Lc, N, s->d_grid, s->cfg.sim_dt, s->d_segF, s->d_lo, s->d_hi,
                                                                      s->d_qoff, s->tk.Qmax, s->d_mslot, s->d_cnt, s->d_flag);
            mtg_count_launch(2);
            if ((rc = tk_launch_columns(s->tk, Lc, sp_val, sp_fold, model ? s->d_coefF : nullptr, d_psd_table, d_ampt,
                                        d_normals ? d_normals + (size_t)l0 * 2 * M : nullptr, seed, (uint64_t)l0, s->d_tw,
                                        s->d_mslot, s->d_Y, stream))) return rc;
            if ((rc = tk_launch_rows(s->tk, Lc, N, s->d_qoff, s->d_mslot, s->d_cnt, s->d_Y, s->cfg.mean, d_out_tm, L, l0, stream))) return rc;
            if (d_seg_out)
                MTG_CUDA(cudaMemcpyAsync(d_seg_out + l0, s->d_segF, sizeof(SegInfo) * Lc, cudaMemcpyDeviceToDevice, stream));
            MTG_CUDA(cudaGetLastError());
        }
        return 0;
    }
    const int chunk = sim_chunk(s, L);
    int rc = sim_reserve(s, chunk);
    if (rc) return rc;
    for (int l0 = 0; l0 < L; l0 += chunk) {
        const int Lc = L - l0 < chunk ? L - l0 : chunk;
        if (model) {
            k_psd_coeffs<<<(Lc + 127) / 128, 128, 0, stream>>>(pl, d_params + (size_t)l0 * pl.n_free, Lc, s->d_coef);
            mtg_count_launch();
        }
        double sp_val, sp_fold;
        spectrum_scales(n_fft, s->cfg.sim_dt, &sp_val, &sp_fold);
        k_spectrum<<<dim3((M + 255) / 256, Lc), 256, 0, stream>>>(
            Lc, n_fft, M, sp_val, sp_fold, model ? s->d_coef : nullptr, d_psd_table,
            d_normals ? d_normals + (size_t)l0 * 2 * M : nullptr, seed, (uint64_t)l0, s->tk.N2 > 0 ? s->tk.N2 : 1, s->d_spec);
        mtg_count_launch();
        cufftHandle plan;
        if ((rc = sim_get_plan(s, s->plans, CUFFT_Z2D, n_fft, Lc, &plan))) return rc;
        cufftSetStream(plan, stream);
        cufftResult fr = cufftExecZ2D(plan, (cufftDoubleComplex *)s->d_spec, s->d_lc);
        if (fr != CUFFT_SUCCESS) { mtg_set_error("cufftExecZ2D failed: %d", (int)fr); return MTG_EINVAL; }
        k_segment<<<(Lc + 127) / 128, 128, 0, stream>>>(Lc, n_fft, s->d_grid, s->cfg.sim_dt, s->cfg.sim_duration,
                                                       s->first_bin_start, d_useg ? d_useg + l0 : nullptr, seed,
                                                       (uint64_t)l0, s->d_seg);
        mtg_count_launch();
        double mean_add = s->cfg.mean;
        if (s->cfg.pdf != MTG_PDF_GAUSSIAN) {
            if ((rc = sim_e13_chunk(s, Lc, l0, seed, d_pdfvar, d_e13_iters ? d_e13_iters + l0 : nullptr, stream))) return rc;
            mean_add = 0.0;                    // E13 output already carries the target mean
        }
        const long long tot = (long long)Lc * N;
        k_bin<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(Lc, N, n_fft, s->d_grid, s->d_lc, s->cfg.sim_dt, s->d_seg,
                                                                 s->d_lo, s->d_hi, mean_add, d_out_tm, L, l0);
        mtg_count_launch();
        if (d_seg_out)
            MTG_CUDA(cudaMemcpyAsync(d_seg_out + l0, s->d_seg, sizeof(SegInfo) * Lc, cudaMemcpyDeviceToDevice, stream));
        MTG_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int mtg_sim_generate_dev(void *sim, int L, const mtg_model_t *model, const double *params_dev,
                                    const double *psd_table_dev, uint64_t seed, double *out_rates_tm, void *stream)
{
    Sim *s = (Sim *)sim;
    if (!s || L < 0 || !out_rates_tm) { mtg_set_error("mtg_sim_generate_dev: bad argument"); return MTG_EINVAL; }
    if (L == 0) return 0;
    return sim_generate_core(s, L, model, params_dev, psd_table_dev, seed, nullptr, nullptr, nullptr, out_rates_tm,
                             nullptr, nullptr, (cudaStream_t)stream);
}

__global__ void k_transpose_tm_to_rows(const double *__restrict__ tm, double *__restrict__ rows, int L, int N)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)L * N) return;
    const int l = (int)(idx % L), n = (int)(idx / L);
    rows[(size_t)l * N + n] = tm[(size_t)n * L + l];
}

extern "C" int mtg_sim_generate_host(void *sim, int L, const mtg_model_t *model, const double *params,
                                     const double *psd_table, uint64_t seed, const double *normals,
                                     const double *u_segment, const double *pdf_variates,
                                     double *out_rates, int32_t *out_seg, int32_t *out_e13_iters)
{
    Sim *s = (Sim *)sim;
    if (!s || L < 0 || !out_rates) { mtg_set_error("mtg_sim_generate_host: bad argument"); return MTG_EINVAL; }
    if (L == 0) return 0;
    cudaStream_t st = s->stream;
    const int N = s->cfg.n_obs, M = s->M;
    double *d_par = nullptr, *d_tab = nullptr, *d_nrm = nullptr, *d_u = nullptr, *d_pv = nullptr, *d_tm = nullptr, *d_rows = nullptr;
    SegInfo *d_seg = nullptr; int32_t *d_it = nullptr;
    int rc = 0;
    auto up = [&](double **d, const double *h, size_t n) { if (!rc && h) rc = upload(d, h, n, st); };
    if (model && model->n_free > 0) up(&d_par, params, (size_t)L * model->n_free);
    up(&d_tab, psd_table, (size_t)M);
    up(&d_nrm, normals, (size_t)L * 2 * M);
    up(&d_u, u_segment, (size_t)L);
    up(&d_pv, pdf_variates, (size_t)L * s->cfg.n_fft);
    if (!rc) rc = sim_alloc((void **)&d_tm, sizeof(double) * (size_t)L * N);
    if (!rc) rc = sim_alloc((void **)&d_rows, sizeof(double) * (size_t)L * N);
    if (!rc) rc = sim_alloc((void **)&d_seg, sizeof(SegInfo) * (size_t)L);
    if (!rc) rc = sim_alloc((void **)&d_it, sizeof(int32_t) * (size_t)L);
    if (!rc) rc = mtg_cuda_check(cudaMemsetAsync(d_it, 0, sizeof(int32_t) * (size_t)L, st), "memset");
    if (!rc) rc = sim_generate_core(s, L, model, d_par, d_tab, seed, d_nrm, d_u, d_pv, d_tm, d_seg, d_it, st);
    if (!rc) {
        const long long tot = (long long)L * N;
        k_transpose_tm_to_rows<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_tm, d_rows, L, N);
        mtg_count_launch();
        rc = mtg_cuda_check(cudaMemcpyAsync(out_rates, d_rows, sizeof(double) * (size_t)L * N, cudaMemcpyDeviceToHost, st), "d2h");
    }
    std::vector<SegInfo> hseg;
    if (!rc && out_seg) { hseg.resize(L); rc = mtg_cuda_check(cudaMemcpyAsync(hseg.data(), d_seg, sizeof(SegInfo) * L, cudaMemcpyDeviceToHost, st), "d2h seg"); }
    if (!rc && out_e13_iters) rc = mtg_cuda_check(cudaMemcpyAsync(out_e13_iters, d_it, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, st), "d2h it");
    int flag = 0;
    if (!rc && s->d_flag) rc = mtg_cuda_check(cudaMemcpyAsync(&flag, s->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st), "d2h flag");
    if (!rc) rc = mtg_cuda_check(cudaStreamSynchronize(st), "sim generate sync");
    else cudaStreamSynchronize(st);
    if (!rc && flag) { mtg_set_error("fused TK95 path: an exposure bin holds more grid points than the slots reserved for it"); rc = MTG_EINVAL; }
    if (!rc && out_seg) for (int l = 0; l < L; ++l) { out_seg[2 * l] = hseg[l].i0; out_seg[2 * l + 1] = hseg[l].i1 - hseg[l].i0; }
    cudaFree(d_par); cudaFree(d_tab); cudaFree(d_nrm); cudaFree(d_u); cudaFree(d_pv); cudaFree(d_tm); cudaFree(d_rows);
    cudaFree(d_seg); cudaFree(d_it);
    return rc;
}

static int noise_launch(Sim *s, int L, size_t ld_n, size_t ld_l, const double *rates, uint64_t seed,
                        const long long *counts_in, const double *normals_in, double *out_rates, double *out_dy,
                        int32_t *status, cudaStream_t st)
{
    const long long tot = (long long)L * s->cfg.n_obs;
    k_noise<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(s->cfg.noise, L, s->cfg.n_obs, ld_n, ld_l, rates, s->d_exp, s->d_bkg,
                                                           s->d_err, s->cfg.sigma_noise, s->cfg.kraft_counts, seed,
                                                           counts_in, normals_in, out_rates, out_dy, status);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k_noise launch");
}

extern "C" int mtg_sim_add_noise_dev(void *sim, int L, const double *rates_tm, uint64_t seed,
                                     double *out_rates_tm, double *out_dy_tm, void *stream)
{
    Sim *s = (Sim *)sim;
    if (!s || L < 0 || !rates_tm || !out_rates_tm || !out_dy_tm) { mtg_set_error("mtg_sim_add_noise_dev: bad argument"); return MTG_EINVAL; }
    if (L == 0) return 0;
    return noise_launch(s, L, (size_t)L, 1, rates_tm, seed, nullptr, nullptr, out_rates_tm, out_dy_tm, nullptr, (cudaStream_t)stream);
}

extern "C" int mtg_sim_add_noise_host(void *sim, int L, const double *rates, uint64_t seed, const int64_t *counts_in,
                                      const double *normals_in, double *out_rates, double *out_dy, int32_t *status)
{
    Sim *s = (Sim *)sim;
    if (!s || L < 0 || !rates || !out_rates || !out_dy) { mtg_set_error("mtg_sim_add_noise_host: bad argument"); return MTG_EINVAL; }
    if (L == 0) return 0;
    cudaStream_t st = s->stream;
    const size_t n = (size_t)L * s->cfg.n_obs;
    double *d_r = nullptr, *d_nrm = nullptr, *d_o = nullptr, *d_dy = nullptr; long long *d_c = nullptr; int32_t *d_st = nullptr;
    int rc = upload(&d_r, rates, n, st);
    if (!rc && normals_in) rc = upload(&d_nrm, normals_in, n, st);
    if (!rc && counts_in) {
        rc = sim_alloc((void **)&d_c, n * sizeof(long long));
        if (!rc) rc = mtg_cuda_check(cudaMemcpyAsync(d_c, counts_in, n * sizeof(long long), cudaMemcpyHostToDevice, st), "h2d counts");
    }
    if (!rc) rc = sim_alloc((void **)&d_o, n * sizeof(double));
    if (!rc) rc = sim_alloc((void **)&d_dy, n * sizeof(double));
    if (!rc) rc = sim_alloc((void **)&d_st, n * sizeof(int32_t));
    // host layout is row-major [L][n_obs]: element (n, l) at l*n_obs + n
    if (!rc) rc = noise_launch(s, L, 1, (size_t)s->cfg.n_obs, d_r, seed, d_c, d_nrm, d_o, d_dy, d_st, st);
    if (!rc) rc = mtg_cuda_check(cudaMemcpyAsync(out_rates, d_o, n * sizeof(double), cudaMemcpyDeviceToHost, st), "d2h");
    if (!rc) rc = mtg_cuda_check(cudaMemcpyAsync(out_dy, d_dy, n * sizeof(double), cudaMemcpyDeviceToHost, st), "d2h");
    if (!rc && status) rc = mtg_cuda_check(cudaMemcpyAsync(status, d_st, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "d2h");
    if (!rc) rc = mtg_cuda_check(cudaStreamSynchronize(st), "noise sync"); else cudaStreamSynchronize(st);
    cudaFree(d_r); cudaFree(d_nrm); cudaFree(d_c); cudaFree(d_o); cudaFree(d_dy); cudaFree(d_st);
    return rc;
}

extern "C" int mtg_downsample_host(int n_seg, const double *time, const double *countrate, int n_obs,
                                   const double *bin_lo, const double *bin_hi, double *out_rates,
                                   int32_t *out_first, int32_t *out_count)
{
    if (n_seg < 0 || n_obs < 0 || !time || !countrate || !bin_lo || !bin_hi || !out_rates) {
        mtg_set_error("mtg_downsample_host: bad argument"); return MTG_EINVAL;
    }
    if (n_obs == 0) return 0;
    if (n_seg == 0) {
        for (int n = 0; n < n_obs; ++n) { out_rates[n] = nan(""); if (out_first) out_first[n] = 0; if (out_count) out_count[n] = 0; }
        return 0;
    }
    double *d_t = nullptr, *d_r = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_o = nullptr; int32_t *d_f = nullptr, *d_c = nullptr;
    cudaStream_t st = 0;
    int rc = upload(&d_t, time, n_seg > 0 ? n_seg : 1, st);
    if (!rc) rc = upload(&d_r, countrate, n_seg > 0 ? n_seg : 1, st);
    if (!rc) rc = upload(&d_lo, bin_lo, n_obs, st);
    if (!rc) rc = upload(&d_hi, bin_hi, n_obs, st);
    if (!rc) rc = sim_alloc((void **)&d_o, sizeof(double) * n_obs);
    if (!rc) rc = sim_alloc((void **)&d_f, sizeof(int32_t) * n_obs);
    if (!rc) rc = sim_alloc((void **)&d_c, sizeof(int32_t) * n_obs);
    if (!rc) {
        k_downsample_explicit<<<(n_obs + 127) / 128, 128, 0, st>>>(n_seg, d_t, d_r, n_obs, d_lo, d_hi, d_o, d_f, d_c);
        mtg_count_launch();
        rc = mtg_cuda_check(cudaGetLastError(), "k_downsample launch");
    }
    if (!rc) rc = mtg_cuda_check(cudaMemcpy(out_rates, d_o, sizeof(double) * n_obs, cudaMemcpyDeviceToHost), "d2h");
    if (!rc && out_first) rc = mtg_cuda_check(cudaMemcpy(out_first, d_f, sizeof(int32_t) * n_obs, cudaMemcpyDeviceToHost), "d2h");
    if (!rc && out_count) rc = mtg_cuda_check(cudaMemcpy(out_count, d_c, sizeof(int32_t) * n_obs, cudaMemcpyDeviceToHost), "d2h");
    cudaFree(d_t); cudaFree(d_r); cudaFree(d_lo); cudaFree(d_hi); cudaFree(d_o); cudaFree(d_f); cudaFree(d_c);
    return rc;
}

// ---- standalone noise on explicit arrays (PoissonNoise / KraftNoise / GaussianNoise.add_noise) -------
extern "C" int mtg_noise_host(int kind, int L, int n_obs, const double *rates, const double *exposures,
                              const double *bkg_counts, const double *bkg_rate_err, double sigma, double kraft_counts,
                              uint64_t seed, const int64_t *counts_in, const double *normals_in,
                              double *out_rates, double *out_dy, int32_t *status)
{
    if (kind < MTG_NOISE_POISSON || kind > MTG_NOISE_GAUSSIAN || L < 0 || n_obs < 1 || !rates || !out_rates || !out_dy ||
        (kind != MTG_NOISE_GAUSSIAN && !exposures)) { mtg_set_error("mtg_noise_host: bad argument"); return MTG_EINVAL; }
    if (L == 0) return 0;
    Sim s;
    s.cfg.noise = kind; s.cfg.n_obs = n_obs; s.cfg.sigma_noise = sigma; s.cfg.kraft_counts = kraft_counts;
    int rc = mtg_cuda_check(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "stream");
    if (!rc && exposures) rc = upload(&s.d_exp, exposures, n_obs, s.stream);
    if (!rc && bkg_counts) rc = upload(&s.d_bkg, bkg_counts, n_obs, s.stream);
    if (!rc && bkg_rate_err) rc = upload(&s.d_err, bkg_rate_err, n_obs, s.stream);
    if (!rc) rc = mtg_sim_add_noise_host(&s, L, rates, seed, counts_in, normals_in, out_rates, out_dy, status);
    cudaFree(s.d_exp); cudaFree(s.d_bkg); cudaFree(s.d_err);
    if (s.stream) cudaStreamDestroy(s.stream);
    return rc;
}

// ---- the regularly sampled TK95 lightcurve itself (Simulator.simulate_regularly_sampled) -------------
__global__ void k_add_const(double *__restrict__ x, size_t n, double c)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] += c;
}

extern "C" int mtg_sim_regular_host(void *sim, int L, const mtg_model_t *model, const double *params,
                                    const double *psd_table, uint64_t seed, const double *normals, double *out_countrate)
{
    Sim *s = (Sim *)sim;
    if (!s || L < 0 || !out_countrate) { mtg_set_error("mtg_sim_regular_host: bad argument"); return MTG_EINVAL; }
    if (!model == !psd_table) { mtg_set_error("mtg_sim_regular_host: give exactly one of model+params / psd_table"); return MTG_EINVAL; }
    if (L == 0) return 0;
    cudaStream_t st = s->stream;
    const int n_fft = s->cfg.n_fft, M = s->M;
    Plan pl;
    int rc = 0;
    if (model) rc = mtg_build_plan(model, &pl);
    double *d_par = nullptr, *d_tab = nullptr, *d_nrm = nullptr;
    if (!rc && model && pl.n_free > 0) rc = upload(&d_par, params, (size_t)L * pl.n_free, st);
    if (!rc && psd_table) rc = upload(&d_tab, psd_table, (size_t)M, st);
    const int chunk = sim_chunk(s, L);
    if (!rc) rc = sim_reserve(s, chunk);
    for (int l0 = 0; l0 < L && !rc; l0 += chunk) {
        const int Lc = L - l0 < chunk ? L - l0 : chunk;
        if (normals) { cudaFree(d_nrm); d_nrm = nullptr; rc = upload(&d_nrm, normals + (size_t)l0 * 2 * M, (size_t)Lc * 2 * M, st); if (rc) break; }
        if (model) { k_psd_coeffs<<<(Lc + 127) / 128, 128, 0, st>>>(pl, d_par + (size_t)l0 * pl.n_free, Lc, s->d_coef); mtg_count_launch(); }
        double sp_val, sp_fold;
        spectrum_scales(n_fft, s->cfg.sim_dt, &sp_val, &sp_fold);
        k_spectrum<<<dim3((M + 255) / 256, Lc), 256, 0, st>>>(Lc, n_fft, M, sp_val, sp_fold, model ? s->d_coef : nullptr, d_tab,
                                                              d_nrm, seed, (uint64_t)l0, s->tk.N2 > 0 ? s->tk.N2 : 1, s->d_spec);
        cufftHandle plan;
        if ((rc = sim_get_plan(s, s->plans, CUFFT_Z2D, n_fft, Lc, &plan))) break;
        cufftSetStream(plan, st);
        if (cufftExecZ2D(plan, (cufftDoubleComplex *)s->d_spec, s->d_lc) != CUFFT_SUCCESS) { mtg_set_error("cufftExecZ2D failed"); rc = MTG_EINVAL; break; }
        k_add_const<<<148 * 8, 256, 0, st>>>(s->d_lc, (size_t)Lc * n_fft, s->cfg.mean);
        mtg_count_launch(2);
        rc = mtg_cuda_check(cudaMemcpyAsync(out_countrate + (size_t)l0 * n_fft, s->d_lc, sizeof(double) * (size_t)Lc * n_fft,
                                            cudaMemcpyDeviceToHost, st), "d2h");
        if (!rc) rc = mtg_cuda_check(cudaStreamSynchronize(st), "regular sync");
    }
    cudaFree(d_par); cudaFree(d_tab); cudaFree(d_nrm);
    return rc;
}
