// mtg_periodogram.cu -- batched periodograms of the simulated ensembles.
//
// The reference validates its simulator, and studies the bias of Lomb-Scargle slopes, by looping over hundreds of
// lightcurves that share their timestamps (tests/simulator_test.py:24-35 `power_spectrum`;
// docs/notebooks/lomb_scargle_biases.ipynb: `LombScargle(t, rate, fit_mean=True, center_data=True,
// normalization=...).power(frequencies)` per lightcurve, then `fitting.fit_lomb_scargle`).  Here the whole ensemble is
// one call:
//   * mtg_power_spectrum_host : |rfft(rate)|^2 of L regularly sampled lightcurves (cuFFT D2Z batched + one kernel);
//   * mtg_lombscargle_host    : the generalised (floating-mean) Lomb-Scargle periodogram of L lightcurves on shared,
//     uneven times at M frequencies.  Everything that depends on the window only (tau, sum cos^2, sum sin^2, ...) is
//     computed once per frequency (k_ls_window), the phases cos / sin(omega_k (t_n - tau_k)) are tabulated as two
//     [N][M] matrices (k_ls_trig), and the data-dependent sums YC = Y COS, YS = Y SIN are a pair of FP64 matrix
//     products done by a register-tiled kernel (k_ls_power: 64 x 64 output tile per CTA, 4 x 4 per thread for both
//     products, operands staged through shared memory) that also applies the normalisation.
//     Formulas: astropy.timeseries LombScargle, `lombscargle_slow` (Zechmeister & Kuerster 2009; VanderPlas 2018) with
//     unit weights -- restated in oracle/periodogram.py and pinned there on scipy.signal.lombscargle(floating_mean=...).
#include "mtg_common.cuh"
#include <cufft.h>
#include <mutex>

#define LS_TILE 64
#define LS_KT 16

// per-lightcurve centring: y <- y - mean(y) (when asked), ybar = mean(y) after that, yy = mean(y^2)
__global__ void k_ls_prepare(int N, int L, int center, double *__restrict__ y_tm /*[N][L]*/, double *__restrict__ ybar,
                             double *__restrict__ yy)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    double s = 0.0;
    for (int n = 0; n < N; ++n) s += y_tm[(size_t)n * L + l];
    const double m = s / N;
    double s1 = 0.0, s2 = 0.0;
    for (int n = 0; n < N; ++n) {
        double v = y_tm[(size_t)n * L + l];
        if (center) { v -= m; y_tm[(size_t)n * L + l] = v; }
        s1 += v; s2 += v * v;
    }
    ybar[l] = s1 / N; yy[l] = s2 / N;
}

// window-only quantities of frequency k: omega tau, and (with w = 1/N) CC, SS, Csum, Ssum at the shifted phases
__global__ void k_ls_window(int N, int M, const double *__restrict__ t, const double *__restrict__ freq, int fit_mean,
                            double *__restrict__ wtau, double *__restrict__ win /*[4][M]: CC SS Csum Ssum*/)
{
    __shared__ double red[4][32];
    const int k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const double omega = 2.0 * M_PI * freq[k];
    auto block_sum4 = [&](double &a, double &b, double &c, double &d) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o);
            c += __shfl_xor_sync(0xffffffffu, c, o); d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        __syncthreads();
        if (lane == 0) { red[0][wid] = a; red[1][wid] = b; red[2][wid] = c; red[3][wid] = d; }
        __syncthreads();
        a = b = c = d = 0.0;
        for (int w = 0; w < nw; ++w) { a += red[0][w]; b += red[1][w]; c += red[2][w]; d += red[3][w]; }
    };
    const double w = 1.0 / N;
    double S = 0.0, C = 0.0, S2 = 0.0, C2 = 0.0;
    for (int n = tid; n < N; n += blockDim.x) {
        double s, c;
        sincos(omega * t[n], &s, &c);
        S += w * s; C += w * c;
        S2 += w * 2.0 * s * c; C2 += w * (c * c - s * s);          // sin / cos of 2 omega t
    }
    block_sum4(S, C, S2, C2);
    if (fit_mean) { S2 -= 2.0 * S * C; C2 -= (C * C - S * S); }
    const double wt0 = 0.5 * atan2(S2, C2);                        // omega tau
    double CC = 0.0, SS = 0.0, Cs = 0.0, Ss = 0.0;
    for (int n = tid; n < N; n += blockDim.x) {
        double s, c;
        sincos(omega * t[n] - wt0, &s, &c);
        CC += w * c * c; SS += w * s * s; Cs += w * c; Ss += w * s;
    }
    block_sum4(CC, SS, Cs, Ss);
    if (fit_mean) { CC -= Cs * Cs; SS -= Ss * Ss; }
    if (tid == 0) { wtau[k] = wt0; win[k] = CC; win[M + k] = SS; win[2 * M + k] = Cs; win[3 * M + k] = Ss; }
}

__global__ void k_ls_trig(int N, int M, const double *__restrict__ t, const double *__restrict__ freq,
                          const double *__restrict__ wtau, double *__restrict__ cosm, double *__restrict__ sinm /*[N][M]*/)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
    if (k >= M) return;
    double s, c;
    sincos(2.0 * M_PI * freq[k] * t[n] - wtau[k], &s, &c);
    cosm[(size_t)n * M + k] = c; sinm[(size_t)n * M + k] = s;
}

// power[l][k] from YC = sum_n y[n][l] cos[n][k] / N and YS likewise: 64 x 64 tile per CTA, 4 x 4 per thread
__global__ void __launch_bounds__(256)
k_ls_power(int N, int L, int M, const double *__restrict__ y_tm, const double *__restrict__ cosm, const double *__restrict__ sinm,
           const double *__restrict__ win, const double *__restrict__ ybar, const double *__restrict__ yy,
           int fit_mean, int normalization, double *__restrict__ power /*[L][M]*/)
{
    __shared__ double ys[LS_KT][LS_TILE], cs[LS_KT][LS_TILE], ss[LS_KT][LS_TILE];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;            // tx: frequency direction, ty: lightcurve direction
    const int k0 = blockIdx.x * LS_TILE, l0 = blockIdx.y * LS_TILE;
    double yc[4][4], ysn[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { yc[i][j] = 0.0; ysn[i][j] = 0.0; }
    for (int n0 = 0; n0 < N; n0 += LS_KT) {
        for (int idx = tid; idx < LS_KT * LS_TILE; idx += 256) {
            const int r = idx / LS_TILE, c = idx % LS_TILE, n = n0 + r;
            ys[r][c] = (n < N && l0 + c < L) ? y_tm[(size_t)n * L + l0 + c] : 0.0;
            cs[r][c] = (n < N && k0 + c < M) ? cosm[(size_t)n * M + k0 + c] : 0.0;
            ss[r][c] = (n < N && k0 + c < M) ? sinm[(size_t)n * M + k0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < LS_KT; ++r) {
            double yv[4], cv[4], sv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { yv[i] = ys[r][ty + 16 * i]; cv[i] = cs[r][tx + 16 * i]; sv[i] = ss[r][tx + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { yc[i][j] = fma(yv[i], cv[j], yc[i][j]); ysn[i][j] = fma(yv[i], sv[j], ysn[i][j]); }
        }
        __syncthreads();
    }
    const double w = 1.0 / N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int l = l0 + ty + 16 * i;
        if (l >= L) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx + 16 * j;
            if (k >= M) continue;
            double YC = yc[i][j] * w, YS = ysn[i][j] * w;
            if (fit_mean) { YC -= ybar[l] * win[2 * M + k]; YS -= ybar[l] * win[3 * M + k]; }
            double p = YC * YC / win[k] + YS * YS / win[M + k];
            const double YY = yy[l];
            if (normalization == MTG_LS_STANDARD) p /= YY;
            else if (normalization == MTG_LS_MODEL) p /= (YY - p);
            else if (normalization == MTG_LS_LOG) p = -log(1.0 - p / YY);
            else p *= 0.5 * (double)N;                                  // psd: 0.5 * sum(dy^-2) with unit errors
            power[(size_t)l * M + k] = p;
        }
    }
}

int mtg_transpose_sn_to_ns(const double *in_sn, double *out_ns, int S, int N, cudaStream_t stream);

extern "C" int mtg_lombscargle_host(int N, int L, int M, const double *t, const double *y_rows, const double *freqs,
                                    int fit_mean, int center_data, int normalization, double *out_power)
{
    if (N < 2 || L < 1 || M < 1 || !t || !y_rows || !freqs || !out_power || normalization < MTG_LS_STANDARD || normalization > MTG_LS_PSD) {
        mtg_set_error("mtg_lombscargle_host: bad argument"); return MTG_EINVAL;
    }
    cudaStream_t st = nullptr;
    MTG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc = mtg_pool_setup();
    double *d = nullptr;
    const size_t n_y = (size_t)N * L, n_trig = (size_t)N * M, n_out = (size_t)L * M;
    const size_t total = N + M + 2 * n_y + 2 * n_trig + 5 * (size_t)M + 2 * (size_t)L + n_out;
    if (!rc) rc = mtg_cuda_check(cudaMallocAsync((void **)&d, total * sizeof(double), st), "cudaMallocAsync");
    if (rc) { cudaStreamDestroy(st); return rc; }
    double *d_t = d, *d_f = d_t + N, *d_rows = d_f + M, *d_tm = d_rows + n_y, *d_cos = d_tm + n_y, *d_sin = d_cos + n_trig,
           *d_wtau = d_sin + n_trig, *d_win = d_wtau + M, *d_ybar = d_win + 4 * (size_t)M, *d_yy = d_ybar + L, *d_out = d_yy + L;
    cudaMemcpyAsync(d_t, t, sizeof(double) * N, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_f, freqs, sizeof(double) * M, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_rows, y_rows, sizeof(double) * n_y, cudaMemcpyHostToDevice, st);
    rc = mtg_transpose_sn_to_ns(d_rows, d_tm, L, N, st);
    if (!rc) {
        k_ls_prepare<<<(L + 127) / 128, 128, 0, st>>>(N, L, (center_data || fit_mean) ? 1 : 0, d_tm, d_ybar, d_yy);
        k_ls_window<<<M, 256, 0, st>>>(N, M, d_t, d_f, fit_mean, d_wtau, d_win);
        k_ls_trig<<<dim3((M + 127) / 128, N), 128, 0, st>>>(N, M, d_t, d_f, d_wtau, d_cos, d_sin);
        k_ls_power<<<dim3((M + LS_TILE - 1) / LS_TILE, (L + LS_TILE - 1) / LS_TILE), 256, 0, st>>>(
            N, L, M, d_tm, d_cos, d_sin, d_win, d_ybar, d_yy, fit_mean, normalization, d_out);
        mtg_count_launch(4);
        rc = mtg_cuda_check(cudaGetLastError(), "lombscargle launch");
    }
    if (!rc) {
        cudaMemcpyAsync(out_power, d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, st);
        rc = mtg_cuda_check(cudaStreamSynchronize(st), "mtg_lombscargle_host sync");
    }
    cudaFreeAsync(d, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

// ---- |rfft|^2 of regularly sampled lightcurves ------------------------------------------------------
__global__ void k_abs2(size_t n, const double2 *__restrict__ z, double *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double2 v = z[i]; out[i] = v.x * v.x + v.y * v.y; }
}

extern "C" int mtg_power_spectrum_host(int L, int n, const double *rates_rows, double *out_power)
{
    if (L < 1 || n < 2 || !rates_rows || !out_power) { mtg_set_error("mtg_power_spectrum_host: bad argument"); return MTG_EINVAL; }
    const int Mh = n / 2 + 1;
    cudaStream_t st = nullptr;
    MTG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    double *d_in = nullptr; double2 *d_spec = nullptr; double *d_out = nullptr;
    cufftHandle plan = 0;
    int rc = 0;
    auto fail = [&](int code) { rc = code; };
    if (cudaMalloc((void **)&d_in, sizeof(double) * (size_t)L * n) != cudaSuccess ||
        cudaMalloc((void **)&d_spec, sizeof(double2) * (size_t)L * Mh) != cudaSuccess ||
        cudaMalloc((void **)&d_out, sizeof(double) * (size_t)L * Mh) != cudaSuccess) { mtg_set_error("mtg_power_spectrum_host: out of memory"); fail(MTG_ENOMEM); }
    int nn[1] = {n};
    if (!rc && cufftPlanMany(&plan, 1, nn, nullptr, 1, n, nullptr, 1, Mh, CUFFT_D2Z, L) != CUFFT_SUCCESS) {
        mtg_set_error("mtg_power_spectrum_host: cufftPlanMany(n=%d, batch=%d) failed", n, L); plan = 0; fail(MTG_ENOMEM);
    }
    if (!rc) {
        cufftSetStream(plan, st);
        cudaMemcpyAsync(d_in, rates_rows, sizeof(double) * (size_t)L * n, cudaMemcpyHostToDevice, st);
        if (cufftExecD2Z(plan, d_in, (cufftDoubleComplex *)d_spec) != CUFFT_SUCCESS) { mtg_set_error("cufftExecD2Z failed"); fail(MTG_EINVAL); }
    }
    if (!rc) {
        const size_t tot = (size_t)L * Mh;
        k_abs2<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(tot, d_spec, d_out);
        mtg_count_launch();
        cudaMemcpyAsync(out_power, d_out, sizeof(double) * tot, cudaMemcpyDeviceToHost, st);
        rc = mtg_cuda_check(cudaStreamSynchronize(st), "mtg_power_spectrum_host sync");
    }
    if (plan) cufftDestroy(plan);
    cudaFree(d_in); cudaFree(d_spec); cudaFree(d_out);
    cudaStreamDestroy(st);
    return rc;
}

// ---- autocorrelation functions of chains (emcee autocorr.function_1d, batched) ------------------------
// GPModelling.derive_posteriors checks convergence every `convergence_steps` steps on the integrated autocorrelation
// time of the WHOLE chain so far (gpmodelling.py:251-263 -> emcee get_autocorr_time): walkers x ndim FFTs of the chain
// length.  With the sampler itself at ~25 000 steps/s that host computation became the bottleneck of the call, so the
// device sampler computes the normalised autocorrelation functions here: mean removal + zero padding to 2 * 2^k, D2Z,
// |.|^2, Z2D, normalisation by lag 0 -- emcee's function_1d for all series at once.
__global__ void k_acf_center_pad(int n_t, int nfft, const double *__restrict__ x /*[L][n_t]*/, double *__restrict__ out /*[L][nfft]*/)
{
    __shared__ double red[32];
    const int l = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const double *row = x + (size_t)l * n_t;
    double s = 0.0;
    for (int i = tid; i < n_t; i += blockDim.x) s += row[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[wid] = s;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < nw; ++w) tot += red[w];
    const double m = tot / n_t;
    double *o = out + (size_t)l * nfft;
    for (int i = tid; i < nfft; i += blockDim.x) o[i] = i < n_t ? row[i] - m : 0.0;
}
__global__ void k_acf_power(size_t n, double2 *__restrict__ z)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const double2 v = z[i]; z[i] = make_double2(v.x * v.x + v.y * v.y, 0.0); }
}
__global__ void k_acf_normalise(int n_t, int nfft, const double *__restrict__ r /*[L][nfft]*/, double *__restrict__ out /*[L][n_t]*/)
{
    const int l = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_t) out[(size_t)l * n_t + i] = r[(size_t)l * nfft + i] / r[(size_t)l * nfft];
}

// Plans and buffers of the autocorrelation entry are kept between calls (derive_posteriors calls it every few hundred
// steps with a slowly growing chain: the padded length changes only at powers of two): plan creation and cudaMalloc cost
// more than the transforms themselves.
struct AcfCtx {
    std::mutex mu;
    int device = -1;
    cudaStream_t st = nullptr;
    int nfft = 0, L = 0;
    cufftHandle fwd = 0, inv = 0;
    double *d_x = nullptr, *d_pad = nullptr, *d_out = nullptr; double2 *d_spec = nullptr;
    size_t cap_x = 0, cap_out = 0, cap_pad = 0, cap_spec = 0;
    void drop_plans() { if (fwd) cufftDestroy(fwd); if (inv) cufftDestroy(inv); fwd = inv = 0; nfft = L = 0; }
    void drop_all()
    {
        drop_plans();
        cudaFree(d_x); cudaFree(d_pad); cudaFree(d_out); cudaFree(d_spec);
        d_x = d_pad = d_out = nullptr; d_spec = nullptr; cap_x = cap_out = cap_pad = cap_spec = 0;
        if (st) cudaStreamDestroy(st);
        st = nullptr;
    }
};
static AcfCtx g_acf;

static int acf_grow(void **p, size_t *cap, size_t bytes)
{
    if (bytes <= *cap) return 0;
    cudaFree(*p); *p = nullptr; *cap = 0;
    const size_t want = bytes + bytes / 2;
    if (cudaMalloc(p, want) != cudaSuccess) { mtg_set_error("mtg_autocorr_host: out of memory (%zu bytes)", want); return MTG_ENOMEM; }
    *cap = want;
    return 0;
}

extern "C" int mtg_autocorr_host(int L, int n_t, const double *x_rows, double *out_acf)
{
    if (L < 1 || n_t < 1 || !x_rows || !out_acf) { mtg_set_error("mtg_autocorr_host: bad argument"); return MTG_EINVAL; }
    int nfft = 2;
    while (nfft < 2 * n_t) nfft <<= 1;                     // 2 * next power of two >= n_t, as emcee pads
    const int Mh = nfft / 2 + 1;
    AcfCtx &c = g_acf;
    std::lock_guard<std::mutex> lock(c.mu);
    int dev = 0;
    MTG_CUDA(cudaGetDevice(&dev));
    if (dev != c.device) { c.drop_all(); c.device = dev; }
    if (!c.st) MTG_CUDA(cudaStreamCreateWithFlags(&c.st, cudaStreamNonBlocking));
    int rc = 0;
    if ((rc = acf_grow((void **)&c.d_out, &c.cap_out, sizeof(double) * (size_t)L * n_t)) ||
        (rc = acf_grow((void **)&c.d_x, &c.cap_x, sizeof(double) * (size_t)L * n_t)) ||
        (rc = acf_grow((void **)&c.d_pad, &c.cap_pad, sizeof(double) * (size_t)L * nfft)) ||
        (rc = acf_grow((void **)&c.d_spec, &c.cap_spec, sizeof(double2) * (size_t)L * Mh))) return rc;
    if (nfft != c.nfft || L != c.L) {
        c.drop_plans();
        int nn[1] = {nfft};
        if (cufftPlanMany(&c.fwd, 1, nn, nullptr, 1, nfft, nullptr, 1, Mh, CUFFT_D2Z, L) != CUFFT_SUCCESS ||
            cufftPlanMany(&c.inv, 1, nn, nullptr, 1, Mh, nullptr, 1, nfft, CUFFT_Z2D, L) != CUFFT_SUCCESS) {
            c.drop_plans();
            mtg_set_error("mtg_autocorr_host: cufftPlanMany(n=%d, batch=%d) failed", nfft, L); return MTG_ENOMEM;
        }
        cufftSetStream(c.fwd, c.st); cufftSetStream(c.inv, c.st);
        c.nfft = nfft; c.L = L;
    }
    cudaStream_t st = c.st;
    cudaMemcpyAsync(c.d_x, x_rows, sizeof(double) * (size_t)L * n_t, cudaMemcpyHostToDevice, st);
    k_acf_center_pad<<<L, 256, 0, st>>>(n_t, nfft, c.d_x, c.d_pad);
    if (cufftExecD2Z(c.fwd, c.d_pad, (cufftDoubleComplex *)c.d_spec) != CUFFT_SUCCESS) { mtg_set_error("cufftExecD2Z failed"); return MTG_EINVAL; }
    const size_t tot = (size_t)L * Mh;
    k_acf_power<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(tot, c.d_spec);
    if (cufftExecZ2D(c.inv, (cufftDoubleComplex *)c.d_spec, c.d_pad) != CUFFT_SUCCESS) { mtg_set_error("cufftExecZ2D failed"); return MTG_EINVAL; }
    k_acf_normalise<<<dim3((n_t + 255) / 256, L), 256, 0, st>>>(n_t, nfft, c.d_pad, c.d_out);
    mtg_count_launch(3);
    cudaMemcpyAsync(out_acf, c.d_out, sizeof(double) * (size_t)L * n_t, cudaMemcpyDeviceToHost, st);
    return mtg_cuda_check(cudaStreamSynchronize(st), "mtg_autocorr_host sync");
}
