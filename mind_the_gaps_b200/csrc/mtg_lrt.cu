// mtg_lrt.cu -- small device helpers around the hot kernels: data preparation for the LRT pipeline,
// best-of-group reduction over multi-starts, and GP.predict at the training times.
#include "mtg_recursion.cuh"

// diag = (dy + 1e-12)^2 (gpmodelling.py:54: yerr = dy + 1e-12, squared by celerite) and the
// per-lightcurve mean of y (the frozen ConstantModel of gpmodelling.py:83-87).  Time-major [N][S].
__global__ void k_prepare_data(int N, int S, const double *__restrict__ y_tm, const double *__restrict__ dy_tm,
                               double *__restrict__ diag_tm, double *__restrict__ mean_out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double acc = 0.0;
    for (int n = 0; n < N; ++n) {
        const size_t at = (size_t)n * S + s;
        const double e = dy_tm[at] + 1e-12;
        diag_tm[at] = e * e;
        acc += y_tm[at];
    }
    if (mean_out) mean_out[s] = acc / (double)N;
}

extern "C" int mtg_prepare_data_dev(int N, int S, const double *y_tm, const double *dy_tm, double *diag_tm,
                                    double *mean_out, void *stream)
{
    if (N < 1 || S < 0 || !y_tm || !dy_tm || !diag_tm) { mtg_set_error("mtg_prepare_data_dev: bad argument"); return MTG_EINVAL; }
    if (S == 0) return 0;
    k_prepare_data<<<(S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(N, S, y_tm, dy_tm, diag_tm, mean_out);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k_prepare_data launch");
}

// best (smallest finite) objective of every group of `group` consecutive problems
__global__ void k_group_best(int S, int group, const double *__restrict__ f, double *__restrict__ best, int32_t *__restrict__ arg)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    double b = INFINITY; int a = 0;
    for (int k = 0; k < group; ++k) {
        const double v = f[(size_t)s * group + k];
        if (isfinite(v) && v < b) { b = v; a = k; }
    }
    best[s] = b;
    if (arg) arg[s] = s * group + a;
}

extern "C" int mtg_group_best_dev(int S, int group, const double *f, double *best, int32_t *arg, void *stream)
{
    if (S < 0 || group < 1 || !f || !best) { mtg_set_error("mtg_group_best_dev: bad argument"); return MTG_EINVAL; }
    if (S == 0) return 0;
    k_group_best<<<(S + 127) / 128, 128, 0, (cudaStream_t)stream>>>(S, group, f, best, arg);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k_group_best launch");
}

// GP.predict(y, return_var=True) at the training times (gpmodelling.py:366):
//   mu_n  = mean(t_n) + r_n - Sigma_n alpha_n,   var_n = Sigma_n - Sigma_n^2 (K^-1)_nn,
// Sigma_n = diag_n + jitter, alpha = K^-1 r.  alpha and diag(K^-1) come out of the reverse sweep of K2.
__global__ void k_predict_finish(const __grid_constant__ Plan pl, int N, const double *__restrict__ t, const double *__restrict__ y,
                                 const double *__restrict__ diag, const double *__restrict__ params,
                                 const double *__restrict__ per_step, double *__restrict__ mu, double *__restrict__ var)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double jitter = 0.0;
    for (int k = 0; k < pl.n_jit; ++k) jitter += plan_jitter(pl, params, k);
    const MeanEval mean = plan_mean(pl, params);
    const double sig = diag[n] + jitter, alpha = per_step[n], kinv = per_step[N + n];
    const double m = mean(t[n]);
    mu[n] = m + (y[n] - m) - sig * alpha;
    var[n] = sig - sig * sig * kinv;
}

extern "C" int mtg_loglike_grad_dev(const mtg_model_t *, int, int, int, int, const double *, const double *,
                                    const double *, const double *, const double *, double *, double *,
                                    int32_t *, uint32_t, void *);

extern "C" int mtg_predict_host(const mtg_model_t *model, int N, const double *t, const double *y, const double *diag,
                                const double *params, double *out_mu, double *out_var)
{
    if (!model || N < 1 || !t || !y || !diag || !out_mu || !out_var) { mtg_set_error("mtg_predict_host: bad argument"); return MTG_EINVAL; }
    Plan pl;
    int rc = mtg_build_plan(model, &pl);
    if (rc) return rc;
    const int nd = pl.n_free > 0 ? pl.n_free : 1;
    double *buf = nullptr;
    const size_t n_dbl = 3 * (size_t)N + nd + 1 + nd + 2 * (size_t)N + 2 * (size_t)N;
    MTG_CUDA(cudaMalloc((void **)&buf, n_dbl * sizeof(double) + 64));
    double *d_t = buf, *d_y = d_t + N, *d_d = d_y + N, *d_p = d_d + N, *d_ll = d_p + nd, *d_g = d_ll + 1,
           *d_ps = d_g + nd, *d_mu = d_ps + 2 * N, *d_var = d_mu + N;
    int32_t *d_st = (int32_t *)(d_var + N);
    cudaStream_t st = 0;
    cudaMemcpyAsync(d_t, t, sizeof(double) * N, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_y, y, sizeof(double) * N, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_d, diag, sizeof(double) * N, cudaMemcpyHostToDevice, st);
    if (pl.n_free > 0) cudaMemcpyAsync(d_p, params, sizeof(double) * pl.n_free, cudaMemcpyHostToDevice, st);
    double *tape = nullptr;
    rc = mtg_cuda_check(cudaMalloc((void **)&tape, mtg_k2_tape_bytes_per_series(pl, N) * 128), "tape");
    if (!rc) {
        BatchArgs a{N, 1, 1, 1, d_t, d_y, d_d, nullptr, d_p, d_ll, d_g, d_st, 0u, nullptr, nullptr, d_ps};
        rc = mtg_k2_run(pl, a, tape, 128, st);
    }
    int32_t hst = 0;
    if (!rc) rc = mtg_cuda_check(cudaMemcpy(&hst, d_st, sizeof(int32_t), cudaMemcpyDeviceToHost), "d2h");
    if (!rc && hst != 0) { mtg_set_error("mtg_predict_host: covariance matrix not positive definite / non-finite (status %d)", hst); rc = MTG_EINVAL; }
    if (!rc) {
        k_predict_finish<<<(N + 127) / 128, 128, 0, st>>>(pl, N, d_t, d_y, d_d, d_p, d_ps, d_mu, d_var);
        mtg_count_launch();
        rc = mtg_cuda_check(cudaMemcpy(out_mu, d_mu, sizeof(double) * N, cudaMemcpyDeviceToHost), "d2h");
        if (!rc) rc = mtg_cuda_check(cudaMemcpy(out_var, d_var, sizeof(double) * N, cudaMemcpyDeviceToHost), "d2h");
    }
    cudaFree(tape); cudaFree(buf);
    return rc;
}
