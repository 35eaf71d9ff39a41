// mtg_ensemble.cu -- K8: device-resident affine-invariant ensemble sampler (stretch move).
//
// Replaces the inner loop of emcee 3.1.4 as driven by GPModelling.derive_posteriors
// (mind_the_gaps/gpmodelling.py:247-264): EnsembleSampler.sample -> RedBlueMove (nsplits = 2, random
// split) -> StretchMove.get_proposal (a = 2) -> log_prob of the half ensemble -> accept / reject, with
// GPModelling._log_probability (:127-152) as the log-probability.  emcee pays one Python round trip and one
// pool.map per HALF step; here a whole run of steps is ONE kernel launch:
//
//   one CTA = one ensemble (one lightcurve + W walkers); thread i of the CTA owns walker S[i] of the half being
//   updated, proposes, evaluates the celerite recursion over the N observations (state in registers, K1's step),
//   accepts / rejects, and the two halves alternate separated by __syncthreads.  Positions and log-probabilities
//   live in shared memory; the chain (steps, walkers, ndim) and its log-probabilities stream to HBM.
//   E ensembles (the simulated lightcurves of the LRT, or independent chains on one lightcurve) run in lock step
//   as E CTAs: the batch dimension the reference spreads over a multiprocessing.Pool.
//
// Random numbers: Philox4x32-10 keyed by the call's seed, counter = (walker slot, half, absolute step, ensemble),
// so a run is reproducible and independent of how it is cut into launches.  The stream differs from emcee's
// MT19937 by construction (parity of the sampler is statistical: tests/test_ensemble_gpu.py).
#include "mtg_recursion.cuh"
#include "mtg_philox.cuh"
#include "mtg_ensemble.cuh"
#include <vector>

#define K8_MAX_THREADS 128          // >= ceil(W/2): at most 256 walkers per ensemble
#ifndef K8_MIN_BLOCKS
#define K8_MIN_BLOCKS 4          // 128 registers, no spills: 16 warps per SM (measured: 1.21e8 vs 1.12e8 evals/s at 2368 ensembles)
#endif
#ifndef K8_INLINE
#define K8_INLINE __forceinline__
#endif


// log posterior of one parameter vector on data set `ds` (box prior + celerite log-likelihood).  Deliberately NOT
// inlined: the recursion loop then gets K1's register allocation (128 registers, 16 warps per SM) instead of sharing
// one with the sampler bookkeeping (194 registers when inlined, 10 warps per SM: 25 % slower at thousands of ensembles).
template <int JR, int JC, bool DUAL>
__device__ K8_INLINE double k8_log_prob(const Plan &pl, const EnsembleArgs &a, const mtgm::Tabs &tb,
                                              const double *x, int ds)
{
    if (!plan_in_prior(pl, x)) return -INFINITY;
    Coef<JR, JC, DUAL> c;
    plan_coefficients<JR, JC, DUAL>(pl, x, c);
    MeanEval mean = plan_mean(pl, x);
    if (a.flags & MTG_FLAG_DS_MEAN) { mean.kind = MTG_MEAN_CONSTANT; mean.p0 = a.ds_mean[ds]; }
    const double *yp = a.y + ds, *dp = a.diag + ds;
    const size_t S = (size_t)a.S;
    const int N = a.N;
    K1State<JR, JC, DUAL> s;
    s.init(__ldg(a.t));
    double t_nx = __ldg(a.t), y_nx = __ldg(yp), d_nx = __ldg(dp);
#pragma unroll 2
    for (int n = 0; n < N; ++n) {
        const double tn = t_nx, yn = y_nx, dn = d_nx;
        if (n + 1 < N) { t_nx = __ldg(a.t + n + 1); y_nx = __ldg(yp + (size_t)(n + 1) * S); d_nx = __ldg(dp + (size_t)(n + 1) * S); }
        s.step(c, tb, tn, yn - mean(tn), dn, n);
    }
    int st;
    return s.loglike(N, &st);          // -inf when not positive definite / non-finite (celerite would raise)
}

template <int JR, int JC, bool DUAL>
__global__ void __launch_bounds__(K8_MAX_THREADS, K8_MIN_BLOCKS)
k8_ensemble(const __grid_constant__ Plan pl, const EnsembleArgs a)
{
    extern __shared__ double2 k8_smem2[];
    double *smem = reinterpret_cast<double *>(k8_smem2);
    const mtgm::Tabs tb = mtgm::tabs_fill(smem);
    const int nd = pl.n_free, W = a.W, T = blockDim.x, tid = threadIdx.x, EPC = a.epc;
    const int e0 = blockIdx.x * EPC;                     // first ensemble of this CTA
    const int n_ens = min(EPC, a.E - e0);                // ensembles this CTA really has
    const int WT = EPC * W;                              // walker slots of the CTA
    double *pos = smem + MTG_TABS_DOUBLES;               // [EPC][W][nd]
    double *lp = pos + (size_t)WT * nd;                  // [EPC][W]
    double *blp = lp + WT;                               // [EPC][W]   best log-probability seen by each walker
    double *bpos = blp + WT;                             // [EPC][W][nd]
    uint32_t *key = reinterpret_cast<uint32_t *>(bpos + (size_t)WT * nd);    // [EPC][W] random keys of the split
    int32_t *perm = reinterpret_cast<int32_t *>(key + WT);                   // [EPC][W]
    int32_t *nacc = perm + WT;                           // [EPC][W]
    const uint2 pkey = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const bool track_best = a.best_lp != nullptr;
    auto ds_of = [&](int e) { return a.S == 1 ? 0 : e; };

    for (int idx = tid; idx < n_ens * W * nd; idx += T) pos[idx] = a.pos[(size_t)e0 * W * nd + idx];
    for (int idx = tid; idx < n_ens * W; idx += T) { lp[idx] = a.init_lp ? 0.0 : a.lp[(size_t)e0 * W + idx]; nacc[idx] = 0; }
    __syncthreads();
    if (a.init_lp)
        for (int idx = tid; idx < n_ens * W; idx += T)
            lp[idx] = k8_log_prob<JR, JC, DUAL>(pl, a, tb, pos + (size_t)idx * nd, ds_of(e0 + idx / W));
    __syncthreads();
    if (track_best)
        for (int idx = tid; idx < n_ens * W; idx += T) {
            blp[idx] = lp[idx];
            for (int k = 0; k < nd; ++k) bpos[(size_t)idx * nd + k] = pos[(size_t)idx * nd + k];
        }

    const int n0 = (W + 1) / 2;                          // emcee: inds = arange(W) % 2, shuffled -> ceil(W/2) walkers in split 0
    const int sub = tid / n0, slot = tid % n0;           // this thread's ensemble within the CTA and walker slot within a half
    const bool lane_on = sub < n_ens;
    const int e = e0 + sub;
    const uint32_t ekey = (uint32_t)e * 0x9E3779B9u;
    double *mpos = pos + (size_t)sub * W * nd, *mlp = lp + (size_t)sub * W;
    const int32_t *mperm = perm + (size_t)sub * W;
    for (int it = 0; it < a.nsteps; ++it) {
        const uint64_t step = a.step0 + (uint64_t)it;
        // ---- random split: rank of a random key (ties broken by index) = a uniformly random permutation ----
        if (!a.inj.split)
            for (int idx = tid; idx < n_ens * W; idx += T) {
                const uint32_t ee = (uint32_t)(e0 + idx / W) * 0x9E3779B9u;
                key[idx] = philox4x32_10(make_uint4((uint32_t)(idx % W), 2u, (uint32_t)step, (uint32_t)(step >> 32) ^ ee), pkey).x;
            }
        __syncthreads();
        for (int idx = tid; idx < n_ens * W; idx += T) {
            const int sb = idx / W, w = idx % W;
            int r = 0;
            if (a.inj.split) r = injected_rank(a.inj.split + ((size_t)(e0 + sb) * a.nsteps + it) * W, w, n0);
            else {
                const uint32_t *kk = key + (size_t)sb * W;
                const uint32_t kw = kk[w];
                for (int k = 0; k < W; ++k) { const uint32_t v = kk[k]; r += (v < kw) || (v == kw && k < w); }
            }
            perm[(size_t)sb * W + r] = w;
        }
        __syncthreads();
        // ---- the two half steps ----
        for (int h = 0; h < 2; ++h) {
            const int s_off = h == 0 ? 0 : n0, s_cnt = h == 0 ? n0 : W - n0;
            const int c_off = h == 0 ? n0 : 0, c_cnt = W - s_cnt;
            bool accept = false; int w = 0; double lq = -INFINITY;
            double q[MTG_MAX_PARAMS];
            if (lane_on && slot < s_cnt) {
                w = mperm[s_off + slot];
                double uz, uacc; int rj;
                if (a.inj.split) {
                    const size_t at = (((size_t)e * a.nsteps + it) * 2 + h) * n0 + slot;
                    uz = a.inj.uz[at]; rj = a.inj.rint[at]; uacc = a.inj.uacc[at];
                } else {
                    const uint4 r0 = philox4x32_10(make_uint4((uint32_t)slot, (uint32_t)h, (uint32_t)step, (uint32_t)(step >> 32) ^ ekey), pkey);
                    const uint4 r1 = philox4x32_10(make_uint4((uint32_t)slot, (uint32_t)h + 4u, (uint32_t)step, (uint32_t)(step >> 32) ^ ekey), pkey);
                    uz = u01_from_bits(r0.x, r0.y);
                    rj = (int)(((uint64_t)r0.z * (uint64_t)c_cnt) >> 32);        // randint(Nc)
                    uacc = u01_from_bits(r1.x, r1.y);
                }
                const double zz = stretch_zz(a.a, uz);                           // StretchMove: ((a-1) U + 1)^2 / a
                const int j = c_off + rj;
                const double *pc = mpos + (size_t)mperm[j] * nd, *ps = mpos + (size_t)w * nd;
                for (int k = 0; k < nd; ++k) q[k] = stretch_q(pc[k], ps[k], zz);
                lq = k8_log_prob<JR, JC, DUAL>(pl, a, tb, q, ds_of(e));
                accept = log(uacc) < stretch_lnpdiff(nd, zz, lq, mlp[w]);        // NaN compares false
            }
            __syncthreads();          // every proposal of this half has read the positions it needs
            if (accept) {
                for (int k = 0; k < nd; ++k) mpos[(size_t)w * nd + k] = q[k];
                mlp[w] = lq; nacc[(size_t)sub * W + w] += 1;
                if (track_best && lq > blp[(size_t)sub * W + w]) {
                    blp[(size_t)sub * W + w] = lq;
                    for (int k = 0; k < nd; ++k) bpos[((size_t)sub * W + w) * nd + k] = q[k];
                }
            }
            __syncthreads();
        }
        if (a.chain)
            for (int idx = tid; idx < n_ens * W * nd; idx += T) {
                const int sb = idx / (W * nd), r = idx % (W * nd);
                __stcs(a.chain + ((size_t)(e0 + sb) * a.nsteps + it) * (size_t)W * nd + r, pos[idx]);
            }
        if (a.chain_lp)
            for (int idx = tid; idx < n_ens * W; idx += T) {
                const int sb = idx / W, w = idx % W;
                __stcs(a.chain_lp + ((size_t)(e0 + sb) * a.nsteps + it) * (size_t)W + w, lp[idx]);
            }
    }
    __syncthreads();
    for (int idx = tid; idx < n_ens * W * nd; idx += T) a.pos[(size_t)e0 * W * nd + idx] = pos[idx];
    for (int idx = tid; idx < n_ens * W; idx += T) {
        a.lp[(size_t)e0 * W + idx] = lp[idx];
        if (a.n_accept) a.n_accept[(size_t)e0 * W + idx] += nacc[idx];
    }
    if (track_best)
        for (int sb = tid; sb < n_ens; sb += T) {        // max over the walkers (and over the previous launches of the run)
            double best = a.init_lp ? -INFINITY : a.best_lp[e0 + sb]; int arg = -1;
            for (int w = 0; w < W; ++w) if (blp[(size_t)sb * W + w] > best) { best = blp[(size_t)sb * W + w]; arg = w; }
            if (arg >= 0) {
                a.best_lp[e0 + sb] = best;
                if (a.best_pos) for (int k = 0; k < nd; ++k) a.best_pos[(size_t)(e0 + sb) * nd + k] = bpos[((size_t)sb * W + arg) * nd + k];
            } else if (a.init_lp) a.best_lp[e0 + sb] = best;
        }
}

// Launch shape: a half ensemble (ceil(W/2) walkers) is updated by consecutive threads; ensembles smaller than a warp
// share one (EPC ensembles per CTA) so that the 32 lanes are used.  T >= EPC * ceil(W/2) always.
template <int JR, int JC, bool DUAL>
static int launch_k8_one(const Plan &pl, EnsembleArgs a, cudaStream_t stream)
{
    const int half = (a.W + 1) / 2;
    int epc = half >= 32 ? 1 : 32 / half;
    if (epc > a.E) epc = a.E;
    a.epc = epc;
    const int T = 32 * ((epc * half + 31) / 32);
    const size_t WT = (size_t)epc * a.W;
    const size_t smem = (MTG_TABS_DOUBLES + 2 * WT * pl.n_free + 2 * WT) * sizeof(double) + 3 * WT * sizeof(int32_t) + 16;
    auto kern = k8_ensemble<JR, JC, DUAL>;
    MTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(a.E + epc - 1) / epc, T, smem, stream>>>(pl, a);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k8_ensemble launch");
}
template <int JR, int JC>
static int launch_k8(const Plan &pl, const EnsembleArgs &a, cudaStream_t stream)
{
    if (pl.has_sho) return launch_k8_one<JR, JC, true>(pl, a, stream);
    return launch_k8_one<JR, JC, false>(pl, a, stream);
}
#define K8_CASE(R, C) if (pl.jr == R && pl.jc == C) return launch_k8<R, C>(pl, a, stream);

static int ensemble_run_dev(const mtg_model_t *model, int N, int S, int E, int W,
                            const double *t, const double *y, const double *diag, const double *ds_mean,
                            double *pos, double *lp, int init_lp, int nsteps, double a_scale, uint64_t seed, uint64_t step0,
                            double *chain, double *chain_lp, int32_t *n_accept, double *best_lp, double *best_pos,
                            uint32_t flags, const EnsembleInject *inj, void *stream_)
{
    if (!model || N < 1 || S < 1 || E < 0 || W < 2 || nsteps < 0 || !t || !y || !diag || !pos || !lp || !(a_scale > 1.0)) {
        mtg_set_error("mtg_ensemble_dev: bad argument"); return MTG_EINVAL;
    }
    if (S != 1 && S != E) { mtg_set_error("mtg_ensemble_dev: S must be 1 (shared lightcurve) or E (one per ensemble)"); return MTG_EINVAL; }
    if ((flags & MTG_FLAG_DS_MEAN) && !ds_mean) { mtg_set_error("MTG_FLAG_DS_MEAN without ds_mean"); return MTG_EINVAL; }
    if (W > 2 * K8_MAX_THREADS) { mtg_set_error("mtg_ensemble_dev: at most %d walkers per ensemble", 2 * K8_MAX_THREADS); return MTG_EINVAL; }
    if (E == 0) return 0;
    Plan pl;
    int rc = mtg_build_plan(model, &pl);
    if (rc) return rc;
    if (pl.n_free < 1) { mtg_set_error("mtg_ensemble_dev: the model has no free parameters"); return MTG_EINVAL; }
    if (W < 2 * pl.n_free) { mtg_set_error("the number of walkers must be at least twice the dimension (got %d for ndim=%d)", W, pl.n_free); return MTG_EINVAL; }
    EnsembleArgs a{N, S, E, W, nsteps, t, y, diag, ds_mean, pos, lp, chain, chain_lp, n_accept, best_lp, best_pos, a_scale, seed, step0, flags, init_lp, 1,
                   inj ? *inj : EnsembleInject{nullptr, nullptr, nullptr, nullptr}};
    cudaStream_t stream = (cudaStream_t)stream_;
    // Few ensembles (derive_posteriors on ONE lightcurve): K8 would run each half ensemble as a 1000-step chain on one
    // warp.  K10 gives every proposal a whole CTA (parallel-in-time scan, mtg_scan.cu) and separates the half steps by a
    // per-ensemble barrier across CTAs: same proposals, same random numbers, ~15x the steps per second.
    const bool can_scan = mtg_k10_fits(pl, N, E, W);
    if ((flags & MTG_FLAG_SCAN) && !can_scan) { mtg_set_error("MTG_FLAG_SCAN: model / ensemble count not covered by the scan sampler"); return MTG_EUNSUP; }
    if ((flags & MTG_FLAG_SCAN) || (can_scan && !(flags & MTG_FLAG_NO_SCAN) && N >= 64)) return mtg_k10_dispatch(pl, a, stream);
    K8_CASE(1, 0) K8_CASE(2, 0) K8_CASE(3, 0)
    K8_CASE(0, 1) K8_CASE(1, 1) K8_CASE(2, 1) K8_CASE(3, 1)
    K8_CASE(0, 2) K8_CASE(1, 2) K8_CASE(2, 2)
    K8_CASE(0, 3) K8_CASE(1, 3) K8_CASE(2, 3)
    K8_CASE(0, 4) K8_CASE(1, 4)
    mtg_set_error("unsupported kernel structure: %d real + %d complex recursion slots", pl.jr, pl.jc);
    return MTG_EUNSUP;
}

extern "C" int mtg_ensemble_dev(const mtg_model_t *model, int N, int S, int E, int W,
                                const double *t, const double *y, const double *diag, const double *ds_mean,
                                double *pos, double *lp, int init_lp, int nsteps, double a_scale, uint64_t seed, uint64_t step0,
                                double *chain, double *chain_lp, int32_t *n_accept, double *best_lp, double *best_pos,
                                uint32_t flags, void *stream_)
{
    return ensemble_run_dev(model, N, S, E, W, t, y, diag, ds_mean, pos, lp, init_lp, nsteps, a_scale, seed, step0, chain, chain_lp,
                            n_accept, best_lp, best_pos, flags, nullptr, stream_);
}

// Host-buffer entry point: y, diag are [S][N] ROW-major on the host (one lightcurve per row), transposed here.
// inj_*: host arrays of injected draws (all or none), see EnsembleInject.
static int ensemble_host_impl(const mtg_model_t *model, int N, int S, int E, int W,
                              const double *t, const double *y_rows, const double *diag_rows, const double *ds_mean,
                              double *pos, double *lp, int init_lp, int nsteps, double a_scale, uint64_t seed, uint64_t step0,
                              double *chain, double *chain_lp, int32_t *n_accept, uint32_t flags,
                              const int32_t *inj_split, const double *inj_uz, const int32_t *inj_rint, const double *inj_uacc)
{
    if (!model || N < 1 || S < 1 || E < 1 || !t || !y_rows || !diag_rows || !pos || !lp) { mtg_set_error("mtg_ensemble_host: bad argument"); return MTG_EINVAL; }
    const int nd = model->n_free;
    cudaStream_t st = nullptr;
    MTG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int rc = mtg_pool_setup();
    // injected draws (test hook): one device block [split | rint | uz | uacc]
    EnsembleInject inj{nullptr, nullptr, nullptr, nullptr};
    unsigned char *d_inj = nullptr;
    if (inj_split) {
        const size_t H = (size_t)(W + 1) / 2, n_sp = (size_t)E * nsteps * W, n_h = (size_t)E * nsteps * 2 * H;
        const size_t off_ri = n_sp * 4, off_uz = (off_ri + n_h * 4 + 7) / 8 * 8, off_ua = off_uz + n_h * 8;
        if (!rc) rc = mtg_cuda_check(cudaMallocAsync((void **)&d_inj, off_ua + n_h * 8 + 8, st), "cudaMallocAsync");
        if (rc) { cudaStreamDestroy(st); return rc; }
        cudaMemcpyAsync(d_inj, inj_split, n_sp * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_inj + off_ri, inj_rint, n_h * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_inj + off_uz, inj_uz, n_h * 8, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_inj + off_ua, inj_uacc, n_h * 8, cudaMemcpyHostToDevice, st);
        inj.split = (const int32_t *)d_inj; inj.rint = (const int32_t *)(d_inj + off_ri);
        inj.uz = (const double *)(d_inj + off_uz); inj.uacc = (const double *)(d_inj + off_ua);
    }
    double *d = nullptr; int32_t *d_acc = nullptr;
    const size_t n_t = N, n_y = (size_t)N * S, n_pos = (size_t)E * W * nd, n_lp = (size_t)E * W, n_mean = ds_mean ? S : 0;
    const size_t n_chain = chain ? (size_t)E * nsteps * W * nd : 0, n_clp = chain_lp ? (size_t)E * nsteps * W : 0;
    const size_t total = n_t + 2 * n_y + n_pos + n_lp + n_mean + n_chain + n_clp;
    if (!rc) rc = mtg_cuda_check(cudaMallocAsync((void **)&d, total * sizeof(double), st), "cudaMallocAsync");
    if (!rc) rc = mtg_cuda_check(cudaMallocAsync((void **)&d_acc, n_lp * sizeof(int32_t), st), "cudaMallocAsync");
    if (rc) { cudaStreamDestroy(st); return rc; }
    double *d_t = d, *d_y = d_t + n_t, *d_dg = d_y + n_y, *d_pos = d_dg + n_y, *d_lp = d_pos + n_pos, *d_mean = d_lp + n_lp,
           *d_chain = d_mean + n_mean, *d_clp = d_chain + n_chain;
    // time-major copies of the lightcurves: column s of [N][S] <- row s
    cudaMemcpyAsync(d_t, t, n_t * sizeof(double), cudaMemcpyHostToDevice, st);
    std::vector<double> ty, td;
    if (S > 1) {
        ty.resize(n_y); td.resize(n_y);
        for (int s = 0; s < S; ++s) for (int n = 0; n < N; ++n) { ty[(size_t)n * S + s] = y_rows[(size_t)s * N + n]; td[(size_t)n * S + s] = diag_rows[(size_t)s * N + n]; }
        cudaMemcpyAsync(d_y, ty.data(), n_y * sizeof(double), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_dg, td.data(), n_y * sizeof(double), cudaMemcpyHostToDevice, st);
    } else {
        cudaMemcpyAsync(d_y, y_rows, n_y * sizeof(double), cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(d_dg, diag_rows, n_y * sizeof(double), cudaMemcpyHostToDevice, st);
    }
    cudaMemcpyAsync(d_pos, pos, n_pos * sizeof(double), cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_lp, lp, n_lp * sizeof(double), cudaMemcpyHostToDevice, st);
    if (n_mean) cudaMemcpyAsync(d_mean, ds_mean, n_mean * sizeof(double), cudaMemcpyHostToDevice, st);
    if (n_accept) cudaMemcpyAsync(d_acc, n_accept, n_lp * sizeof(int32_t), cudaMemcpyHostToDevice, st);
    else cudaMemsetAsync(d_acc, 0, n_lp * sizeof(int32_t), st);
    rc = ensemble_run_dev(model, N, S, E, W, d_t, d_y, d_dg, n_mean ? d_mean : nullptr, d_pos, d_lp, init_lp, nsteps, a_scale, seed, step0,
                          chain ? d_chain : nullptr, chain_lp ? d_clp : nullptr, d_acc, nullptr, nullptr, flags, inj_split ? &inj : nullptr, st);
    if (!rc) {
        cudaMemcpyAsync(pos, d_pos, n_pos * sizeof(double), cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(lp, d_lp, n_lp * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (chain) cudaMemcpyAsync(chain, d_chain, n_chain * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (chain_lp) cudaMemcpyAsync(chain_lp, d_clp, n_clp * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (n_accept) cudaMemcpyAsync(n_accept, d_acc, n_lp * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        rc = mtg_cuda_check(cudaStreamSynchronize(st), "mtg_ensemble_host sync");
    }
    cudaFreeAsync(d, st); cudaFreeAsync(d_acc, st);
    if (d_inj) cudaFreeAsync(d_inj, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

extern "C" int mtg_ensemble_host(const mtg_model_t *model, int N, int S, int E, int W,
                                 const double *t, const double *y_rows, const double *diag_rows, const double *ds_mean,
                                 double *pos, double *lp, int init_lp, int nsteps, double a_scale, uint64_t seed, uint64_t step0,
                                 double *chain, double *chain_lp, int32_t *n_accept, uint32_t flags)
{
    return ensemble_host_impl(model, N, S, E, W, t, y_rows, diag_rows, ds_mean, pos, lp, init_lp, nsteps, a_scale, seed, step0,
                              chain, chain_lp, n_accept, flags, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int mtg_ensemble_injected_host(const mtg_model_t *model, int N, int S, int E, int W,
                                          const double *t, const double *y_rows, const double *diag_rows, const double *ds_mean,
                                          double *pos, double *lp, int nsteps, double a_scale,
                                          const int32_t *split, const double *uz, const int32_t *rint, const double *uacc,
                                          double *chain, double *chain_lp, int32_t *n_accept, uint32_t flags)
{
    if (!split || !uz || !rint || !uacc) { mtg_set_error("mtg_ensemble_injected_host: the four draw arrays are required"); return MTG_EINVAL; }
    const int H = (W + 1) / 2;
    for (size_t es = 0; es < (size_t)E * nsteps; ++es) {           // validate: ceil(W/2) zeros per split, partners inside the complement
        int zeros = 0;
        for (int w = 0; w < W; ++w) {
            const int v = split[es * W + w];
            if (v != 0 && v != 1) { mtg_set_error("mtg_ensemble_injected_host: split entries must be 0 or 1"); return MTG_EINVAL; }
            zeros += (v == 0);
        }
        if (zeros != H) { mtg_set_error("mtg_ensemble_injected_host: every split needs ceil(W/2) zeros"); return MTG_EINVAL; }
        for (int h = 0; h < 2; ++h) {
            const int ns = h == 0 ? H : W - H, nc = W - ns;
            for (int i = 0; i < ns; ++i) {
                const int r = rint[(es * 2 + h) * H + i];
                if (r < 0 || r >= nc) { mtg_set_error("mtg_ensemble_injected_host: rint out of range"); return MTG_EINVAL; }
            }
        }
    }
    return ensemble_host_impl(model, N, S, E, W, t, y_rows, diag_rows, ds_mean, pos, lp, 1, nsteps, a_scale, 0, 0,
                              chain, chain_lp, n_accept, flags, split, uz, rint, uacc);
}
