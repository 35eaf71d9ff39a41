// mtg_scan.cu -- K9: the celerite log-likelihood of ONE series spread over a whole CTA by a parallel-in-time scan.
//
// K1 gives every series one thread and walks its N observations in order: perfect when there are >= 10^5 series, but
// the calls the reference makes most often are tiny -- the 6..32 walkers of a half ensemble of emcee
// (mind_the_gaps/gpmodelling.py:247-264 -> _log_probability :127-152), or the single evaluation of
// GPModelling._neg_log_like (:155-169) -- and then K1 is a 1000-step dependent chain on one warp: 0.21 ms whatever
// the batch.  Here the chain itself is parallelised.
//
// Formulation.  A celerite kernel k(tau) = sum_j a_j e^{-c_j tau} + sum_j e^{-c_j tau}(a_j cos d_j tau + b_j sin d_j tau)
// (Foreman-Mackey et al. 2017, eq. 8/9) is the covariance of a linear-Gaussian state-space model with a block-diagonal
// transition per term (in the frame rotating with each complex term):
//     real term     x' = e^{-c dt} x,            P_inf = a,                    h = 1
//     complex term  x' = e^{-c dt} R(d dt) x,    P_inf = [[a, b], [b, a + 2cb/d]],  h = (1, 0)
// (P_inf - F P_inf F^T is the process noise; the (2,2) entry is the one that makes its generator diagonal, and is the
// stationary covariance of the SHO in these coordinates), so the log-likelihood is the prediction-error decomposition
// of a Kalman filter -- the same number as celerite's Cholesky recursion (and as the dense oracle: tests/test_scan_gpu.py,
// 1e-10 relative).  Kalman filtering is a prefix sum over the associative operator of Sarkka & Garcia-Fernandez (IEEE TAC
// 2021, "Temporal parallelization of Bayesian smoothers", sec. IV) on elements (A, b, C, eta, J):
//     (i then j):  M = (I + C_i J_j)^-1,  A = A_j M A_i,  b = A_j M (b_i + C_i eta_j) + b_j,  C = A_j M C_i A_j^T + C_j,
//                  eta = A_i^T M^T (eta_j - J_j b_i) + eta_i,  J = A_i^T M^T J_j A_i + J_i.
// One CTA = one series, 256 threads, thread k owns CH = ceil(N / 256) consecutive observations:
//   phase 1  composite element of the thread's observations, appended one by one with the rank-one form of the operator
//            (a Kalman step on the affine map + information accumulation: O(J^2) per observation, no inverse);
//   phase 2  inclusive scan of the 256 composites: 5 shuffle levels inside each warp, the 8 warp totals through shared
//            memory, one more combine for the exclusive prefix (only its mean b and covariance C are needed);
//   phase 3  every thread filters its own observations from that prefix and accumulates -1/2 (v^2/s + log s);
//            block reduction.
// Work is ~12x K1's per series, depth ~ (2 CH + 9) combines instead of N steps: for the C2 half ensemble (32 series,
// N = 1000) the call drops from 0.28 ms to 19 us.  Series of up to 4096 observations are staged in shared memory first.
// mtg_loglike_dev picks K9 for batches of up to max(148, 2 N) series of kernels with J <= 4 state dimensions (closed-form
// 1x1 .. 4x4 inverses in the operator) and K1 otherwise; MTG_FLAG_SCAN / MTG_FLAG_NO_SCAN force either.  The NumPy statement
// of the formulation is oracle/kalman.py (tests/test_oracle.py).
#include "mtg_philox.cuh"
#include "mtg_ensemble.cuh"

#include "mtg_scan.cuh"

namespace k9 {

template <int JR, int JC, bool DUAL>
__global__ void __launch_bounds__(K9_THREADS, 1)
k9_loglike_scan(const __grid_constant__ Plan pl, const BatchArgs a, int staged)
{
    __shared__ Scratch<JR + 2 * JC> sc;
    extern __shared__ double k9_series_sm[];          // [3][N] when staged
    const int b = blockIdx.x;
    const double *x = a.params + (size_t)b * pl.n_free;
    const bool inside = !(a.flags & MTG_FLAG_APPLY_PRIOR) || plan_in_prior(pl, x);
    const int ds = b / a.group;
    const bool dsm = (a.flags & MTG_FLAG_DS_MEAN) != 0;
    int st;
    const double *st_t = a.t, *st_y = a.y + ds, *st_d = a.diag + ds;
    size_t stride = (size_t)a.S;
    if (staged) {
        stage_series(k9_series_sm, a.N, a.t, a.y + ds, a.diag + ds, (size_t)a.S);
        st_t = k9_series_sm; st_y = k9_series_sm + a.N; st_d = k9_series_sm + 2 * a.N; stride = 1;
    }
    double ll = series_loglike<JR, JC, DUAL, false>(pl, x, a.N, st_t, st_y, st_d, stride, dsm, dsm ? a.ds_mean[ds] : 0.0, sc, &st);
    if (threadIdx.x == 0) {
        if (!inside) { ll = -INFINITY; st = 3; }
        a.out_ll[b] = ll;
        if (a.status) a.status[b] = st;
    }
}

}  // namespace k9

template <int JR, int JC, bool DUAL>
static int launch_k9_one(const Plan &pl, const BatchArgs &a, cudaStream_t stream)
{
    const int staged = a.N <= K9_STAGE_MAX_N;
    const size_t dyn = staged ? sizeof(double) * 3 * (size_t)a.N : 0;
    static bool opted[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !opted[dev]) {
        MTG_CUDA(cudaFuncSetAttribute(k9::k9_loglike_scan<JR, JC, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(double) * 3 * K9_STAGE_MAX_N)));
        opted[dev] = true;
    }
    k9::k9_loglike_scan<JR, JC, DUAL><<<a.B, K9_THREADS, dyn, stream>>>(pl, a, staged);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k9_loglike_scan launch");
}
template <int JR, int JC>
static int launch_k9(const Plan &pl, const BatchArgs &a, cudaStream_t stream)
{
    if (pl.has_sho) return launch_k9_one<JR, JC, true>(pl, a, stream);
    return launch_k9_one<JR, JC, false>(pl, a, stream);
}

// true when the scan kernel covers this model: <= 4 state dimensions.  (celerite's Matern-3/2 term is a complex term with
// d = eps and b / a = w0 / eps: its stationary covariance a [[1, x], [x, 1 + 2 x^2]], x = w0 / eps, has condition number
// ~4 x^2, harmless for any correlation length longer than a few hundred time units -- tests/test_scan_gpu.py.)
bool mtg_k9_supports(const Plan &pl)
{
    return pl.jr + 2 * pl.jc <= 4 && pl.jr + 2 * pl.jc >= 1 && pl.jr <= 3;       // same structures as K1's dispatch
}

#define K9_CASE(R, C) if (pl.jr == R && pl.jc == C) return launch_k9<R, C>(pl, a, stream);

int mtg_k9_dispatch(const Plan &pl, const BatchArgs &a, cudaStream_t stream)
{
    K9_CASE(1, 0) K9_CASE(2, 0) K9_CASE(3, 0) K9_CASE(0, 1) K9_CASE(1, 1)
    K9_CASE(2, 1) K9_CASE(0, 2)
    mtg_set_error("scan kernel: unsupported kernel structure: %d real + %d complex slots", pl.jr, pl.jc);
    return MTG_EUNSUP;
}

// =====================================================================================================
// K10: the stretch-move ensemble sampler for FEW ensembles -- one CTA per proposal, K9's scan as the log-probability.
//
// Same algorithm, same Philox counters and the same proposal arithmetic as K8 (mtg_ensemble.cu; emcee 3.1.4
// RedBlueMove / StretchMove as driven by mind_the_gaps/gpmodelling.py:247-264), so a run reproduces K8's chain except
// where a log-probability differing in the last bits flips an accept decision.  What changes is the mapping: K8 gives a
// proposal ONE thread (a 1000-step dependent chain: 0.28 ms per half step whatever the ensemble size); here ensemble e
// owns H = ceil(W/2) CTAs, CTA (e, slot) evaluates proposal `slot` of the half being updated with the parallel-in-time
// scan, its thread 0 accepts / rejects and writes the walker, and the H CTAs meet at a per-ensemble barrier in global
// memory before the other half moves.  Inside a half step nobody reads what anybody else writes (proposals read the
// complementary half and their own walker, and write their own walker), so ONE barrier per half step is enough.
// All E * H CTAs must be resident at once (one per SM at 255 registers): the launch is cooperative and
// mtg_ensemble_dev only routes here when they fit.
namespace k9 {

__device__ __forceinline__ void ens_barrier(unsigned *ctr, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
    }
    __syncthreads();
}

#define K10_MAX_W 512

template <int JR, int JC, bool DUAL>
__global__ void __launch_bounds__(K9_THREADS, 1)
k10_ensemble_scan(const __grid_constant__ Plan pl, const EnsembleArgs a, unsigned *barrier, double *cta_best, int staged)
{
    __shared__ Scratch<JR + 2 * JC> sc;
    extern __shared__ double k10_series_sm[];         // [3][N] when staged: the lightcurve is read from HBM / L2 once per CTA
    __shared__ uint32_t key[K10_MAX_W];
    __shared__ int32_t perm[K10_MAX_W];
    __shared__ double q[MTG_MAX_PARAMS], ps[MTG_MAX_PARAMS], bq[MTG_MAX_PARAMS];
    const int nd = pl.n_free, W = a.W, T = blockDim.x, tid = threadIdx.x;
    const int H = (W + 1) / 2;
    const int e = blockIdx.x / H, slot = blockIdx.x % H;
    const int ds = a.S == 1 ? 0 : e;
    const bool dsm = (a.flags & MTG_FLAG_DS_MEAN) != 0;
    const double dsv = dsm ? a.ds_mean[ds] : 0.0;
    double *pos_e = a.pos + (size_t)e * W * nd, *lp_e = a.lp + (size_t)e * W;
    unsigned *ctr = barrier + e;
    unsigned gen = 0;
    const uint2 pkey = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    const uint32_t ekey = (uint32_t)e * 0x9E3779B9u;
    const bool track_best = a.best_lp != nullptr;
    double my_best = -INFINITY;                        // thread 0: best log-probability this CTA has seen

    const double *st_t = a.t, *st_y = a.y + ds, *st_d = a.diag + ds;
    size_t stride = (size_t)a.S;
    if (staged) {
        stage_series(k10_series_sm, a.N, a.t, a.y + ds, a.diag + ds, (size_t)a.S);
        st_t = k10_series_sm; st_y = k10_series_sm + a.N; st_d = k10_series_sm + 2 * a.N; stride = 1;
    }
    auto log_prob = [&](const double *x) -> double {   // CTA-collective; x in shared memory
        if (!plan_in_prior(pl, x)) return -INFINITY;   // uniform over the CTA
        int st;
        return series_loglike_call<JR, JC, DUAL>(pl, x, a.N, st_t, st_y, st_d, stride, dsm, dsv, sc, &st);
    };

    // ---- initial log-probabilities: this CTA owns walkers slot and slot + H at the start ----------------
    for (int w = slot; w < W; w += H) {
        __syncthreads();
        if (tid < nd) q[tid] = __ldcg(pos_e + (size_t)w * nd + tid);
        __syncthreads();
        double l0;
        if (a.init_lp) { l0 = log_prob(q); if (tid == 0) __stcg(lp_e + w, l0); }
        else l0 = __ldcg(lp_e + w);
        if (tid == 0 && track_best && l0 > my_best) { my_best = l0; for (int k = 0; k < nd; ++k) bq[k] = q[k]; }
    }
    ++gen; ens_barrier(ctr, gen * (unsigned)H);

    const int n0 = H;                                    // emcee: ceil(W/2) walkers in split 0
    for (int it = 0; it < a.nsteps; ++it) {
        const uint64_t step = a.step0 + (uint64_t)it;
        // random split of the ensemble: rank of a random key = a uniformly random permutation (every CTA redoes it)
        if (!a.inj.split)
            for (int idx = tid; idx < W; idx += T)
                key[idx] = philox4x32_10(make_uint4((uint32_t)idx, 2u, (uint32_t)step, (uint32_t)(step >> 32) ^ ekey), pkey).x;
        __syncthreads();
        for (int w = tid; w < W; w += T) {
            int r = 0;
            if (a.inj.split) r = injected_rank(a.inj.split + ((size_t)e * a.nsteps + it) * W, w, n0);
            else {
                const uint32_t kw = key[w];
                for (int k = 0; k < W; ++k) { const uint32_t v = key[k]; r += (v < kw) || (v == kw && k < w); }
            }
            perm[r] = w;
        }
        __syncthreads();
        for (int h = 0; h < 2; ++h) {
            const int s_off = h == 0 ? 0 : n0, s_cnt = h == 0 ? n0 : W - n0;
            const int c_off = h == 0 ? n0 : 0, c_cnt = W - s_cnt;
            if (slot < s_cnt) {
                const int w = perm[s_off + slot];
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
                if (tid < nd) {
                    const double pc = __ldcg(pos_e + (size_t)perm[j] * nd + tid), p0 = __ldcg(pos_e + (size_t)w * nd + tid);
                    ps[tid] = p0;
                    q[tid] = stretch_q(pc, p0, zz);
                }
                __syncthreads();
                const double lq = log_prob(q);
                if (tid == 0) {
                    const double lpw = __ldcg(lp_e + w);
                    const bool accept = log(uacc) < stretch_lnpdiff(nd, zz, lq, lpw);       // NaN compares false
                    if (accept) {
                        for (int k = 0; k < nd; ++k) __stcg(pos_e + (size_t)w * nd + k, q[k]);
                        __stcg(lp_e + w, lq);
                        if (a.n_accept) a.n_accept[(size_t)e * W + w] += 1;
                        if (track_best && lq > my_best) { my_best = lq; for (int k = 0; k < nd; ++k) bq[k] = q[k]; }
                    }
                    if (a.chain)
                        for (int k = 0; k < nd; ++k)
                            __stcs(a.chain + ((size_t)e * a.nsteps + it) * (size_t)W * nd + (size_t)w * nd + k, accept ? q[k] : ps[k]);
                    if (a.chain_lp) __stcs(a.chain_lp + ((size_t)e * a.nsteps + it) * (size_t)W + w, accept ? lq : lpw);
                }
            }
            ++gen; ens_barrier(ctr, gen * (unsigned)H);
        }
    }
    // ---- running maximum over the ensemble ----------------------------------------------------------------
    if (track_best) {
        double *mine = cta_best + ((size_t)e * H + slot) * (size_t)(nd + 1);
        if (tid == 0) { __stcg(mine, my_best); for (int k = 0; k < nd; ++k) __stcg(mine + 1 + k, bq[k]); }
        ++gen; ens_barrier(ctr, gen * (unsigned)H);
        if (slot == 0 && tid == 0) {
            double best = a.init_lp ? -INFINITY : a.best_lp[e]; int arg = -1;
            for (int s2 = 0; s2 < H; ++s2) {
                const double v = __ldcg(cta_best + ((size_t)e * H + s2) * (size_t)(nd + 1));
                if (v > best) { best = v; arg = s2; }
            }
            if (arg >= 0) {
                a.best_lp[e] = best;
                if (a.best_pos) for (int k = 0; k < nd; ++k) a.best_pos[(size_t)e * nd + k] = __ldcg(cta_best + ((size_t)e * H + arg) * (size_t)(nd + 1) + 1 + k);
            } else if (a.init_lp) a.best_lp[e] = best;
        }
    }
}

}  // namespace k9

static int k10_device_sms()
{
    static int sms[64]; static bool have[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 0;
    if (!have[dev]) { cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev); have[dev] = true; }
    return sms[dev];
}

bool mtg_k10_fits(const Plan &pl, int N, int E, int W)
{
    if (!mtg_k9_supports(pl) || W > K10_MAX_W || N < 1) return false;
    const long long ctas = (long long)E * ((W + 1) / 2);
    return ctas >= 1 && ctas <= k10_device_sms();          // one CTA per SM (255 registers x 256 threads)
}

template <int JR, int JC, bool DUAL>
static int launch_k10_one(const Plan &pl, const EnsembleArgs &a, cudaStream_t stream)
{
    const int H = (a.W + 1) / 2, nd = pl.n_free;
    if ((2.0 * a.nsteps + 4.0) * H >= 4.0e9) {       // the per-ensemble barrier counts arrivals in 32 bits
        mtg_set_error("mtg_ensemble_dev: %d steps per launch overflow the barrier counter of the scan sampler; use blocks of <= %d steps",
                      a.nsteps, (int)(1.9e9 / H));
        return MTG_EINVAL;
    }
    const size_t nb = sizeof(unsigned) * (size_t)a.E, nbest = sizeof(double) * (size_t)a.E * H * (nd + 1);
    int rc = mtg_pool_setup();
    if (rc) return rc;
    unsigned char *scratch = nullptr;
    const size_t off_best = (nb + 255) / 256 * 256;
    MTG_CUDA(cudaMallocAsync((void **)&scratch, off_best + nbest, stream));
    MTG_CUDA(cudaMemsetAsync(scratch, 0, off_best, stream));
    unsigned *barrier = reinterpret_cast<unsigned *>(scratch);
    double *cta_best = reinterpret_cast<double *>(scratch + off_best);
    Plan plc = pl; EnsembleArgs ac = a;
    int staged = a.N <= K9_STAGE_MAX_N;
    const size_t dyn = staged ? sizeof(double) * 3 * (size_t)a.N : 0;
    MTG_CUDA(cudaFuncSetAttribute(k9::k10_ensemble_scan<JR, JC, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(double) * 3 * K9_STAGE_MAX_N)));
    void *params[] = {&plc, &ac, &barrier, &cta_best, &staged};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k9::k10_ensemble_scan<JR, JC, DUAL>, dim3(a.E * H), dim3(K9_THREADS), params, dyn, stream);
    mtg_count_launch();
    cudaFreeAsync(scratch, stream);
    return mtg_cuda_check(e, "k10_ensemble_scan launch");
}
template <int JR, int JC>
static int launch_k10(const Plan &pl, const EnsembleArgs &a, cudaStream_t stream)
{
    if (pl.has_sho) return launch_k10_one<JR, JC, true>(pl, a, stream);
    return launch_k10_one<JR, JC, false>(pl, a, stream);
}
#define K10_CASE(R, C) if (pl.jr == R && pl.jc == C) return launch_k10<R, C>(pl, a, stream);

int mtg_k10_dispatch(const Plan &pl, const EnsembleArgs &a, cudaStream_t stream)
{
    K10_CASE(1, 0) K10_CASE(2, 0) K10_CASE(3, 0) K10_CASE(0, 1) K10_CASE(1, 1)
    K10_CASE(2, 1) K10_CASE(0, 2)
    mtg_set_error("scan sampler: unsupported kernel structure: %d real + %d complex slots", pl.jr, pl.jc);
    return MTG_EUNSUP;
}
