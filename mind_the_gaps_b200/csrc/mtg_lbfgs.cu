// mtg_lbfgs.cu -- K4: batched, box-constrained L-BFGS maximum-likelihood fits on the device.
//
// Replaces scipy.optimize.minimize(self._neg_log_like, x0, method="L-BFGS-B", bounds=...) of
// GPModelling.fit (mind_the_gaps/gpmodelling.py:172-194) for B independent problems (the null and
// alternative fits of every simulated lightcurve of the LRT, times their multi-starts).
//
// Structure: every problem is a small state machine kept in SoA device arrays.  One ROUND =
//   (1) K2 evaluates -loglike and its gradient at the trial point of every still-active problem
//       (one thread per problem, active list compacted in order with cub::DeviceSelect);
//   (2) k_lbfgs_update advances each problem: Armijo test -> accept (history update, convergence
//       tests, new two-loop direction with a projected-gradient active set) or shrink the step.
// The host only launches rounds and reads the active count every few rounds.
//
// TAIL: a K2 round costs ~1.1 ms at N = 1000 whatever the batch, and most rounds of a batched fit serve a few hundred slowly
// converging problems.  Once the active count drops below MTG_LBFGS_SWITCH (default 6144: about what makes one pass of the
// persistent kernel over the active problems as cheap as one K2 round) the remaining problems are handed to ONE launch of
// k_lbfgs_persist: every CTA takes a problem from a work queue, copies its optimiser state into shared memory and
// runs its whole remaining L-BFGS loop there -- gradient by the parallel-in-time kernel K11 (series_loglike_grad, ~30 us),
// update by thread 0 with the same state machine (lbfgs_advance) -- then writes the state back and takes the next one.
// No rounds, no compaction, no host synchronisation, and stragglers no longer hold the batch.  Small batches
// (GPModelling.fit on one lightcurve) run entirely in this kernel.
//
// Differences from scipy's L-BFGS-B (Byrd-Lu-Nocedal-Zhu): the generalised-Cauchy-point /
// subspace-minimisation step is replaced by a projected L-BFGS direction (variables at a bound with
// the gradient pointing outwards are frozen for the iteration) and the More-Thuente line search by
// Armijo backtracking with quadratic interpolation, an EXPANSION phase (the step is tripled while the
// Armijo point still violates the weak-Wolfe curvature condition, slope_t < 0.9 slope_0: without it the
// rare starts on a likelihood plateau crawl with steps of the size of their gradient for hundreds of
// iterations, and every straggler costs the whole batch one latency-bound round) and cautious (s.y > 0) updates.  Stopping tests
// are scipy's: projected-gradient inf-norm <= gtol, relative decrease <= ftol.  Gradients are
// analytic (K2) instead of 2-point finite differences.
#include "mtg_scan_grad.cuh"
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#define LB_THREADS 128
#define LB_MAXM 10
#define LB_EXPAND 3.0          // step-length growth factor of the expansion phase
#define LB_MAX_EXPAND 8       // at most 3^8 times the initial step

struct LbfgsState {
    int B, nd, m;
    double *X, *G, *Dir;          // [nd][B]
    double *SH, *YH;              // [m][nd][B]
    double *RHO;                  // [m][B]
    double *F, *alpha, *gd;       // [B]
    int32_t *hist_n, *hist_head, *ls_count, *iter, *phase, *status, *restarted;   // [B]
    double *XC, *GC, *FC;         // [nd][B], [nd][B], [B]  best Armijo point of the current line search (expansion phase)
    int32_t *has_cand, *n_exp;    // [B]
    double *XT, *LL, *GT;         // [B][nd], [B], [B][nd]   K2 input / outputs
    int32_t *k2_status;           // [B]
    int32_t *flags;               // [B] 1 = still active
};

struct LbfgsParams {
    int max_iter, max_ls; double ftol, gtol, max_step;
    int nd; double lo[MTG_MAX_PARAMS], hi[MTG_MAX_PARAMS];
};

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

__global__ void k_lbfgs_init(LbfgsState s, LbfgsParams p, const double *__restrict__ x0, int32_t *list)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    for (int k = 0; k < p.nd; ++k) s.XT[(size_t)b * p.nd + k] = clampd(x0[(size_t)b * p.nd + k], p.lo[k], p.hi[k]);
    s.phase[b] = 0; s.status[b] = 1; s.iter[b] = 0; s.hist_n[b] = 0; s.hist_head[b] = 0; s.ls_count[b] = 0;
    s.restarted[b] = 0; s.flags[b] = 1; s.F[b] = INFINITY; s.alpha[b] = 1.0; s.gd[b] = 0.0;
    s.has_cand[b] = 0; s.n_exp[b] = 0;
    list[b] = b;
}

// loop over the parameters: unrolled with a guard when NDMAX is small (register arrays), a plain loop otherwise
#define LB_STR2(x) #x
#define LB_STR(x) LB_STR2(x)
#define LB_K(k) _Pragma("unroll (NDMAX <= 8 ? NDMAX : 1)") for (int k = 0; k < (NDMAX <= 8 ? NDMAX : nd); ++k) if (NDMAX > 8 || k < nd)
#define LB_SMALL 8

// new search direction at the current point; returns false if the point is stationary
template <int NDMAX>
__device__ bool lbfgs_direction(const LbfgsState &s, const LbfgsParams &p, int b, double *alpha_out)
{
    const int nd = p.nd, B = s.B;
    double q[NDMAX], g[NDMAX];
    bool fixed[NDMAX];
    double qq = 0.0;
    LB_K(k) {
        const double xk = s.X[(size_t)k * B + b];
        g[k] = s.G[(size_t)k * B + b];
        fixed[k] = (xk <= p.lo[k] && g[k] > 0.0) || (xk >= p.hi[k] && g[k] < 0.0);
        q[k] = fixed[k] ? 0.0 : g[k];
        qq += q[k] * q[k];
    }
    if (!(qq > 0.0)) return false;
    int hn = s.hist_n[b];
    const int head = s.hist_head[b];       // slot of the NEXT write; newest = head-1
    double al[LB_MAXM];
    double d[NDMAX];
    LB_K(k) d[k] = q[k];
    for (int i = 0; i < hn; ++i) {
        const int slot = (head - 1 - i + 2 * s.m) % s.m;
        double sd = 0.0;
        LB_K(k) if (!fixed[k]) sd += s.SH[((size_t)slot * nd + k) * B + b] * d[k];
        al[i] = s.RHO[(size_t)slot * B + b] * sd;
        LB_K(k) if (!fixed[k]) d[k] -= al[i] * s.YH[((size_t)slot * nd + k) * B + b];
    }
    if (hn > 0) {
        const int slot = (head - 1 + s.m) % s.m;
        double sy = 0.0, yy = 0.0;
        LB_K(k) if (!fixed[k]) {
            const double yk = s.YH[((size_t)slot * nd + k) * B + b];
            sy += s.SH[((size_t)slot * nd + k) * B + b] * yk; yy += yk * yk;
        }
        const double gamma = (yy > 0.0 && sy > 0.0) ? sy / yy : 1.0;
        LB_K(k) d[k] *= gamma;
    }
    for (int i = hn - 1; i >= 0; --i) {
        const int slot = (head - 1 - i + 2 * s.m) % s.m;
        double yd = 0.0;
        LB_K(k) if (!fixed[k]) yd += s.YH[((size_t)slot * nd + k) * B + b] * d[k];
        const double be = s.RHO[(size_t)slot * B + b] * yd;
        LB_K(k) if (!fixed[k]) d[k] += (al[i] - be) * s.SH[((size_t)slot * nd + k) * B + b];
    }
    double gdir = 0.0;
    LB_K(k) { d[k] = fixed[k] ? 0.0 : -d[k]; gdir += g[k] * d[k]; }
    if (!(gdir < 0.0) || !isfinite(gdir)) {          // not a descent direction: steepest descent restart
        hn = 0; s.hist_n[b] = 0;
        gdir = 0.0;
        LB_K(k) { d[k] = -q[k]; gdir += g[k] * d[k]; }
    }
    double alpha = (hn == 0) ? fmin(1.0, 1.0 / sqrt(qq)) : 1.0;
    double dmax = 0.0;
    LB_K(k) dmax = fmax(dmax, fabs(d[k]));
    if (alpha * dmax > p.max_step) alpha = p.max_step / dmax;
    LB_K(k) s.Dir[(size_t)k * B + b] = d[k];
    s.gd[b] = gdir;
    *alpha_out = alpha;
    return true;
}

template <int NDMAX>
__device__ __forceinline__ void lbfgs_set_trial(const LbfgsState &s, const LbfgsParams &p, int b, double alpha)
{
    const int nd = p.nd, B = s.B;
    s.alpha[b] = alpha;
    LB_K(k)
        s.XT[(size_t)b * nd + k] = clampd(s.X[(size_t)k * B + b] + alpha * s.Dir[(size_t)k * B + b], p.lo[k], p.hi[k]);
}

// one transition of problem b's state machine after its trial point was evaluated (LL, GT, k2_status hold the result);
// returns true when the problem is finished (status / flags written).
// NDMAX: compile-time bound on the number of parameters.  NDMAX <= 8: every per-parameter array lives in registers and the
// loops are unrolled (guarded by k < nd) -- in the persistent kernel this update is a SERIAL section on one thread while
// 255 wait at a barrier (ncu r02c: 43 % of its stall samples), and the generic form spends it in local-memory traffic and
// loop overhead; NDMAX = MTG_MAX_PARAMS: the generic form.
template <int NDMAX>
__device__ bool lbfgs_advance(const LbfgsState &s, const LbfgsParams &p, int b)
{
    const int nd = p.nd, B = s.B;
    double ft = -s.LL[b];
    if (s.k2_status[b] != 0 || !isfinite(ft)) ft = INFINITY;
    double gt[NDMAX], xt[NDMAX];
    LB_K(k) { gt[k] = -s.GT[(size_t)b * nd + k]; xt[k] = s.XT[(size_t)b * nd + k]; }
    bool done = false; int status = 1;
    bool need_dir = false;

    if (s.phase[b] == 0) {
        if (!isfinite(ft)) { done = true; status = 2; LB_K(k) s.X[(size_t)k * B + b] = xt[k]; s.F[b] = ft; }
        else {
            LB_K(k) { s.X[(size_t)k * B + b] = xt[k]; s.G[(size_t)k * B + b] = gt[k]; }
            s.F[b] = ft; s.phase[b] = 1; need_dir = true;
        }
    } else {
        const double F = s.F[b];
        const bool armijo_at = [&]() {
            double gdx = 0.0;
            LB_K(k) gdx += s.G[(size_t)k * B + b] * (xt[k] - s.X[(size_t)k * B + b]);
            return isfinite(ft) && ft <= F + 1e-4 * gdx;
        }();
        // expansion test (weak Wolfe curvature condition): the slope along the direction at the trial point is still
        // nearly as steep as at the base point -> the step was too short (plateaus, negative curvature).  Only for
        // trial points that were not clipped by the box.
        auto wants_expansion = [&]() {
            if (s.n_exp[b] >= LB_MAX_EXPAND) return false;
            const double a0 = s.alpha[b];
            double slope = 0.0, dmax = 0.0;
            LB_K(k) {
                const double dk = s.Dir[(size_t)k * B + b];
                if (dk != 0.0 && (xt[k] <= p.lo[k] || xt[k] >= p.hi[k])) return false;
                slope += gt[k] * dk; dmax = fmax(dmax, fabs(dk));
            }
            if (!(slope < 0.9 * s.gd[b])) return false;
            return LB_EXPAND * a0 * dmax <= p.max_step;
        };
        // make (xa, fa, ga) the new iterate: history update, convergence tests
        auto accept = [&](const double *xa, double fa, const double *ga) {
            double ss = 0.0, yy = 0.0, sy = 0.0;
            double dx[NDMAX], dg[NDMAX];
            LB_K(k) {
                dx[k] = xa[k] - s.X[(size_t)k * B + b]; dg[k] = ga[k] - s.G[(size_t)k * B + b];
                ss += dx[k] * dx[k]; yy += dg[k] * dg[k]; sy += dx[k] * dg[k];
            }
            if (sy > 1e-10 * sqrt(ss) * sqrt(yy) && sy > 0.0) {
                const int slot = s.hist_head[b];
                LB_K(k) { s.SH[((size_t)slot * nd + k) * B + b] = dx[k]; s.YH[((size_t)slot * nd + k) * B + b] = dg[k]; }
                s.RHO[(size_t)slot * B + b] = 1.0 / sy;
                s.hist_head[b] = (slot + 1) % s.m;
                if (s.hist_n[b] < s.m) s.hist_n[b] += 1;
            }
            double pg = 0.0;
            LB_K(k) {
                s.X[(size_t)k * B + b] = xa[k]; s.G[(size_t)k * B + b] = ga[k];
                pg = fmax(pg, fabs(xa[k] - clampd(xa[k] - ga[k], p.lo[k], p.hi[k])));
            }
            s.F[b] = fa;
            const int it = s.iter[b] + 1; s.iter[b] = it;
            s.restarted[b] = 0; s.has_cand[b] = 0; s.n_exp[b] = 0;
            if (pg <= p.gtol) { done = true; status = 0; }
            else if ((F - fa) <= p.ftol * fmax(fmax(fabs(F), fabs(fa)), 1.0)) { done = true; status = 0; }
            else if (it >= p.max_iter) { done = true; status = 1; }
            else need_dir = true;
        };
        auto store_candidate_and_expand = [&]() {
            LB_K(k) { s.XC[(size_t)k * B + b] = xt[k]; s.GC[(size_t)k * B + b] = gt[k]; }
            s.FC[b] = ft; s.has_cand[b] = 1; s.n_exp[b] += 1;
            lbfgs_set_trial<NDMAX>(s, p, b, LB_EXPAND * s.alpha[b]);
        };
        if (s.has_cand[b]) {
            // this evaluation was an expansion trial; the stored candidate already satisfies Armijo
            if (armijo_at && ft < s.FC[b]) {
                if (wants_expansion()) store_candidate_and_expand();
                else accept(xt, ft, gt);
            } else {
                double xc[NDMAX], gc[NDMAX];
                LB_K(k) { xc[k] = s.XC[(size_t)k * B + b]; gc[k] = s.GC[(size_t)k * B + b]; }
                accept(xc, s.FC[b], gc);
            }
        } else if (armijo_at) {
            if (wants_expansion()) store_candidate_and_expand();
            else accept(xt, ft, gt);
        } else {
            const int lc = s.ls_count[b] + 1; s.ls_count[b] = lc;
            if (lc >= p.max_ls) {
                if (s.hist_n[b] > 0 && !s.restarted[b]) { s.hist_n[b] = 0; s.restarted[b] = 1; need_dir = true; }
                else { done = true; status = 2; }
            } else {
                const double a0 = s.alpha[b], gd = s.gd[b];
                double an = 0.25 * a0;
                if (isfinite(ft)) {
                    const double den = 2.0 * (ft - F - gd * a0);
                    if (den > 0.0) an = -gd * a0 * a0 / den;
                    an = fmin(fmax(an, 0.1 * a0), 0.5 * a0);
                }
                lbfgs_set_trial<NDMAX>(s, p, b, an);
            }
        }
    }
    if (need_dir) {
        double alpha;
        if (!lbfgs_direction<NDMAX>(s, p, b, &alpha)) { done = true; status = 0; }
        else { s.ls_count[b] = 0; s.n_exp[b] = 0; s.has_cand[b] = 0; lbfgs_set_trial<NDMAX>(s, p, b, alpha); }
    }
    if (done) { s.status[b] = status; s.flags[b] = 0; }
    return done;
}

__global__ void k_lbfgs_update(LbfgsState s, LbfgsParams p, const int32_t *__restrict__ list,
                               const int32_t *__restrict__ count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *count) return;
    if (p.nd <= LB_SMALL) lbfgs_advance<LB_SMALL>(s, p, list[i]);
    else lbfgs_advance<MTG_MAX_PARAMS>(s, p, list[i]);
}

// ---- the persistent tail kernel -------------------------------------------------------------------------------------------
// per-problem optimiser state in shared memory: the same SoA arrays with B = 1
struct LbfgsLocal {
    double X[MTG_MAX_PARAMS], G[MTG_MAX_PARAMS], Dir[MTG_MAX_PARAMS];
    double SH[LB_MAXM * MTG_MAX_PARAMS], YH[LB_MAXM * MTG_MAX_PARAMS], RHO[LB_MAXM];
    double F, alpha, gd;
    double XT[MTG_MAX_PARAMS], GT[MTG_MAX_PARAMS], LL;
    double XC[MTG_MAX_PARAMS], GC[MTG_MAX_PARAMS], FC;
    int32_t hist_n, hist_head, ls_count, iter, phase, status, restarted, k2_status, flags, has_cand, n_exp;
};

template <int JR, int JC, bool DUAL>
__global__ void __launch_bounds__(K9_THREADS, 1)
k_lbfgs_persist(const __grid_constant__ Plan pl, const BatchArgs a, const LbfgsState s, const __grid_constant__ LbfgsParams p,
                const int32_t *__restrict__ list, const int32_t *__restrict__ count, int32_t *cursor, long long max_rounds,
                const K11Layout lay, double *__restrict__ scratch)
{
    constexpr int J = JR + 2 * JC;
    __shared__ k9::Scratch<J> sc;
    __shared__ k9::GScratch<JR, JC, DUAL> gs;
    __shared__ LbfgsLocal loc;
    __shared__ int cur, done_flag;
    extern __shared__ double lb_dyn[];               // [3][N] series when staged, then K11's tape when it fits (mtg_k11_layout)
    double *series = lb_dyn;
    double *tape = lay.tape_smem ? lb_dyn + (lay.stage ? 3 * (size_t)a.N : 0) : scratch + (size_t)blockIdx.x * lay.scratch_doubles;
    const int tid = threadIdx.x, T = blockDim.x, nd = p.nd, m = s.m;
    const size_t B = (size_t)s.B;
    const bool dsm = (a.flags & MTG_FLAG_DS_MEAN) != 0;
    LbfgsState ls = s;
    ls.B = 1;
    ls.X = loc.X; ls.G = loc.G; ls.Dir = loc.Dir; ls.SH = loc.SH; ls.YH = loc.YH; ls.RHO = loc.RHO;
    ls.F = &loc.F; ls.alpha = &loc.alpha; ls.gd = &loc.gd; ls.XT = loc.XT; ls.GT = loc.GT; ls.LL = &loc.LL;
    ls.XC = loc.XC; ls.GC = loc.GC; ls.FC = &loc.FC;
    ls.hist_n = &loc.hist_n; ls.hist_head = &loc.hist_head; ls.ls_count = &loc.ls_count; ls.iter = &loc.iter; ls.phase = &loc.phase;
    ls.status = &loc.status; ls.restarted = &loc.restarted; ls.k2_status = &loc.k2_status; ls.flags = &loc.flags;
    ls.has_cand = &loc.has_cand; ls.n_exp = &loc.n_exp;
    int staged_ds = -1;
    for (;;) {
        __syncthreads();
        if (tid == 0) { const int i = atomicAdd(cursor, 1); cur = i < *count ? list[i] : -1; }
        __syncthreads();
        const int b = cur;
        if (b < 0) break;
        // ---- state in ----
        for (int k = tid; k < nd; k += T) {
            loc.X[k] = s.X[k * B + b]; loc.G[k] = s.G[k * B + b]; loc.Dir[k] = s.Dir[k * B + b];
            loc.XC[k] = s.XC[k * B + b]; loc.GC[k] = s.GC[k * B + b];
            loc.XT[k] = s.XT[(size_t)b * nd + k];
        }
        for (int k = tid; k < m * nd; k += T) { loc.SH[k] = s.SH[k * B + b]; loc.YH[k] = s.YH[k * B + b]; }
        for (int k = tid; k < m; k += T) loc.RHO[k] = s.RHO[k * B + b];
        if (tid == 0) {
            loc.F = s.F[b]; loc.alpha = s.alpha[b]; loc.gd = s.gd[b]; loc.FC = s.FC[b];
            loc.hist_n = s.hist_n[b]; loc.hist_head = s.hist_head[b]; loc.ls_count = s.ls_count[b]; loc.iter = s.iter[b];
            loc.phase = s.phase[b]; loc.status = s.status[b]; loc.restarted = s.restarted[b]; loc.flags = s.flags[b];
            loc.has_cand = s.has_cand[b]; loc.n_exp = s.n_exp[b];
        }
        const int ds = b / a.group;
        const double *st_t = a.t, *st_y = a.y + ds, *st_d = a.diag + ds;
        size_t stride = (size_t)a.S;
        if (lay.stage) {
            if (ds != staged_ds) { k9::stage_series(series, a.N, a.t, a.y + ds, a.diag + ds, (size_t)a.S); staged_ds = ds; }
            else __syncthreads();
            st_t = series; st_y = series + a.N; st_d = series + 2 * a.N; stride = 1;
        } else __syncthreads();
        const double dsv = dsm ? a.ds_mean[ds] : 0.0;
        // ---- the problem's remaining L-BFGS loop ----
        for (long long round = 0; round < max_rounds; ++round) {
            int st;
            const double ll = k9::series_loglike_grad<JR, JC, DUAL>(pl, loc.XT, a.N, st_t, st_y, st_d, stride, dsm, dsv,
                                                                    sc, gs, tape, &st, loc.GT);
            if (tid == 0) {
                loc.LL = ll; loc.k2_status = st;
                done_flag = (nd <= LB_SMALL ? lbfgs_advance<LB_SMALL>(ls, p, 0) : lbfgs_advance<MTG_MAX_PARAMS>(ls, p, 0)) ? 1 : 0;
            }
            __syncthreads();
            if (done_flag) break;
        }
        // ---- state out ----
        for (int k = tid; k < nd; k += T) {
            s.X[k * B + b] = loc.X[k]; s.G[k * B + b] = loc.G[k]; s.Dir[k * B + b] = loc.Dir[k];
            s.XC[k * B + b] = loc.XC[k]; s.GC[k * B + b] = loc.GC[k];
            s.XT[(size_t)b * nd + k] = loc.XT[k];
        }
        for (int k = tid; k < m * nd; k += T) { s.SH[k * B + b] = loc.SH[k]; s.YH[k * B + b] = loc.YH[k]; }
        for (int k = tid; k < m; k += T) s.RHO[k * B + b] = loc.RHO[k];
        if (tid == 0) {
            s.F[b] = loc.F; s.alpha[b] = loc.alpha; s.gd[b] = loc.gd; s.FC[b] = loc.FC;
            s.hist_n[b] = loc.hist_n; s.hist_head[b] = loc.hist_head; s.ls_count[b] = loc.ls_count; s.iter[b] = loc.iter;
            s.phase[b] = loc.phase; s.status[b] = loc.status; s.restarted[b] = loc.restarted; s.flags[b] = loc.flags;
            s.has_cand[b] = loc.has_cand; s.n_exp[b] = loc.n_exp;
        }
    }
}

template <int JR, int JC, bool DUAL>
static int launch_persist_one(const Plan &pl, const BatchArgs &a, const LbfgsState &s, const LbfgsParams &p, const int32_t *list,
                              const int32_t *count, int32_t *cursor, int bound, long long max_rounds, cudaStream_t stream)
{
    const K11Layout lay = mtg_k11_layout(JR + 2 * JC, a.N);
    auto kern = k_lbfgs_persist<JR, JC, DUAL>;
    MTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K11_SMEM_LIMIT));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = bound < sms ? bound : sms;
    MTG_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t), stream));
    double *scratch = nullptr;
    if (lay.scratch_doubles) MTG_CUDA(cudaMallocAsync((void **)&scratch, sizeof(double) * lay.scratch_doubles * (size_t)grid, stream));
    kern<<<grid, K9_THREADS, lay.dyn_bytes, stream>>>(pl, a, s, p, list, count, cursor, max_rounds, lay, scratch);
    mtg_count_launch();
    int rc = mtg_cuda_check(cudaGetLastError(), "k_lbfgs_persist launch");
    if (scratch) cudaFreeAsync(scratch, stream);
    return rc;
}
template <int JR, int JC>
static int launch_persist(const Plan &pl, const BatchArgs &a, const LbfgsState &s, const LbfgsParams &p, const int32_t *list,
                          const int32_t *count, int32_t *cursor, int bound, long long max_rounds, cudaStream_t stream)
{
    if (pl.has_sho) return launch_persist_one<JR, JC, true>(pl, a, s, p, list, count, cursor, bound, max_rounds, stream);
    return launch_persist_one<JR, JC, false>(pl, a, s, p, list, count, cursor, bound, max_rounds, stream);
}
#define LBP_CASE(R, C) if (pl.jr == R && pl.jc == C) return launch_persist<R, C>(pl, a, s, p, list, count, cursor, bound, max_rounds, stream);
static int persist_dispatch(const Plan &pl, const BatchArgs &a, const LbfgsState &s, const LbfgsParams &p, const int32_t *list,
                            const int32_t *count, int32_t *cursor, int bound, long long max_rounds, cudaStream_t stream)
{
    LBP_CASE(1, 0) LBP_CASE(2, 0) LBP_CASE(3, 0) LBP_CASE(0, 1) LBP_CASE(1, 1) LBP_CASE(2, 1) LBP_CASE(0, 2)
    mtg_set_error("persistent L-BFGS: unsupported kernel structure: %d real + %d complex slots", pl.jr, pl.jc);
    return MTG_EUNSUP;
}

// active count at and below which the remaining problems go to the persistent kernel (0 disables it)
static int lbfgs_switch_count()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("MTG_LBFGS_SWITCH"); v = e ? atoi(e) : 6144; }
    return v;
}

__global__ void k_lbfgs_finish(LbfgsState s, double *__restrict__ out_x, double *__restrict__ out_nll,
                               int32_t *__restrict__ out_iters, int32_t *__restrict__ out_status)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= s.B) return;
    for (int k = 0; k < s.nd; ++k) out_x[(size_t)b * s.nd + k] = s.X[(size_t)k * s.B + b];
    out_nll[b] = s.F[b];
    if (out_iters) out_iters[b] = s.iter[b];
    if (out_status) out_status[b] = s.status[b];        // still-active problems keep 1 (iteration limit)
}

extern "C" int mtg_lbfgs_dev(const mtg_model_t *model, int N, int S, int group, int B,
                             const double *t, const double *y, const double *diag, const double *ds_mean,
                             const double *x0, const mtg_lbfgs_opts_t *opts,
                             double *out_x, double *out_nll, int32_t *out_iters, int32_t *out_status,
                             uint32_t flags, void *stream_)
{
    if (!model || N < 1 || S < 1 || group < 1 || B < 0 || !t || !y || !diag || !x0 || !out_x || !out_nll) {
        mtg_set_error("mtg_lbfgs_dev: bad argument"); return MTG_EINVAL;
    }
    if ((flags & MTG_FLAG_DS_MEAN) && !ds_mean) { mtg_set_error("MTG_FLAG_DS_MEAN without ds_mean"); return MTG_EINVAL; }
    if (B > 0 && (long long)(B - 1) / group >= S) { mtg_set_error("B/group exceeds the number of data sets S"); return MTG_EINVAL; }
    if (B == 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    Plan pl;
    int rc = mtg_build_plan(model, &pl);
    if (rc) return rc;
    const int nd = pl.n_free;
    if (nd < 1) { mtg_set_error("mtg_lbfgs_dev: the model has no free parameters"); return MTG_EINVAL; }
    LbfgsParams p;
    p.max_iter = opts && opts->max_iter > 0 ? opts->max_iter : 200;
    p.max_ls = opts && opts->max_ls > 0 ? opts->max_ls : 20;
    p.ftol = opts && opts->ftol > 0 ? opts->ftol : 2.220446049250313e-09;
    p.gtol = opts && opts->gtol > 0 ? opts->gtol : 1e-5;
    p.max_step = 8.0;
    p.nd = nd;
    for (int k = 0; k < nd; ++k) { p.lo[k] = pl.lo[k]; p.hi[k] = pl.hi[k]; }
    int m = opts && opts->history > 0 ? opts->history : 10;
    if (m > LB_MAXM) m = LB_MAXM;

    // ---- one allocation for all state ---------------------------------------------------------
    const size_t Bz = (size_t)B;
    size_t nd_ = (size_t)nd;
    size_t n_dbl = 3 * nd_ * Bz + 2 * (size_t)m * nd_ * Bz + (size_t)m * Bz + 3 * Bz + 2 * nd_ * Bz + Bz + 2 * nd_ * Bz + Bz;
    size_t n_int = 11 * Bz + 2 * Bz + 8;
    size_t cub_bytes = 0;
    cub::CountingInputIterator<int32_t> iota(0);
    cub::DeviceSelect::Flagged((void *)nullptr, cub_bytes, iota, (int32_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, B, stream);
    const size_t total = n_dbl * 8 + n_int * 4 + cub_bytes + 1024;
    if ((rc = mtg_pool_setup())) return rc;
    char *base = nullptr;
    MTG_CUDA(cudaMallocAsync((void **)&base, total, stream));
    double *dptr = (double *)base;
    auto take_d = [&](size_t n) { double *r = dptr; dptr += n; return r; };
    LbfgsState s;
    s.B = B; s.nd = nd; s.m = m;
    s.X = take_d(nd_ * Bz); s.G = take_d(nd_ * Bz); s.Dir = take_d(nd_ * Bz);
    s.SH = take_d((size_t)m * nd_ * Bz); s.YH = take_d((size_t)m * nd_ * Bz); s.RHO = take_d((size_t)m * Bz);
    s.F = take_d(Bz); s.alpha = take_d(Bz); s.gd = take_d(Bz);
    s.XT = take_d(nd_ * Bz); s.GT = take_d(nd_ * Bz); s.LL = take_d(Bz);
    s.XC = take_d(nd_ * Bz); s.GC = take_d(nd_ * Bz); s.FC = take_d(Bz);
    int32_t *iptr = (int32_t *)dptr;
    auto take_i = [&](size_t n) { int32_t *r = iptr; iptr += n; return r; };
    s.hist_n = take_i(Bz); s.hist_head = take_i(Bz); s.ls_count = take_i(Bz); s.iter = take_i(Bz);
    s.phase = take_i(Bz); s.status = take_i(Bz); s.restarted = take_i(Bz); s.k2_status = take_i(Bz); s.flags = take_i(Bz);
    s.has_cand = take_i(Bz); s.n_exp = take_i(Bz);
    int32_t *list_a = take_i(Bz), *list_b = take_i(Bz);
    int32_t *d_count = take_i(8);           // [0], [1]: ping-pong counts, [4]: work-queue cursor of the persistent kernel
    void *cub_tmp = (void *)(((uintptr_t)iptr + 255) & ~(uintptr_t)255);

    const int chunk = mtg_k2_chunk(pl, N, B);
    double *tape = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&tape, mtg_k2_tape_bytes_per_series(pl, N) * (size_t)chunk, stream);
    if (e != cudaSuccess) { cudaFreeAsync(base, stream); return mtg_cuda_check(e, "cudaMallocAsync(tape)"); }
    // pinned word for the active count: one per host thread, kept for the life of the thread (cudaMallocHost /
    // cudaFreeHost synchronise the device, which would serialise concurrently running fits on other streams)
    static thread_local int32_t *h_count = nullptr;
    if (!h_count) {
        e = cudaMallocHost((void **)&h_count, sizeof(int32_t));
        if (e != cudaSuccess) { h_count = nullptr; cudaFreeAsync(tape, stream); cudaFreeAsync(base, stream); return mtg_cuda_check(e, "cudaMallocHost"); }
    }

    const dim3 blk(LB_THREADS);
    k_lbfgs_init<<<(B + LB_THREADS - 1) / LB_THREADS, blk, 0, stream>>>(s, p, x0, list_a);
    *h_count = B;
    cudaMemcpyAsync(d_count, h_count, sizeof(int32_t), cudaMemcpyHostToDevice, stream);
    mtg_count_launch();

    const bool debug = getenv("MTG_LBFGS_DEBUG") != nullptr;
    int bound = B;                          // host-side upper bound on the active count
    int32_t *list = list_a, *list_next = list_b;
    int32_t *cnt = d_count, *cnt_next = d_count + 1;
    const long long max_rounds = (long long)p.max_iter * (p.max_ls + LB_MAX_EXPAND + 1) + 2;
    rc = 0;
    const bool can_persist = lbfgs_switch_count() > 0 && mtg_k11_supports(pl, N) && N >= 64;
    for (long long round = 0; round < max_rounds && bound > 0; ++round) {
        BatchArgs a{N, S, group, bound, t, y, diag, ds_mean, s.XT, s.LL, s.GT, s.k2_status, flags, list, cnt, nullptr};
        if (can_persist && bound <= lbfgs_switch_count()) {
            // the tail (or a small batch from its first evaluation on): one launch runs every remaining problem to completion
            if (debug) fprintf(stderr, "[mtg_lbfgs] B=%d round %lld: %d active -> persistent kernel\n", B, round, bound);
            rc = persist_dispatch(pl, a, s, p, list, cnt, d_count + 4, bound, max_rounds - round, stream);
            break;
        }
        if ((rc = mtg_k2_run(pl, a, tape, chunk, stream))) break;
        k_lbfgs_update<<<(bound + LB_THREADS - 1) / LB_THREADS, blk, 0, stream>>>(s, p, list, cnt);
        // ordered compaction of the still-active problems
        size_t tmp_bytes = cub_bytes;
        cub::DeviceSelect::Flagged(cub_tmp, tmp_bytes, iota, s.flags, list_next, cnt_next, B, stream);
        mtg_count_launch(2);
        { int32_t *sw = list; list = list_next; list_next = sw; sw = cnt; cnt = cnt_next; cnt_next = sw; }
        if ((round & 3) == 3 || bound <= 4096) {
            cudaMemcpyAsync(h_count, cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, stream);
            if ((rc = mtg_cuda_check(cudaStreamSynchronize(stream), "lbfgs round sync"))) break;
            bound = *h_count;
            if (debug) fprintf(stderr, "[mtg_lbfgs] B=%d round %lld active %d\n", B, round + 1, bound);
        }
    }
    if (!rc) {
        k_lbfgs_finish<<<(B + LB_THREADS - 1) / LB_THREADS, blk, 0, stream>>>(s, out_x, out_nll, out_iters, out_status);
        mtg_count_launch();
        rc = mtg_cuda_check(cudaGetLastError(), "lbfgs finish");
    }
    cudaFreeAsync(tape, stream);
    cudaFreeAsync(base, stream);
    if (!rc) rc = mtg_cuda_check(cudaStreamSynchronize(stream), "lbfgs final sync");
    return rc;
}
