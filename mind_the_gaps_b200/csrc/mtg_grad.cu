// mtg_grad.cu -- K2: log-likelihood AND its gradient in one kernel (forward + reverse sweep).
//
// The reference never forms a gradient: GPModelling.fit hands scipy's L-BFGS-B no `jac`
// (mind_the_gaps/gpmodelling.py:192-193) so every optimiser iteration costs ndim+1 likelihood
// evaluations by finite differences.  Here one thread runs the celerite recursion forward, writing
// the per-step state (S, f, W, D, z) to a time-major scratch tape [N][K][Bc] (coalesced), and then
// sweeps the tape backwards accumulating the adjoints of the celerite coefficients, which are
// chained to the log-parameters on the device (plan_chain).  exp / sincos are recomputed in the
// reverse sweep rather than taped: the reverse sweep is HBM-bound on the tape, so the FP64 pipe has
// the slack.
//
// Reverse-mode derivation (adjoints written with a leading "b"): for step n with previous state
// (S', f', W', D', z'), p_j = exp(-c_j dt_n), A = S' + D' W' W'^T:
//   S = (p p^T) o A, f = p o (f' + W' z'), su = S u, D = d_n + sum a - u.su, W = (v - su)/D,
//   z = r_n - u.f, l += -1/2 (z^2/D + log D)
// gives
//   bz += -z/D ; bD += 1/2 z^2/D^2 - 1/2 /D - (bW.W)/D ; bsu = -bW/D - bD u ; bv = bW/D
//   bu = -bz f - bD su + S bsu ; bf += -bz u ; bS_jk += bsu_j u_k (+ sym.)
//   c_j : bc_j += -dt (sum_k bS_jk S_jk [diag twice] + bf_j f_j)
//   bA = (p p^T) o bS -> bS' = bA, bD' = W'^T bA W', bW' = 2 D' bA W' + bg z', bz' = bg.W', bf' = bg,
//   with bg = p o bf.
#include "mtg_recursion.cuh"

#define K2_THREADS 128

// coefficient adjoints
template <int JR, int JC, bool DUAL>
struct CoefAdj {
    double ar[JR > 0 ? JR : 1], cr[JR > 0 ? JR : 1];
    double ac[JC > 0 ? JC : 1], bc[JC > 0 ? JC : 1], cc[JC > 0 ? JC : 1], dc[JC > 0 ? JC : 1];
    double a2[(DUAL && JC > 0) ? JC : 1], c2[(DUAL && JC > 0) ? JC : 1];
    double a_sum;
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int j = 0; j < JR; ++j) { ar[j] = 0.0; cr[j] = 0.0; }
#pragma unroll
        for (int j = 0; j < JC; ++j) { ac[j] = bc[j] = cc[j] = dc[j] = 0.0; if (DUAL) { a2[j] = 0.0; c2[j] = 0.0; } }
        a_sum = 0.0;
    }
};

// d loglike / d free parameter from the coefficient adjoints (chain rule through plan_coefficients)
template <int JR, int JC, bool DUAL>
__device__ __forceinline__ void plan_chain(const Plan &pl, const double *__restrict__ x,
                                           const Coef<JR, JC, DUAL> &c, const CoefAdj<JR, JC, DUAL> &g,
                                           double *__restrict__ grad /* [n_free], pre-zeroed */)
{
    auto add = [&](int full_idx, double v) { const int pos = pl.src_pos[full_idx]; if (pos >= 0) grad[pos] += v; };
    for (int k = 0; k < pl.n_jit; ++k)
        add(pl.jit_idx[k], g.a_sum * 2.0 * exp(2.0 * plan_param(pl, x, pl.jit_idx[k])));
#pragma unroll
    for (int j = 0; j < JR; ++j) {
        add(pl.r_ia[j], (g.ar[j] + g.a_sum) * c.ar[j]);
        add(pl.r_ic[j], g.cr[j] * (-c.ncr[j] * MTG_LN2));
    }
#pragma unroll
    for (int j = 0; j < JC; ++j) {
        const double a = c.ac[j], b = c.bc[j], cc = -c.ncc[j] * MTG_LN2, d = c.dc[j];
        const double ga = g.ac[j] + g.a_sum, gb = g.bc[j], gc = g.cc[j], gd = g.dc[j];
        const int *ix = pl.c_idx[j];
        switch (pl.c_kind[j]) {
        case CS_COMPLEX3: add(ix[0], ga * a); add(ix[1], gc * cc); add(ix[2], gd * d); break;
        case CS_COMPLEX4: add(ix[0], ga * a); add(ix[1], gb * b); add(ix[2], gc * cc); add(ix[3], gd * d); break;
        case CS_LORENTZIAN: add(ix[0], ga * a); add(ix[1], -gc * cc); add(ix[2], gc * cc + gd * d); break;
        case CS_COSINUS: add(ix[0], ga * a); add(ix[1], gd * d); break;
        case CS_BENDINGPL: add(ix[0], ga * a); add(ix[1], gb * b); add(ix[2], (gc + gd) * d); break;
        case CS_MATERN32: add(ix[0], 2.0 * (ga * a + gb * b)); add(ix[1], -(gb * b + gc * cc)); break;
        default: /* CS_SHO */ {
            const double Q = exp(plan_param(pl, x, ix[1]));
            if (!DUAL || c.pair[j]) {
                const double f2 = 4.0 * Q * Q - 1.0, r = 4.0 * Q * Q / f2;
                add(ix[0], ga * a + gb * b);
                add(ix[1], ga * a + gb * b * (1.0 - r) - gc * cc + gd * d * (r - 1.0));
                add(ix[2], ga * a + gb * b + gc * cc + gd * d);
            } else {
                const double a2 = c.a2[j], c2 = -c.nc2[j] * MTG_LN2;
                const double ga2 = g.a2[j] + g.a_sum, gc2 = g.c2[j];
                const double f = sqrt(1.0 - 4.0 * Q * Q);
                const double h = 0.5 * (a + a2);                       // = 1/2 S0 w0 Q
                const double da = h * 4.0 * Q * Q / (f * f * f);
                const double w0 = exp(plan_param(pl, x, ix[2]));
                const double dcq = 2.0 * w0 * Q / f;
                add(ix[0], ga * a + ga2 * a2);
                add(ix[1], ga * (a + da) + ga2 * (a2 - da) + gc * (-cc + dcq) + gc2 * (-c2 - dcq));
                add(ix[2], ga * a + ga2 * a2 + gc * cc + gc2 * c2);
            } } break;
        }
    }
}

template <int JR, int JC, bool DUAL>
__global__ void __launch_bounds__(K2_THREADS)
k2_loglike_grad(const __grid_constant__ Plan pl, const BatchArgs a, double *__restrict__ tape, int b0, int Bc)
{
    constexpr int J = JR + 2 * JC;
    constexpr int NS = J * (J + 1) / 2;
    constexpr int K = NS + 2 * J + 2;                       // tape doubles per step: S, f, W, D, z
    __shared__ alignas(16) double tabs_sm[MTG_TABS_DOUBLES];
    const mtgm::Tabs tb = mtgm::tabs_fill(tabs_sm);
    __syncthreads();
    const int lb = blockIdx.x * K2_THREADS + threadIdx.x;   // index inside the chunk
    if (lb >= Bc) return;
    if (a.count && b0 + lb >= *a.count) return;
    const int b = a.index ? a.index[b0 + lb] : b0 + lb;
    const int N = a.N;
    const double *x = a.params + (size_t)b * pl.n_free;
    Coef<JR, JC, DUAL> c;
    plan_coefficients<JR, JC, DUAL>(pl, x, c);
    const bool inside = !(a.flags & MTG_FLAG_APPLY_PRIOR) || plan_in_prior(pl, x);
    const int ds = b / a.group;
    const bool ds_mean = (a.flags & MTG_FLAG_DS_MEAN) != 0;
    MeanEval mean = plan_mean(pl, x);
    if (ds_mean) { mean.kind = MTG_MEAN_CONSTANT; mean.p0 = a.ds_mean[ds]; }
    const double *yp = a.y + ds, *dp = a.diag + ds;
    const size_t S = (size_t)a.S;
    double *tp = tape + lb;
    const size_t Bs = (size_t)Bc;

    // ---------------- forward sweep, taping the state after every step -------------------------
    K1State<JR, JC, DUAL> s;
    s.init(__ldg(a.t));
    double t_nx = __ldg(a.t), y_nx = __ldg(yp), d_nx = __ldg(dp);
#pragma unroll 1
    for (int n = 0; n < N; ++n) {
        const double tn = t_nx, yn = y_nx, dn = d_nx;
        if (n + 1 < N) { t_nx = __ldg(a.t + n + 1); y_nx = __ldg(yp + (size_t)(n + 1) * S); d_nx = __ldg(dp + (size_t)(n + 1) * S); }
        s.step(c, tb, tn, yn - mean(tn), dn, n);
        double *rec = tp + (size_t)n * K * Bs;
#pragma unroll
        for (int i = 0; i < NS; ++i) __stcs(rec + (size_t)i * Bs, s.S[i]);
#pragma unroll
        for (int j = 0; j < J; ++j) { __stcs(rec + (size_t)(NS + j) * Bs, s.f[j]); __stcs(rec + (size_t)(NS + J + j) * Bs, s.W[j] * s.rDp); }
        __stcs(rec + (size_t)(NS + 2 * J) * Bs, s.Dp);
        __stcs(rec + (size_t)(NS + 2 * J + 1) * Bs, s.zp);
    }
    int stt;
    double ll = s.loglike(N, &stt);
    if (!inside) { ll = -INFINITY; stt = 3; }
    a.out_ll[b] = ll;
    if (a.status) a.status[b] = stt;
    double *grad = a.out_grad + (size_t)b * pl.n_free;
    for (int k = 0; k < pl.n_free; ++k) grad[k] = 0.0;
    if (stt != 0) return;

    // ---------------- reverse sweep -----------------------------------------------------------
    double bS[NS > 0 ? NS : 1], bf[J > 0 ? J : 1], bW[J > 0 ? J : 1], bD = 0.0, bz = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) bS[i] = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) { bf[j] = 0.0; bW[j] = 0.0; }
    CoefAdj<JR, JC, DUAL> g; g.zero();
    double gm0 = 0.0, gm1 = 0.0, gm2 = 0.0, gm3 = 0.0;       // mean-parameter gradients

    // current record (step n) in registers; previous record loaded while step n is processed
    double cS[NS > 0 ? NS : 1], cf[J > 0 ? J : 1], cW[J > 0 ? J : 1], cD, cz;
    {
        const double *rec = tp + (size_t)(N - 1) * K * Bs;
#pragma unroll
        for (int i = 0; i < NS; ++i) cS[i] = __ldcs(rec + (size_t)i * Bs);
#pragma unroll
        for (int j = 0; j < J; ++j) { cf[j] = __ldcs(rec + (size_t)(NS + j) * Bs); cW[j] = __ldcs(rec + (size_t)(NS + J + j) * Bs); }
        cD = __ldcs(rec + (size_t)(NS + 2 * J) * Bs); cz = __ldcs(rec + (size_t)(NS + 2 * J + 1) * Bs);
    }
    double tn = __ldg(a.t + N - 1);
#pragma unroll 1
    for (int n = N - 1; n >= 0; --n) {
        double pS[NS > 0 ? NS : 1], pf[J > 0 ? J : 1], pW[J > 0 ? J : 1], pD = 0.0, pz = 0.0, tprev = tn;
        if (n > 0) {
            const double *rec = tp + (size_t)(n - 1) * K * Bs;
#pragma unroll
            for (int i = 0; i < NS; ++i) pS[i] = __ldcs(rec + (size_t)i * Bs);
#pragma unroll
            for (int j = 0; j < J; ++j) { pf[j] = __ldcs(rec + (size_t)(NS + j) * Bs); pW[j] = __ldcs(rec + (size_t)(NS + J + j) * Bs); }
            pD = __ldcs(rec + (size_t)(NS + 2 * J) * Bs); pz = __ldcs(rec + (size_t)(NS + 2 * J + 1) * Bs);
            tprev = __ldg(a.t + n - 1);
        } else {
#pragma unroll
            for (int i = 0; i < NS; ++i) pS[i] = 0.0;
#pragma unroll
            for (int j = 0; j < J; ++j) { pf[j] = 0.0; pW[j] = 0.0; }
        }
        const double dt = tn - tprev;
        // recompute p, u, v for step n
        double p[J > 0 ? J : 1], u[J > 0 ? J : 1], v[J > 0 ? J : 1];
        double cdv[JC > 0 ? JC : 1], sdv[JC > 0 ? JC : 1];
#pragma unroll
        for (int j = 0; j < JR; ++j) { p[j] = mtgm::exp2_nonpos_tab(c.ncr[j] * dt, tb); u[j] = c.ar[j]; v[j] = 1.0; }
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const double e = mtgm::exp2_nonpos_tab(c.ncc[j] * dt, tb);
            if (!DUAL || c.pair[j]) {
                double sd, cd; mtgm::sincos_tab(c.dc[j] * tn, tb, &sd, &cd);
                cdv[j] = cd; sdv[j] = sd;
                p[JR + 2 * j] = e; p[JR + 2 * j + 1] = e;
                u[JR + 2 * j] = fma(c.ac[j], cd, c.bc[j] * sd); u[JR + 2 * j + 1] = fma(c.ac[j], sd, -c.bc[j] * cd);
                v[JR + 2 * j] = cd; v[JR + 2 * j + 1] = sd;
            } else {
                cdv[j] = 1.0; sdv[j] = 0.0;
                p[JR + 2 * j] = e; p[JR + 2 * j + 1] = mtgm::exp2_nonpos_tab(c.nc2[j] * dt, tb);
                u[JR + 2 * j] = c.ac[j]; u[JR + 2 * j + 1] = c.a2[j];
                v[JR + 2 * j] = 1.0; v[JR + 2 * j + 1] = 1.0;
            }
        }
        const double rD = mtgm::rcp_pos(cD);
        // local likelihood terms
        bz = fma(-cz, rD, bz);
        double bWW = 0.0;
#pragma unroll
        for (int j = 0; j < J; ++j) bWW = fma(bW[j], cW[j], bWW);
        bD += (0.5 * cz * cz * rD - 0.5 - bWW) * rD;
        // mean parameters: r_n = y_n - mean(t_n), br_n = bz
        if (!ds_mean) {
            if (mean.kind == MTG_MEAN_CONSTANT) gm0 -= bz;
            else if (mean.kind == MTG_MEAN_LINEAR) { gm0 = fma(-bz, tn, gm0); gm1 -= bz; }
            else {
                const double q = tn - mean.p0, s2 = mean.p1 * mean.p1;
                const double ex = exp(-q * q / (2.0 * s2)), amp = mean.p2 / (2.0 * M_PI * mean.p1);
                gm0 -= bz * amp * ex * q / s2;
                gm1 -= bz * amp * ex * (q * q / (s2 * mean.p1) - 1.0 / mean.p1);
                gm2 -= bz * ex / (2.0 * M_PI * mean.p1);
                gm3 -= bz;
            }
        }
        // su = v - W D ; bsu = -bW/D - bD u ; bv = bW/D
        double su[J > 0 ? J : 1], bsu[J > 0 ? J : 1], bu[J > 0 ? J : 1], bv[J > 0 ? J : 1];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            su[j] = fma(-cW[j], cD, v[j]);
            bv[j] = bW[j] * rD;
            bsu[j] = fma(-bD, u[j], -bv[j]);
            bu[j] = fma(-bz, cf[j], -bD * su[j]);
            bf[j] = fma(-bz, u[j], bf[j]);
        }
        g.a_sum += bD;
        if (a.per_step) {       // d loglike / d r_n = -alpha_n ;  d loglike / d diag_n = 1/2 alpha_n^2 - 1/2 (K^-1)_nn
            a.per_step[(size_t)n * a.B + b] = -bz;
            a.per_step[((size_t)N + n) * a.B + b] = fma(bz, bz, -2.0 * bD);
        }
        // bu += S bsu ; bS += bsu u^T (symmetrised)
#pragma unroll
        for (int j = 0; j < J; ++j) {
            double acc = bu[j];
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(cS[j <= k ? tri<J>(j, k) : tri<J>(k, j)], bsu[k], acc);
            bu[j] = acc;
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            bS[tri<J>(j, j)] = fma(bsu[j], u[j], bS[tri<J>(j, j)]);
#pragma unroll
            for (int k = j + 1; k < J; ++k)
                bS[tri<J>(j, k)] += fma(bsu[j], u[k], bsu[k] * u[j]);
        }
        // decay rates: bc_j += -dt * (sum_k bS_jk S_jk [diagonal twice] + bf_j f_j)
        double pc[J > 0 ? J : 1];
#pragma unroll
        for (int j = 0; j < J; ++j) pc[j] = bf[j] * cf[j];
#pragma unroll
        for (int j = 0; j < J; ++j) {
#pragma unroll
            for (int k = j; k < J; ++k) {
                const double w = bS[tri<J>(j, k)] * cS[tri<J>(j, k)];
                pc[j] += w; pc[k] += w;
            }
        }
#pragma unroll
        for (int j = 0; j < JR; ++j) g.cr[j] = fma(-dt, pc[j], g.cr[j]);
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            if (!DUAL || c.pair[j]) g.cc[j] = fma(-dt, pc[JR + 2 * j] + pc[JR + 2 * j + 1], g.cc[j]);
            else { g.cc[j] = fma(-dt, pc[JR + 2 * j], g.cc[j]); g.c2[j] = fma(-dt, pc[JR + 2 * j + 1], g.c2[j]); }
        }
        // u, v -> a, b, d
#pragma unroll
        for (int j = 0; j < JR; ++j) g.ar[j] += bu[j];
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const double bu0 = bu[JR + 2 * j], bu1 = bu[JR + 2 * j + 1];
            if (!DUAL || c.pair[j]) {
                const double cd = cdv[j], sd = sdv[j];
                g.ac[j] += fma(bu0, cd, bu1 * sd);
                g.bc[j] += fma(bu0, sd, -bu1 * cd);
                const double bcd = fma(bu0, c.ac[j], -bu1 * c.bc[j]) + bv[JR + 2 * j];
                const double bsd = fma(bu0, c.bc[j], bu1 * c.ac[j]) + bv[JR + 2 * j + 1];
                g.dc[j] = fma(tn, fma(cd, bsd, -sd * bcd), g.dc[j]);
            } else { g.ac[j] += bu0; g.a2[j] += bu1; }
        }
        // propagate to step n-1:  bg = p o bf ; bA = (p p^T) o bS
        double bg[J > 0 ? J : 1];
#pragma unroll
        for (int j = 0; j < J; ++j) bg[j] = p[j] * bf[j];
        double nbD = 0.0, nbz = 0.0, nbW[J > 0 ? J : 1];
#pragma unroll
        for (int j = 0; j < J; ++j) { nbW[j] = bg[j] * pz; nbz = fma(bg[j], pW[j], nbz); }
#pragma unroll
        for (int j = 0; j < J; ++j) {
#pragma unroll
            for (int k = j; k < J; ++k) {
                const double bA = (p[j] * p[k]) * bS[tri<J>(j, k)];
                bS[tri<J>(j, k)] = bA;
                nbD = fma(bA * pW[j], pW[k], nbD);
                const double m = (j == k) ? 2.0 * bA : bA;
                nbW[j] = fma(pD * m, pW[k], nbW[j]);
                if (k != j) nbW[k] = fma(pD * bA, pW[j], nbW[k]);
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) { bf[j] = bg[j]; bW[j] = nbW[j]; cf[j] = pf[j]; cW[j] = pW[j]; }
#pragma unroll
        for (int i = 0; i < NS; ++i) cS[i] = pS[i];
        bD = nbD; bz = nbz; cD = pD; cz = pz; tn = tprev;
    }

    plan_chain<JR, JC, DUAL>(pl, x, c, g, grad);
    if (!ds_mean) {
        auto addm = [&](int i, double v) { const int pos = pl.src_pos[pl.mean_idx[i]]; if (pos >= 0) grad[pos] += v; };
        addm(0, gm0);
        if (mean.kind != MTG_MEAN_CONSTANT) addm(1, gm1);
        if (mean.kind == MTG_MEAN_GAUSSIAN) { addm(2, gm2); addm(3, gm3); }
    }
}

// ---- dispatch ---------------------------------------------------------------------------------
template <int JR, int JC>
static int launch_k2(const Plan &pl, const BatchArgs &a, double *tape, int b0, int Bc, cudaStream_t stream)
{
    const dim3 grid((Bc + K2_THREADS - 1) / K2_THREADS), block(K2_THREADS);
    if (pl.has_sho) k2_loglike_grad<JR, JC, true><<<grid, block, 0, stream>>>(pl, a, tape, b0, Bc);
    else            k2_loglike_grad<JR, JC, false><<<grid, block, 0, stream>>>(pl, a, tape, b0, Bc);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k2_loglike_grad launch");
}

#define K2_CASE(R, C) if (pl.jr == R && pl.jc == C) return launch_k2<R, C>(pl, a, tape, b0, Bc, stream);

static int k2_dispatch(const Plan &pl, const BatchArgs &a, double *tape, int b0, int Bc, cudaStream_t stream)
{
    K2_CASE(1, 0) K2_CASE(2, 0) K2_CASE(3, 0)
    K2_CASE(0, 1) K2_CASE(1, 1) K2_CASE(2, 1) K2_CASE(3, 1)
    K2_CASE(0, 2) K2_CASE(1, 2) K2_CASE(2, 2)
    K2_CASE(0, 3) K2_CASE(1, 3) K2_CASE(2, 3)
    K2_CASE(0, 4) K2_CASE(1, 4)
    mtg_set_error("unsupported kernel structure: %d real + %d complex recursion slots", pl.jr, pl.jc);
    return MTG_EUNSUP;
}

// tape bytes for one series
size_t mtg_k2_tape_bytes_per_series(const Plan &pl, int N)
{
    const int J = pl.jr + 2 * pl.jc;
    return sizeof(double) * (size_t)N * (size_t)(J * (J + 1) / 2 + 2 * J + 2);
}

// series per chunk such that the tape fits the scratch budget (MTG_TAPE_MB env, default 24 GiB)
int mtg_k2_chunk(const Plan &pl, int N, int B)
{
    size_t budget = (size_t)24 << 30;
    if (const char *e = getenv("MTG_TAPE_MB")) { long v = atol(e); if (v > 0) budget = (size_t)v << 20; }
    size_t per = mtg_k2_tape_bytes_per_series(pl, N);
    size_t n = budget / per;
    n = n / K2_THREADS * K2_THREADS;
    if (n < K2_THREADS) n = K2_THREADS;
    return (int)(n < (size_t)B ? n : (size_t)B);
}

int mtg_k2_run(const Plan &pl, const BatchArgs &a, double *tape, int chunk, cudaStream_t stream)
{
    for (int b0 = 0; b0 < a.B; b0 += chunk) {
        const int Bc = a.B - b0 < chunk ? a.B - b0 : chunk;
        int rc = k2_dispatch(pl, a, tape, b0, Bc, stream);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int mtg_loglike_grad_dev(const mtg_model_t *model, int N, int S, int group, int B,
                                    const double *t, const double *y, const double *diag, const double *ds_mean,
                                    const double *params, double *out_ll, double *out_grad, int32_t *status,
                                    uint32_t flags, void *stream_)
{
    if (!model || N < 1 || S < 1 || group < 1 || B < 0 || !t || !y || !diag || !params || !out_ll || !out_grad) {
        mtg_set_error("mtg_loglike_grad_dev: bad argument"); return MTG_EINVAL;
    }
    if ((flags & MTG_FLAG_DS_MEAN) && !ds_mean) { mtg_set_error("MTG_FLAG_DS_MEAN without ds_mean"); return MTG_EINVAL; }
    if (B > 0 && (long long)(B - 1) / group >= S) { mtg_set_error("B/group exceeds the number of data sets S"); return MTG_EINVAL; }
    if (B == 0) return 0;
    cudaStream_t stream = (cudaStream_t)stream_;
    Plan pl;
    int rc = mtg_build_plan(model, &pl);
    if (rc) return rc;
    if ((rc = mtg_pool_setup())) return rc;
    const int chunk = mtg_k2_chunk(pl, N, B);
    double *tape = nullptr;
    MTG_CUDA(cudaMallocAsync((void **)&tape, mtg_k2_tape_bytes_per_series(pl, N) * (size_t)chunk, stream));
    BatchArgs a{N, S, group, B, t, y, diag, ds_mean, params, out_ll, out_grad, status, flags, nullptr, nullptr, nullptr};
    rc = mtg_k2_run(pl, a, tape, chunk, stream);
    cudaError_t e = cudaFreeAsync(tape, stream);
    if (rc) return rc;
    return mtg_cuda_check(e, "cudaFreeAsync");
}
