// mtg_recursion.cuh -- the celerite semiseparable recursion state shared by K1 (log-likelihood),
// K2 (gradient) and the L-BFGS driver.  One thread owns one series; everything here lives in
// registers (fully unrolled, compile-time J).
//
// Algorithm: Foreman-Mackey et al. 2017, sec. 5 (celerite CholeskySolver::compute / solve /
// log_determinant, un-vendored dependency of /root/reference, reached from
// mind_the_gaps/gpmodelling.py:54,152,169); restated in SURVEY.md 3.1.
#pragma once
#include "mtg_common.cuh"
#include "mtg_math.cuh"

template <int J> __device__ __forceinline__ constexpr int tri(int j, int k)   // j <= k, upper triangle
{ return j * J - (j * (j - 1)) / 2 + (k - j); }

template <int JR, int JC, bool DUAL>
struct K1State {
    static constexpr int J = JR + 2 * JC;
    static constexpr int NS = J * (J + 1) / 2;
    double S[NS > 0 ? NS : 1];
    double f[J > 0 ? J : 1], W[J > 0 ? J : 1];   // W is kept UNNORMALISED: v - S u (celerite's W times D)
    double Dp, rDp, zp, tp;     // previous D, 1/D, z, t
    double chi2, prod; int esum; bool bad;

    __device__ __forceinline__ void init(double t0)
    {
#pragma unroll
        for (int i = 0; i < NS; ++i) S[i] = 0.0;
#pragma unroll
        for (int j = 0; j < J; ++j) { f[j] = 0.0; W[j] = 0.0; }
        Dp = 0.0; rDp = 0.0; zp = 0.0; tp = t0; chi2 = 0.0; prod = 1.0; esum = 0; bad = false;
    }

    // The per-step quantities that do not depend on the recursion state: phi = exp(-c dt) and the generator rows u, v.
    // K1 computes them one step AHEAD of the state update (software pipelining): the ~40 FP64 instructions of the
    // transcendentals then fill the dependency bubbles of the previous step's S -> D -> 1/D -> W chain.
    struct Pre { double p[J > 0 ? J : 1], u[J > 0 ? J : 1], v[J > 0 ? J : 1]; };

    template <class TB>
    __device__ __forceinline__ void prepare(const Coef<JR, JC, DUAL> &c, const TB &tb, double tn, double tprev, Pre &q) const
    {
        const double dt = tn - tprev;
        double *p = q.p, *u = q.u, *v = q.v;
#pragma unroll
        for (int j = 0; j < JR; ++j) { p[j] = mtgm::exp2_nonpos_tab(c.ncr[j] * dt, tb); u[j] = c.ar[j]; v[j] = 1.0; }
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const double e = mtgm::exp2_nonpos_tab(c.ncc[j] * dt, tb);
            if (!DUAL || c.pair[j]) {
                double sd, cd;
                mtgm::sincos_tab(c.dc[j] * tn, tb, &sd, &cd);
                p[JR + 2 * j] = e; p[JR + 2 * j + 1] = e;
                if (c.b_zero) { u[JR + 2 * j] = c.ac[j] * cd; u[JR + 2 * j + 1] = c.ac[j] * sd; }
                else {
                    u[JR + 2 * j] = fma(c.ac[j], cd, c.bc[j] * sd);
                    u[JR + 2 * j + 1] = fma(c.ac[j], sd, -c.bc[j] * cd);
                }
                v[JR + 2 * j] = cd; v[JR + 2 * j + 1] = sd;
            } else {
                p[JR + 2 * j] = e; p[JR + 2 * j + 1] = mtgm::exp2_nonpos_tab(c.nc2[j] * dt, tb);
                u[JR + 2 * j] = c.ac[j]; u[JR + 2 * j + 1] = c.a2[j];
                v[JR + 2 * j] = 1.0; v[JR + 2 * j + 1] = 1.0;
            }
        }
    }

    // one observation: (tn, residual rn, diagonal dn).  With the zero-initialised state and
    // tp = t0 the same code performs the n = 0 start-up (S = 0, f = 0, D = diag + sum a).
    // TB: mtgm::Tabs (256-entry tables; K2, K8) or mtgm::TabsBig (2048-entry tables; K1)
    template <class TB>
    __device__ __forceinline__ void step(const Coef<JR, JC, DUAL> &c, const TB &tb, double tn, double rn, double dn, int n)
    {
        Pre q;
        prepare(c, tb, tn, tp, q);
        tp = tn;
        advance(c, q, rn, dn, n);
    }

    __device__ __forceinline__ void advance(const Coef<JR, JC, DUAL> &c, const Pre &q, double rn, double dn, int n)
    {
        const double *p = q.p, *u = q.u, *v = q.v;
        // S <- (phi phi^T) o (S + D W W^T)   [previous D, W];   f <- phi o (f + W z).  With the
        // unnormalised W' = D W:  D W W^T = W' (W'/D)^T and W z = W' (z/D); chi^2 picks up z^2/D here too.
        const double zr = zp * rDp;
        chi2 = fma(zp, zr, chi2);
        double wr[J > 0 ? J : 1];
#pragma unroll
        for (int j = 0; j < J; ++j) wr[j] = W[j] * rDp;
#pragma unroll
        for (int j = 0; j < J; ++j) {
#pragma unroll
            for (int k = j; k < J; ++k)
                S[tri<J>(j, k)] = (p[j] * p[k]) * fma(W[j], wr[k], S[tri<J>(j, k)]);
            f[j] = p[j] * fma(W[j], zr, f[j]);
        }
        // D = diag + sum a - u^T S u ;  W = (v - S u)/D ;  z = r - u^T f
        double D = dn + c.a_sum, z = rn;
        double su[J > 0 ? J : 1];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(S[j <= k ? tri<J>(j, k) : tri<J>(k, j)], u[k], acc);
            su[j] = acc;
        }
#pragma unroll
        for (int j = 0; j < J; ++j) { D = fma(-u[j], su[j], D); z = fma(-u[j], f[j], z); }
#pragma unroll
        for (int j = 0; j < J; ++j) W[j] = v[j] - su[j];
        bad = bad || (__double2hiint(D) <= 0);          // D <= 0 (or denormal): not positive definite; NaN surfaces in the final isfinite test
        // sum log D as the log of a running product, renormalised every 4 steps
        prod *= D;
        if ((n & 3) == 3) {                      // frexp on a positive normal number, branch free
            const int hi = __double2hiint(prod);
            esum += ((hi >> 20) & 0x7ff) - 1022;
            prod = __hiloint2double((hi & 0x800fffff) | 0x3fe00000, __double2loint(prod));
        }
        Dp = D; rDp = mtgm::rcp_pos(D); zp = z;
    }

    __device__ __forceinline__ double loglike(int N, int *st) const
    {
        const double logdet = log(prod) + (double)esum * 0.6931471805599453094;
        const double chi2_all = fma(zp * zp, rDp, chi2);             // the last step's z^2/D
        double ll = -0.5 * (chi2_all + logdet + (double)N * 1.8378770664093454836);
        int s = 0;
        if (bad) s = 1; else if (!isfinite(ll)) s = 2;
        if (s) ll = -INFINITY;
        *st = s;
        return ll;
    }
};

