// mtg_scan_grad.cuh -- K11: log-likelihood AND gradient of ONE series by a whole CTA, parallel in time.
//
// K2 (mtg_grad.cu) differentiates the celerite recursion with one thread per series: three dependent sweeps over the N
// observations, ~1.1 ms at N = 1000 whatever the batch -- the latency every round of the batched L-BFGS pays once only a
// few hundred fits are still running (GPModelling.fit, mind_the_gaps/gpmodelling.py:172-194, as used for every simulated
// lightcurve of the LRT, docs/notebooks/tutorial_ppp.ipynb cells 11-13).  Here the reverse sweep is parallelised as well.
//
// Forward: K9's scan (mtg_scan.cuh) gives every thread the filtered state (m, P) before its chunk of CH = ceil(N / 256)
// observations; the thread filters its chunk, keeps (m, P) before every step on a shared-memory tape, and accumulates the
// chunk's closed-loop transition and backward information
//     Phi = prod_n (I - K_n h^T) F_n,     q = sum_n w_n v_n / s_n,     Omega = sum_n w_n w_n^T / s_n,    w_n = Psi_{n-1}^T F_n^T h
// (exactly what `append` builds as (A, eta, J) when it starts from the thread's true prior instead of a point mass).
// Reverse: with lambda, Lambda the adjoints of a filtered state (d loglike = lambda . dm + sum_ij Lambda_ij dP_ij), one
// filter step  m- = F m, P- = F (P - Pinf) F^T + Pinf, g = P- h, s = h.g + R, v = y - h.m-, m+ = m- + g v / s,
// P+ = P- - g g^T / s, l = -1/2 (v^2 / s + log s)  has the adjoint
//     vb = (lambda.g - v) / s;   sb = (g^T Lambda g - (lambda.g) v) / s^2 + v^2 / (2 s^2) - 1 / (2 s)
//     gb = lambda v / s - 2 Lambda g / s + sb h;   mb = lambda - vb h;   Pb = Lambda + sym(gb h^T)
//     lambda' = F^T mb;   Lambda' = F^T Pb F;   Pinf_b += Pb - Lambda';   F_b += 2 Pb F (P - Pinf) + mb m^T;   R_b += sb
// and over a whole chunk this map is affine with a closed form in the chunk's (Phi, q, Omega):
//     lambda_start = Phi^T lambda_end + q,    Lambda_start = Phi^T Lambda_end Phi + sym((Phi^T lambda_end) q^T) + (q q^T - Omega) / 2
// (Lambda - lambda lambda^T / 2 decouples: it is the backward information filter).  Chunks compose associatively,
//     (a then b) = (Phi_b Phi_a,  Phi_a^T q_b + q_a,  Phi_a^T Omega_b Phi_a + Omega_a),
// so ONE suffix scan over the 256 threads (no inverse in the operator) hands every thread the adjoint of its chunk's end
// state, lambda = q_suf, Lambda = (q_suf q_suf^T - Omega_suf) / 2, and the 256 chunk adjoint sweeps run independently.
// Coefficient adjoints are reduced over the CTA and chained to the log-parameters by plan_chain (mtg_chain.cuh), as in K2.
// NumPy statement: oracle/kalman.py grad_chunked_adjoint (tests/test_oracle.py); parity vs complex-step of the celerite
// recursion: tests/test_scan_grad_gpu.py.
#pragma once
#include "mtg_scan.cuh"
#include "mtg_chain.cuh"

namespace k9 {

// backward element of a chunk
template <int J>
struct BElem {
    static constexpr int NS = J * (J + 1) / 2;
    double A[J * J], q[J], Om[NS];
    __device__ __forceinline__ void identity()
    {
#pragma unroll
        for (int i = 0; i < J * J; ++i) A[i] = 0.0;
#pragma unroll
        for (int i = 0; i < J; ++i) { A[i * J + i] = 1.0; q[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < NS; ++i) Om[i] = 0.0;
    }
};

template <int J>
__device__ __forceinline__ BElem<J> shfl_down(const BElem<J> &e, int delta)
{
    BElem<J> r;
#pragma unroll
    for (int i = 0; i < J * J; ++i) r.A[i] = __shfl_down_sync(0xffffffffu, e.A[i], delta, 32);
#pragma unroll
    for (int i = 0; i < J; ++i) r.q[i] = __shfl_down_sync(0xffffffffu, e.q[i], delta, 32);
#pragma unroll
    for (int i = 0; i < BElem<J>::NS; ++i) r.Om[i] = __shfl_down_sync(0xffffffffu, e.Om[i], delta, 32);
    return r;
}

// a (earlier chunk) then b (later chunk).  FULL = false: only q and Omega of the result.
template <int J>
__device__ __forceinline__ void compose(const BElem<J> &a, const BElem<J> &b, BElem<J> &out, bool FULL)
{
    double nq[J], X[J * J], nO[BElem<J>::NS];
#pragma unroll
    for (int r = 0; r < J; ++r) {
        double acc = a.q[r];
#pragma unroll
        for (int k = 0; k < J; ++k) acc = fma(a.A[k * J + r], b.q[k], acc);
        nq[r] = acc;
    }
    // X = Omega_b Phi_a ;  Omega = Phi_a^T X + Omega_a
#pragma unroll
    for (int r = 0; r < J; ++r)
#pragma unroll
        for (int c = 0; c < J; ++c) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(b.Om[sy<J>(r, k)], a.A[k * J + c], acc);
            X[r * J + c] = acc;
        }
#pragma unroll
    for (int r = 0; r < J; ++r)
#pragma unroll
        for (int c = r; c < J; ++c) {
            double acc = a.Om[sy<J>(r, c)];
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(a.A[k * J + r], X[k * J + c], acc);
            nO[sy<J>(r, c)] = acc;
        }
    if (FULL) {
        double nA[J * J];
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int c = 0; c < J; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < J; ++k) acc = fma(b.A[r * J + k], a.A[k * J + c], acc);
                nA[r * J + c] = acc;
            }
#pragma unroll
        for (int i = 0; i < J * J; ++i) out.A[i] = nA[i];
    }
#pragma unroll
    for (int i = 0; i < J; ++i) out.q[i] = nq[i];
#pragma unroll
    for (int i = 0; i < BElem<J>::NS; ++i) out.Om[i] = nO[i];
}

// number of per-thread accumulators reduced over the CTA: coefficient adjoints, R_b, four mean-parameter gradients
template <int JR, int JC, bool DUAL> struct GradCount { static constexpr int N = 2 * JR + (DUAL ? 6 : 4) * JC + 1 + 4; };

template <int JR, int JC, bool DUAL>
struct GScratch {
    static constexpr int J = JR + 2 * JC;
    double warp_tot[K9_WARPS][sizeof(BElem<J>) / sizeof(double)];
    double red[K9_WARPS][GradCount<JR, JC, DUAL>::N];
};

// tape: (m, P) before every step of the thread's chunk; component k of slot i at tape[(i * KT + k) * T + tid]
template <int J> struct TapeDims { static constexpr int KT = J + J * (J + 1) / 2; };

}  // namespace k9

// Where one K11 CTA keeps the series and the tape: both in shared memory when they fit (N <= ~2000 at J = 3), else the tape
// in a per-CTA global scratch, and for series beyond K9_STAGE_MAX_N points the data is read from global memory as well.
struct K11Layout { int stage; int tape_smem; size_t dyn_bytes; size_t scratch_doubles; };      // scratch per CTA, in doubles
#define K11_SMEM_LIMIT (200 * 1024)          // of the 227 KB an sm_100a CTA may use; the rest is static scratch
static inline K11Layout mtg_k11_layout(int J, int N)
{
    const size_t KT = (size_t)J + (size_t)J * (J + 1) / 2;
    const size_t CH = ((size_t)N + K9_THREADS - 1) / K9_THREADS;
    const size_t series = sizeof(double) * 3 * (size_t)N, tape = sizeof(double) * KT * CH * K9_THREADS;
    K11Layout L;
    L.stage = N <= K9_STAGE_MAX_N;
    const size_t base = L.stage ? series : 0;
    L.tape_smem = base + tape <= K11_SMEM_LIMIT;
    L.dyn_bytes = base + (L.tape_smem ? tape : 0);
    L.scratch_doubles = L.tape_smem ? 0 : KT * CH * K9_THREADS;
    return L;
}

namespace k9 {

// X <- X F  (J x J row-major times the block-diagonal transition)
template <int JR, int JC, bool DUAL>
__device__ __forceinline__ void right_mul(const typename Model<JR, JC, DUAL>::Trans &F, double *X)
{
    constexpr int J = JR + 2 * JC;
#pragma unroll
    for (int r = 0; r < J; ++r) {
#pragma unroll
        for (int j = 0; j < JR; ++j) X[r * J + j] *= F.phi[j];
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const int c = JR + 2 * j;
            const double x0 = X[r * J + c], x1 = X[r * J + c + 1];
            X[r * J + c] = fma(x0, F.f[j][0], x1 * F.f[j][2]);
            X[r * J + c + 1] = fma(x0, F.f[j][1], x1 * F.f[j][3]);
        }
    }
}
// X <- F^T X  (NC columns)
template <int JR, int JC, bool DUAL, int NC>
__device__ __forceinline__ void left_t(const typename Model<JR, JC, DUAL>::Trans &F, double *X)
{
#pragma unroll
    for (int j = 0; j < JR; ++j)
#pragma unroll
        for (int k = 0; k < NC; ++k) X[j * NC + k] *= F.phi[j];
#pragma unroll
    for (int j = 0; j < JC; ++j) {
        const int r = JR + 2 * j;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            const double x0 = X[r * NC + k], x1 = X[(r + 1) * NC + k];
            X[r * NC + k] = fma(F.f[j][0], x0, F.f[j][2] * x1);
            X[(r + 1) * NC + k] = fma(F.f[j][1], x0, F.f[j][3] * x1);
        }
    }
}

// log-likelihood and gradient of ONE series by the whole CTA (K9_THREADS threads; every thread calls it and gets the
// log-likelihood back; thread 0 writes grad[0 .. n_free), zeros when *st != 0).  t / yp / dp: the series, generic pointers --
// its shared-memory copy (S = 1) or global memory (yp, dp strided by S).  tape: TapeDims<J>::KT * CH * K9_THREADS doubles,
// CH = ceil(N / K9_THREADS), in shared memory when it fits and in a per-CTA global scratch otherwise (long series: written
// and read once per evaluation, coalesced over the threads; it stays in L2).
template <int JR, int JC, bool DUAL>
__device__ __forceinline__ double series_loglike_grad(const Plan &pl, const double *__restrict__ x, int N, const double *__restrict__ t,
                                                      const double *__restrict__ yp, const double *__restrict__ dp, size_t S,
                                                      bool ds_mean, double ds_mean_value, Scratch<JR + 2 * JC> &sc,
                                                      GScratch<JR, JC, DUAL> &gs, double *__restrict__ tape, int *st,
                                                      double *__restrict__ grad)
{
    constexpr int J = JR + 2 * JC;
    constexpr int NS = J * (J + 1) / 2;
    constexpr int KT = TapeDims<J>::KT;
    constexpr int NB = sizeof(BElem<J>) / sizeof(double);
    constexpr int NG = GradCount<JR, JC, DUAL>::N;
    constexpr int T = K9_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int CH = (N + T - 1) / T;
    const int n0 = min(tid * CH, N), n1 = min(n0 + CH, N);
    double *tp = tape + tid;
    auto slot = [&](int i, int k) -> double & { return tp[(size_t)(i * KT + k) * T]; };

    Model<JR, JC, DUAL> m;
    m.setup(pl, x);
    MeanEval mean = plan_mean(pl, x);
    if (ds_mean) { mean.kind = MTG_MEAN_CONSTANT; mean.p0 = ds_mean_value; }

    // ---- forward: chunk-start state from the scan, then the chunk itself ----------------------------------------
    Elem<J> e;
    e.identity();
    scan_prefix<JR, JC, DUAL, false>(m, mean, N, t, yp, dp, S, sc, e.b, e.C);
    double chi2 = 0.0, prod = 1.0; int esum = 0; bool bad = false;
    for (int n = n0; n < n1; ++n) {
        const int i = n - n0;
#pragma unroll
        for (int k = 0; k < J; ++k) slot(i, k) = e.b[k];
#pragma unroll
        for (int k = 0; k < NS; ++k) slot(i, J + k) = e.C[k];
        const double tn = t[n];
        const double dt = n > 0 ? tn - t[n - 1] : 0.0;
        double s, v;
        append_sv<JR, JC, DUAL>(m, e, dt, yp[(size_t)n * S] - mean(tn), dp[(size_t)n * S] + m.jit, &s, &v);
        bad = bad || !(s > 0.0);
        chi2 = fma(v * v, 1.0 / s, chi2);
        prod *= s;
        { int ex2; prod = frexp(prod, &ex2); esum += ex2; }
    }
    double part = chi2 + log(prod) + (double)esum * MTG_LN2;
    int badi = bad ? 1 : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        part += __shfl_xor_sync(0xffffffffu, part, d);
        badi |= __shfl_xor_sync(0xffffffffu, badi, d);
    }
    if (lane == 0) { sc.red_ll[wid] = part; sc.red_bad[wid] = badi; }
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0; int bb = 0;
#pragma unroll
        for (int w = 0; w < K9_WARPS; ++w) { tot += sc.red_ll[w]; bb |= sc.red_bad[w]; }
        double ll = -0.5 * (tot + (double)N * 1.8378770664093454836);
        int s2 = 0;
        if (bb) s2 = 1; else if (!isfinite(ll)) s2 = 2;
        if (s2) ll = -INFINITY;
        sc.result = ll; sc.status = s2;
    }
    __syncthreads();
    const double ll = sc.result;
    *st = sc.status;
    if (sc.status != 0) {                                // uniform over the CTA
        if (tid == 0) for (int k = 0; k < pl.n_free; ++k) grad[k] = 0.0;
        return ll;
    }

    // ---- suffix scan of the chunk elements (Phi, q, Omega) -------------------------------------------------------
    BElem<J> be;
#pragma unroll
    for (int i = 0; i < J * J; ++i) be.A[i] = e.A[i];
#pragma unroll
    for (int i = 0; i < J; ++i) be.q[i] = e.eta[i];
#pragma unroll
    for (int i = 0; i < NS; ++i) be.Om[i] = e.Jm[i];
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        const BElem<J> nx = shfl_down<J>(be, d);
        if (lane + d < 32) { BElem<J> r; compose<J>(be, nx, r, true); be = r; }
    }
    if (lane == 0) {
        const double *src = reinterpret_cast<const double *>(&be);
#pragma unroll
        for (int i = 0; i < NB; ++i) gs.warp_tot[wid][i] = src[i];
    }
    __syncthreads();
    if (wid == 0) {
        BElem<J> w;
        w.identity();
        if (lane < K9_WARPS) {
            double *dst = reinterpret_cast<double *>(&w);
#pragma unroll
            for (int i = 0; i < NB; ++i) dst[i] = gs.warp_tot[lane][i];
        }
#pragma unroll 1
        for (int d = 1; d < K9_WARPS; d <<= 1) {
            const BElem<J> nx = shfl_down<J>(w, d);
            if (lane + d < K9_WARPS) { BElem<J> r; compose<J>(w, nx, r, true); w = r; }
        }
        if (lane < K9_WARPS) {
            const double *src = reinterpret_cast<const double *>(&w);
#pragma unroll
            for (int i = 0; i < NB; ++i) gs.warp_tot[lane][i] = src[i];
        }
    }
    __syncthreads();
    // exclusive suffix of this thread: (lanes after it in the warp) then (warps after it)
    BElem<J> ex = shfl_down<J>(be, 1);
    if (lane == 31) ex.identity();
    if (wid < K9_WARPS - 1) {
        BElem<J> wp;
        double *dst = reinterpret_cast<double *>(&wp);
#pragma unroll
        for (int i = 0; i < NB; ++i) dst[i] = gs.warp_tot[wid + 1][i];
        BElem<J> r;
        compose<J>(ex, wp, r, false);
#pragma unroll
        for (int i = 0; i < J; ++i) ex.q[i] = r.q[i];
#pragma unroll
        for (int i = 0; i < NS; ++i) ex.Om[i] = r.Om[i];
    }

    // ---- reverse: adjoint sweep over this thread's chunk ------------------------------------------------------------
    double hv[J];
#pragma unroll
    for (int j = 0; j < JR; ++j) hv[j] = 1.0;
#pragma unroll
    for (int j = 0; j < JC; ++j) { hv[JR + 2 * j] = 1.0; hv[JR + 2 * j + 1] = m.h2[j]; }
    double lam[J], Lam[NS];
#pragma unroll
    for (int r = 0; r < J; ++r) lam[r] = ex.q[r];
#pragma unroll
    for (int r = 0; r < J; ++r)
#pragma unroll
        for (int c = r; c < J; ++c) Lam[sy<J>(r, c)] = 0.5 * fma(ex.q[r], ex.q[c], -ex.Om[sy<J>(r, c)]);

    double g_ar[JR > 0 ? JR : 1], g_cr[JR > 0 ? JR : 1];
    double g_ac[JC > 0 ? JC : 1], g_bc[JC > 0 ? JC : 1], g_cc[JC > 0 ? JC : 1], g_dc[JC > 0 ? JC : 1];
    double g_a2[JC > 0 ? JC : 1], g_c2[JC > 0 ? JC : 1];
    double g_R = 0.0, gm0 = 0.0, gm1 = 0.0, gm2 = 0.0, gm3 = 0.0;
#pragma unroll
    for (int j = 0; j < JR; ++j) { g_ar[j] = 0.0; g_cr[j] = 0.0; }
#pragma unroll
    for (int j = 0; j < JC; ++j) { g_ac[j] = g_bc[j] = g_cc[j] = g_dc[j] = g_a2[j] = g_c2[j] = 0.0; }

    // adjoint of P_inf from a symmetric J x J adjoint (block-diagonal entries only)
    auto add_pinf = [&](const double *Pb_, const double *Lp_, bool with_lp) {
#pragma unroll
        for (int j = 0; j < JR; ++j) g_ar[j] += Pb_[sy<J>(j, j)] - (with_lp ? Lp_[sy<J>(j, j)] : 0.0);
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const int r = JR + 2 * j;
            const double p00 = Pb_[sy<J>(r, r)] - (with_lp ? Lp_[sy<J>(r, r)] : 0.0);
            const double p01 = Pb_[sy<J>(r, r + 1)] - (with_lp ? Lp_[sy<J>(r, r + 1)] : 0.0);
            const double p11 = Pb_[sy<J>(r + 1, r + 1)] - (with_lp ? Lp_[sy<J>(r + 1, r + 1)] : 0.0);
            bool pair = true;
            if (DUAL) pair = m.c.pair[j];
            if (pair) {
                g_ac[j] += p00 + p11;
                g_bc[j] += 2.0 * p01;
                if (m.c.bc[j] != 0.0) {                  // P_inf(2,2) = a + 2 c b / d
                    const double cc = -m.c.ncc[j] * MTG_LN2, d = m.c.dc[j], b = m.c.bc[j];
                    g_bc[j] += p11 * 2.0 * cc / d; g_cc[j] += p11 * 2.0 * b / d; g_dc[j] -= p11 * 2.0 * cc * b / (d * d);
                }
            } else { g_ac[j] += p00; g_a2[j] += p11; }
        }
    };

#pragma unroll 1
    for (int n = n1 - 1; n >= n0; --n) {
        const int i = n - n0;
        double mp[J], Pp[NS];
#pragma unroll
        for (int k = 0; k < J; ++k) mp[k] = slot(i, k);
#pragma unroll
        for (int k = 0; k < NS; ++k) Pp[k] = slot(i, J + k);
        const double tn = t[n];
        const double dt = n > 0 ? tn - t[n - 1] : 0.0;
        typename Model<JR, JC, DUAL>::Trans F;
        m.transition(dt, F);
        // forward quantities of the step
        double mm[J], Pm[NS], g[J];
#pragma unroll
        for (int k = 0; k < J; ++k) mm[k] = mp[k];
        m.template left<1>(F, mm);
#pragma unroll
        for (int k = 0; k < NS; ++k) Pm[k] = Pp[k];
        m.predict_cov(F, Pm);
        m.Ph(Pm, g);
        const double s = m.hdot(g) + dp[(size_t)n * S] + m.jit, rs = 1.0 / s;
        const double v = yp[(size_t)n * S] - mean(tn) - m.hdot(mm);
        // adjoint of the update
        double lg = 0.0, gLg = 0.0, Lg[J];
#pragma unroll
        for (int r = 0; r < J; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(Lam[sy<J>(r, k)], g[k], acc);
            Lg[r] = acc; lg = fma(lam[r], g[r], lg);
        }
#pragma unroll
        for (int r = 0; r < J; ++r) gLg = fma(g[r], Lg[r], gLg);
        const double vb = (lg - v) * rs;
        const double sb = ((gLg - lg * v) * rs + 0.5 * v * v * rs - 0.5) * rs;
        double gb[J], mb[J], Pb[NS];
#pragma unroll
        for (int r = 0; r < J; ++r) {
            gb[r] = fma(lam[r], v * rs, fma(-2.0 * rs, Lg[r], sb * hv[r]));
            mb[r] = fma(-vb, hv[r], lam[r]);
        }
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int c = r; c < J; ++c) Pb[sy<J>(r, c)] = Lam[sy<J>(r, c)] + 0.5 * fma(gb[r], hv[c], hv[r] * gb[c]);
        g_R += sb;
        // mean parameters: residual = y - mean(t), its adjoint is vb
        if (!ds_mean) {
            if (mean.kind == MTG_MEAN_CONSTANT) gm0 -= vb;
            else if (mean.kind == MTG_MEAN_LINEAR) { gm0 = fma(-vb, tn, gm0); gm1 -= vb; }
            else {
                const double qq = tn - mean.p0, s2 = mean.p1 * mean.p1;
                const double ex_ = exp(-qq * qq / (2.0 * s2)), amp = mean.p2 / (2.0 * M_PI * mean.p1);
                gm0 -= vb * amp * ex_ * qq / s2;
                gm1 -= vb * amp * ex_ * (qq * qq / (s2 * mean.p1) - 1.0 / mean.p1);
                gm2 -= vb * ex_ / (2.0 * M_PI * mean.p1);
                gm3 -= vb;
            }
        }
        // T1 = Pb F ;  Lambda' = F^T T1 ;  lambda' = F^T mb
        double T1[J * J];
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int c = 0; c < J; ++c) T1[r * J + c] = Pb[sy<J>(r, c)];
        right_mul<JR, JC, DUAL>(F, T1);
        // adjoint of F on its block diagonal: F_b = 2 T1 X + mb m^T, X = P - P_inf
        double Xf[J * J];
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int c = 0; c < J; ++c) Xf[r * J + c] = Pp[sy<J>(r, c)];
#pragma unroll
        for (int j = 0; j < JR; ++j) Xf[j * J + j] -= m.c.ar[j];
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const int r = JR + 2 * j;
            Xf[r * J + r] -= m.c.ac[j]; Xf[r * J + r + 1] -= m.p12[j]; Xf[(r + 1) * J + r] -= m.p12[j]; Xf[(r + 1) * J + r + 1] -= m.p22[j];
        }
        auto fb = [&](int r, int c) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(T1[r * J + k], Xf[k * J + c], acc);
            return fma(2.0, acc, mb[r] * mp[c]);
        };
#pragma unroll
        for (int j = 0; j < JR; ++j) g_cr[j] = fma(-dt * F.phi[j], fb(j, j), g_cr[j]);
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const int r = JR + 2 * j;
            bool pair = true;
            if (DUAL) pair = m.c.pair[j];
            const double f00 = fb(r, r), f11 = fb(r + 1, r + 1);
            if (pair) {
                const double f01 = fb(r, r + 1), f10 = fb(r + 1, r);
                g_cc[j] = fma(-dt, fma(f00, F.f[j][0], fma(f01, F.f[j][1], fma(f10, F.f[j][2], f11 * F.f[j][3]))), g_cc[j]);
                // dF/dd = dt [[-F01, F00], [-F00, -F01]]   (F00 = e cos, F01 = e sin)
                g_dc[j] = fma(dt, fma(-F.f[j][1], f00 + f11, F.f[j][0] * (f01 - f10)), g_dc[j]);
            } else {
                g_cc[j] = fma(-dt * F.f[j][0], f00, g_cc[j]);
                g_c2[j] = fma(-dt * F.f[j][3], f11, g_c2[j]);
            }
        }
        left_t<JR, JC, DUAL, J>(F, T1);              // T1 = F^T Pb F (symmetric)
        double Lp[NS];
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int c = r; c < J; ++c) Lp[sy<J>(r, c)] = 0.5 * (T1[r * J + c] + T1[c * J + r]);
        add_pinf(Pb, Lp, true);
        left_t<JR, JC, DUAL, 1>(F, mb);
#pragma unroll
        for (int r = 0; r < J; ++r) lam[r] = mb[r];
#pragma unroll
        for (int k = 0; k < NS; ++k) Lam[k] = Lp[k];
    }
    if (tid == 0 && n1 > n0) add_pinf(Lam, Lam, false);      // the prior covariance of the first observation IS P_inf

    // ---- reduce the adjoints over the CTA, chain rule to the free parameters -------------------------------------
    double acc[NG];
    {
        int k = 0;
#pragma unroll
        for (int j = 0; j < JR; ++j) { acc[k++] = g_ar[j]; acc[k++] = g_cr[j]; }
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            acc[k++] = g_ac[j]; acc[k++] = g_bc[j]; acc[k++] = g_cc[j]; acc[k++] = g_dc[j];
            if (DUAL) { acc[k++] = g_a2[j]; acc[k++] = g_c2[j]; }
        }
        acc[k++] = g_R; acc[k++] = gm0; acc[k++] = gm1; acc[k++] = gm2; acc[k++] = gm3;
    }
#pragma unroll
    for (int k = 0; k < NG; ++k)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NG; ++k) gs.red[wid][k] = acc[k];
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < K9_WARPS; ++w) tot += gs.red[w][k];
            acc[k] = tot;
        }
        CoefAdj<JR, JC, DUAL> ga;
        ga.zero();
        int k = 0;
        const double Rb = acc[NG - 5];
        // plan_chain adds a_sum to every amplitude (celerite's diagonal carries jitter + sum a); here the amplitudes'
        // adjoints are complete on their own and only the jitter sees R_b
#pragma unroll
        for (int j = 0; j < JR; ++j) { ga.ar[j] = acc[k++] - Rb; ga.cr[j] = acc[k++]; }
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            ga.ac[j] = acc[k++] - Rb; ga.bc[j] = acc[k++]; ga.cc[j] = acc[k++]; ga.dc[j] = acc[k++];
            if (DUAL) { ga.a2[j] = acc[k++] - Rb; ga.c2[j] = acc[k++]; }
        }
        ga.a_sum = Rb;
        for (int q = 0; q < pl.n_free; ++q) grad[q] = 0.0;
        plan_chain<JR, JC, DUAL>(pl, x, m.c, ga, grad);
        if (!ds_mean) {
            auto addm = [&](int i, double v) { const int pos = pl.src_pos[pl.mean_idx[i]]; if (pos >= 0) grad[pos] += v; };
            addm(0, acc[NG - 4]);
            if (mean.kind != MTG_MEAN_CONSTANT) addm(1, acc[NG - 3]);
            if (mean.kind == MTG_MEAN_GAUSSIAN) { addm(2, acc[NG - 2]); addm(3, acc[NG - 1]); }
        }
    }
    __syncthreads();
    return ll;
}

}  // namespace k9
