// mtg_scan.cuh -- the parallel-in-time (associative scan) machinery shared by K9 / K10 (mtg_scan.cu), the scan gradient K11
// (mtg_scan_grad.cu) and the persistent L-BFGS kernel (mtg_lbfgs.cu): state-space model of a celerite kernel, the scan
// element and its operator (Sarkka & Garcia-Fernandez 2021), the CTA-collective log-likelihood of one series.
// See the header of mtg_scan.cu for the formulation.
#pragma once
#include "mtg_common.cuh"
#include "mtg_math.cuh"

#define K9_THREADS 256
#define K9_WARPS (K9_THREADS / 32)

namespace k9 {

template <int J> __device__ __forceinline__ constexpr int sy(int i, int j)      // packed upper triangle, any order
{ return i <= j ? i * J - (i * (i - 1)) / 2 + (j - i) : j * J - (j * (j - 1)) / 2 + (i - j); }

template <int J>
struct Elem {
    static constexpr int NS = J * (J + 1) / 2;
    double A[J * J], b[J], C[NS], eta[J], Jm[NS];
    __device__ __forceinline__ void identity()
    {
#pragma unroll
        for (int i = 0; i < J * J; ++i) A[i] = 0.0;
#pragma unroll
        for (int i = 0; i < J; ++i) { A[i * J + i] = 1.0; b[i] = 0.0; eta[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < NS; ++i) { C[i] = 0.0; Jm[i] = 0.0; }
    }
};

template <int J>
__device__ __forceinline__ Elem<J> shfl_up(const Elem<J> &e, int delta, int width)
{
    Elem<J> r;
#pragma unroll
    for (int i = 0; i < J * J; ++i) r.A[i] = __shfl_up_sync(0xffffffffu, e.A[i], delta, width);
#pragma unroll
    for (int i = 0; i < J; ++i) { r.b[i] = __shfl_up_sync(0xffffffffu, e.b[i], delta, width); r.eta[i] = __shfl_up_sync(0xffffffffu, e.eta[i], delta, width); }
#pragma unroll
    for (int i = 0; i < Elem<J>::NS; ++i) { r.C[i] = __shfl_up_sync(0xffffffffu, e.C[i], delta, width); r.Jm[i] = __shfl_up_sync(0xffffffffu, e.Jm[i], delta, width); }
    return r;
}

// inverse of a J x J matrix with det >= 1 (I + C J with C, J positive semi-definite): closed forms, J <= 4
template <int J> __device__ __forceinline__ void inverse(const double *T, double *M);
template <> __device__ __forceinline__ void inverse<1>(const double *T, double *M) { M[0] = 1.0 / T[0]; }
template <> __device__ __forceinline__ void inverse<2>(const double *T, double *M)
{
    const double r = 1.0 / (T[0] * T[3] - T[1] * T[2]);
    M[0] = T[3] * r; M[1] = -T[1] * r; M[2] = -T[2] * r; M[3] = T[0] * r;
}
template <> __device__ __forceinline__ void inverse<3>(const double *T, double *M)
{
    const double c00 = T[4] * T[8] - T[5] * T[7], c01 = T[5] * T[6] - T[3] * T[8], c02 = T[3] * T[7] - T[4] * T[6];
    const double r = 1.0 / (T[0] * c00 + T[1] * c01 + T[2] * c02);
    M[0] = c00 * r; M[3] = c01 * r; M[6] = c02 * r;
    M[1] = (T[2] * T[7] - T[1] * T[8]) * r; M[4] = (T[0] * T[8] - T[2] * T[6]) * r; M[7] = (T[1] * T[6] - T[0] * T[7]) * r;
    M[2] = (T[1] * T[5] - T[2] * T[4]) * r; M[5] = (T[2] * T[3] - T[0] * T[5]) * r; M[8] = (T[0] * T[4] - T[1] * T[3]) * r;
}

template <> __device__ __forceinline__ void inverse<4>(const double *a, double *b)      // Laplace expansion by 2 x 2 minors
{
    const double s0 = a[0] * a[5] - a[4] * a[1], s1 = a[0] * a[6] - a[4] * a[2], s2 = a[0] * a[7] - a[4] * a[3];
    const double s3 = a[1] * a[6] - a[5] * a[2], s4 = a[1] * a[7] - a[5] * a[3], s5 = a[2] * a[7] - a[6] * a[3];
    const double c5 = a[10] * a[15] - a[14] * a[11], c4 = a[9] * a[15] - a[13] * a[11], c3 = a[9] * a[14] - a[13] * a[10];
    const double c2 = a[8] * a[15] - a[12] * a[11], c1 = a[8] * a[14] - a[12] * a[10], c0 = a[8] * a[13] - a[12] * a[9];
    const double r = 1.0 / (s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0);
    b[0] = (a[5] * c5 - a[6] * c4 + a[7] * c3) * r;    b[1] = (-a[1] * c5 + a[2] * c4 - a[3] * c3) * r;
    b[2] = (a[13] * s5 - a[14] * s4 + a[15] * s3) * r; b[3] = (-a[9] * s5 + a[10] * s4 - a[11] * s3) * r;
    b[4] = (-a[4] * c5 + a[6] * c2 - a[7] * c1) * r;   b[5] = (a[0] * c5 - a[2] * c2 + a[3] * c1) * r;
    b[6] = (-a[12] * s5 + a[14] * s2 - a[15] * s1) * r; b[7] = (a[8] * s5 - a[10] * s2 + a[11] * s1) * r;
    b[8] = (a[4] * c4 - a[5] * c2 + a[7] * c0) * r;    b[9] = (-a[0] * c4 + a[1] * c2 - a[3] * c0) * r;
    b[10] = (a[12] * s4 - a[13] * s2 + a[15] * s0) * r; b[11] = (-a[8] * s4 + a[9] * s2 - a[11] * s0) * r;
    b[12] = (-a[4] * c3 + a[5] * c1 - a[6] * c0) * r;  b[13] = (a[0] * c3 - a[1] * c1 + a[2] * c0) * r;
    b[14] = (-a[12] * s3 + a[13] * s1 - a[14] * s0) * r; b[15] = (a[8] * s3 - a[9] * s1 + a[10] * s0) * r;
}

// e_i (earlier) then e_j (later).  FULL = false (run time: the operator is inlined once): only b and C of the result are formed.
template <int J>
__device__ __forceinline__ void combine(const Elem<J> &ei, const Elem<J> &ej, Elem<J> &out, bool FULL)
{
    double T[J * J], M[J * J], AjM[J * J];
#pragma unroll
    for (int r = 0; r < J; ++r)
#pragma unroll
        for (int c = 0; c < J; ++c) {
            double acc = (r == c) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(ei.C[sy<J>(r, k)], ej.Jm[sy<J>(k, c)], acc);
            T[r * J + c] = acc;
        }
    inverse<J>(T, M);
#pragma unroll
    for (int r = 0; r < J; ++r)
#pragma unroll
        for (int c = 0; c < J; ++c) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(ej.A[r * J + k], M[k * J + c], acc);
            AjM[r * J + c] = acc;
        }
    // b = AjM (b_i + C_i eta_j) + b_j
    double u[J];
#pragma unroll
    for (int r = 0; r < J; ++r) {
        double acc = ei.b[r];
#pragma unroll
        for (int k = 0; k < J; ++k) acc = fma(ei.C[sy<J>(r, k)], ej.eta[k], acc);
        u[r] = acc;
    }
    double nb[J];
#pragma unroll
    for (int r = 0; r < J; ++r) {
        double acc = ej.b[r];
#pragma unroll
        for (int k = 0; k < J; ++k) acc = fma(AjM[r * J + k], u[k], acc);
        nb[r] = acc;
    }
    // C = AjM C_i A_j^T + C_j
    double X[J * J];
#pragma unroll
    for (int r = 0; r < J; ++r)
#pragma unroll
        for (int c = 0; c < J; ++c) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(AjM[r * J + k], ei.C[sy<J>(k, c)], acc);
            X[r * J + c] = acc;
        }
    double nC[Elem<J>::NS];
#pragma unroll
    for (int r = 0; r < J; ++r)
#pragma unroll
        for (int c = r; c < J; ++c) {
            double acc = ej.C[sy<J>(r, c)];
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(X[r * J + k], ej.A[c * J + k], acc);
            nC[sy<J>(r, c)] = acc;
        }
    if (FULL) {
        // A = AjM A_i
        double nA[J * J];
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int c = 0; c < J; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < J; ++k) acc = fma(AjM[r * J + k], ei.A[k * J + c], acc);
                nA[r * J + c] = acc;
            }
        // eta = A_i^T M^T (eta_j - J_j b_i) + eta_i
        double v[J], v2[J], ne[J];
#pragma unroll
        for (int r = 0; r < J; ++r) {
            double acc = ej.eta[r];
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(-ej.Jm[sy<J>(r, k)], ei.b[k], acc);
            v[r] = acc;
        }
#pragma unroll
        for (int r = 0; r < J; ++r) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(M[k * J + r], v[k], acc);
            v2[r] = acc;
        }
#pragma unroll
        for (int r = 0; r < J; ++r) {
            double acc = ei.eta[r];
#pragma unroll
            for (int k = 0; k < J; ++k) acc = fma(ei.A[k * J + r], v2[k], acc);
            ne[r] = acc;
        }
        // J = A_i^T (M^T J_j) A_i + J_i
        double Y[J * J], Z[J * J];
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int c = 0; c < J; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < J; ++k) acc = fma(M[k * J + r], ej.Jm[sy<J>(k, c)], acc);
                Y[r * J + c] = acc;
            }
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int c = 0; c < J; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < J; ++k) acc = fma(Y[r * J + k], ei.A[k * J + c], acc);
                Z[r * J + c] = acc;
            }
        double nJ[Elem<J>::NS];
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int c = r; c < J; ++c) {
                double acc = ei.Jm[sy<J>(r, c)];
#pragma unroll
                for (int k = 0; k < J; ++k) acc = fma(ei.A[k * J + r], Z[k * J + c], acc);
                nJ[sy<J>(r, c)] = acc;
            }
#pragma unroll
        for (int i = 0; i < J * J; ++i) out.A[i] = nA[i];
#pragma unroll
        for (int i = 0; i < J; ++i) out.eta[i] = ne[i];
#pragma unroll
        for (int i = 0; i < Elem<J>::NS; ++i) out.Jm[i] = nJ[i];
    }
#pragma unroll
    for (int i = 0; i < J; ++i) out.b[i] = nb[i];
#pragma unroll
    for (int i = 0; i < Elem<J>::NS; ++i) out.C[i] = nC[i];
}

// ---- the state-space model of one series ------------------------------------------------------------
template <int JR, int JC, bool DUAL>
struct Model {
    static constexpr int J = JR + 2 * JC;
    static constexpr int NS = J * (J + 1) / 2;
    Coef<JR, JC, DUAL> c;
    double p12[JC > 0 ? JC : 1], p22[JC > 0 ? JC : 1], h2[JC > 0 ? JC : 1];     // complex slot: P_inf off-diagonal, (2,2); h = (1, h2)
    double jit;

    __device__ __forceinline__ void setup(const Plan &pl, const double *__restrict__ x)
    {
        plan_coefficients<JR, JC, DUAL>(pl, x, c);
        jit = 0.0;
        for (int k = 0; k < pl.n_jit; ++k) jit += plan_jitter(pl, x, k);
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            bool pair = true;
            if (DUAL) pair = c.pair[j];
            if (pair) {
                p12[j] = c.bc[j];
                // a + 2 c b / d; b == 0 (ComplexTerm(a, c, d), Lorentzian, Cosinus) keeps it isotropic also when d == 0
                p22[j] = (c.bc[j] == 0.0) ? c.ac[j] : c.ac[j] + 2.0 * (-c.ncc[j] * MTG_LN2) * c.bc[j] / c.dc[j];
                h2[j] = 0.0;
            } else {            // overdamped SHO: two independent real terms share the slot
                p12[j] = 0.0; p22[j] = DUAL ? c.a2[j] : 0.0; h2[j] = 1.0;
            }
        }
    }

    // transition blocks at lag dt: real phi[j]; complex f11 f12 f21 f22 per slot
    struct Trans { double phi[JR > 0 ? JR : 1]; double f[JC > 0 ? JC : 1][4]; };

    __device__ __forceinline__ void transition(double dt, Trans &F) const
    {
#pragma unroll
        for (int j = 0; j < JR; ++j) F.phi[j] = mtgm::exp2_nonpos(c.ncr[j] * dt);
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const double e = mtgm::exp2_nonpos(c.ncc[j] * dt);
            bool pair = true;
            if (DUAL) pair = c.pair[j];
            if (pair) {
                double sn, cs;
                mtgm::sincos_fast(c.dc[j] * dt, &sn, &cs);
                F.f[j][0] = e * cs; F.f[j][1] = e * sn; F.f[j][2] = -e * sn; F.f[j][3] = e * cs;
            } else {
                F.f[j][0] = e; F.f[j][1] = 0.0; F.f[j][2] = 0.0; F.f[j][3] = DUAL ? mtgm::exp2_nonpos(c.nc2[j] * dt) : 0.0;
            }
        }
    }

    // X <- F X for a J x NC row-major matrix (rows are state components)
    template <int NC>
    __device__ __forceinline__ void left(const Trans &F, double *X) const
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
                X[r * NC + k] = fma(F.f[j][0], x0, F.f[j][1] * x1);
                X[(r + 1) * NC + k] = fma(F.f[j][2], x0, F.f[j][3] * x1);
            }
        }
    }

    // P <- F P F^T + Q,  Q = P_inf - F P_inf F^T  (packed symmetric)
    __device__ __forceinline__ void predict_cov(const Trans &F, double *P) const
    {
        double X[J * J];
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int k = 0; k < J; ++k) X[r * J + k] = P[sy<J>(r, k)];
        // D = P - P_inf on the diagonal blocks, then F D F^T + P_inf
#pragma unroll
        for (int j = 0; j < JR; ++j) X[j * J + j] -= c.ar[j];
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const int r = JR + 2 * j;
            X[r * J + r] -= c.ac[j]; X[r * J + r + 1] -= p12[j]; X[(r + 1) * J + r] -= p12[j]; X[(r + 1) * J + r + 1] -= p22[j];
        }
        left<J>(F, X);                         // F D
        // (F D) F^T = (F (F D)^T)^T: transpose, apply, read back symmetric
        double Y[J * J];
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int k = 0; k < J; ++k) Y[r * J + k] = X[k * J + r];
        left<J>(F, Y);
#pragma unroll
        for (int r = 0; r < J; ++r)
#pragma unroll
            for (int k = r; k < J; ++k) P[sy<J>(r, k)] = Y[k * J + r];
#pragma unroll
        for (int j = 0; j < JR; ++j) P[sy<J>(j, j)] += c.ar[j];
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const int r = JR + 2 * j;
            P[sy<J>(r, r)] += c.ac[j]; P[sy<J>(r, r + 1)] += p12[j]; P[sy<J>(r + 1, r + 1)] += p22[j];
        }
    }

    __device__ __forceinline__ void stationary(double *P) const
    {
#pragma unroll
        for (int i = 0; i < NS; ++i) P[i] = 0.0;
#pragma unroll
        for (int j = 0; j < JR; ++j) P[sy<J>(j, j)] = c.ar[j];
#pragma unroll
        for (int j = 0; j < JC; ++j) {
            const int r = JR + 2 * j;
            P[sy<J>(r, r)] = c.ac[j]; P[sy<J>(r, r + 1)] = p12[j]; P[sy<J>(r + 1, r + 1)] = p22[j];
        }
    }

    // h^T x  and  P h  (h has ones at the first component of every slot, h2 at the second of a complex slot)
    __device__ __forceinline__ double hdot(const double *x) const
    {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < JR; ++j) s += x[j];
#pragma unroll
        for (int j = 0; j < JC; ++j) s += fma(h2[j], x[JR + 2 * j + 1], x[JR + 2 * j]);
        return s;
    }
    __device__ __forceinline__ void Ph(const double *P, double *out) const
    {
#pragma unroll
        for (int r = 0; r < J; ++r) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < JR; ++j) s += P[sy<J>(r, j)];
#pragma unroll
            for (int j = 0; j < JC; ++j) s += fma(h2[j], P[sy<J>(r, JR + 2 * j + 1)], P[sy<J>(r, JR + 2 * j)]);
            out[r] = s;
        }
    }
};

// append one observation (lag dt, residual y, variance R) to the composite element: rank-one form of the operator
// *s_out = innovation variance h^T P h + R, *w_out = innovation y - h^T b' (what the filter's likelihood term needs)
template <int JR, int JC, bool DUAL>
__device__ __forceinline__ void append_sv(const Model<JR, JC, DUAL> &m, Elem<JR + 2 * JC> &e, double dt, double y, double R,
                                          double *s_out, double *w_out)
{
    constexpr int J = JR + 2 * JC;
    typename Model<JR, JC, DUAL>::Trans F;
    m.transition(dt, F);
    m.template left<J>(F, e.A);                 // A' = F A
    m.template left<1>(F, e.b);                 // b' = F b
    m.predict_cov(F, e.C);                      // P  = F C F^T + Q
    double Ph[J];
    m.Ph(e.C, Ph);
    const double s = m.hdot(Ph) + R;            // h^T P h + R
    const double rs = 1.0 / s;
    const double w = y - m.hdot(e.b);
    // ga = A'^T h
    double ga[J];
#pragma unroll
    for (int k = 0; k < J; ++k) {
        double col[J];
#pragma unroll
        for (int r = 0; r < J; ++r) col[r] = e.A[r * J + k];
        ga[k] = m.hdot(col);
    }
#pragma unroll
    for (int k = 0; k < J; ++k) e.eta[k] = fma(ga[k], w * rs, e.eta[k]);
#pragma unroll
    for (int r = 0; r < J; ++r)
#pragma unroll
        for (int k = r; k < J; ++k) e.Jm[sy<J>(r, k)] = fma(ga[r] * rs, ga[k], e.Jm[sy<J>(r, k)]);
    // Kalman update of the affine map: K = P h / s
#pragma unroll
    for (int r = 0; r < J; ++r) {
        const double K = Ph[r] * rs;
#pragma unroll
        for (int k = 0; k < J; ++k) e.A[r * J + k] = fma(-K, ga[k], e.A[r * J + k]);
        e.b[r] = fma(K, w, e.b[r]);
#pragma unroll
        for (int k = r; k < J; ++k) e.C[sy<J>(r, k)] = fma(-K, Ph[k], e.C[sy<J>(r, k)]);
    }
    *s_out = s; *w_out = w;
}
template <int JR, int JC, bool DUAL>
__device__ __forceinline__ void append(const Model<JR, JC, DUAL> &m, Elem<JR + 2 * JC> &e, double dt, double y, double R)
{
    double s, w;
    append_sv<JR, JC, DUAL>(m, e, dt, y, R, &s, &w);
}

// CTA-collective scratch of one evaluation
template <int J>
struct Scratch {
    double warp_tot[K9_WARPS][sizeof(Elem<J>) / sizeof(double)];
    double red_ll[K9_WARPS];
    int red_bad[K9_WARPS];
    double result; int status;
};

// Phases 1 and 2 of the scan: composite element of this thread's observations, inclusive scan over the CTA, and from it the
// FILTERED state (mean mu, covariance P) just before the thread's first observation (thread 0: the prior 0, P_inf).
// CTA-collective: every thread calls it.
template <int JR, int JC, bool DUAL, bool STAGED>
__device__ __forceinline__ void scan_prefix(const Model<JR, JC, DUAL> &m, const MeanEval &mean, int N, const double *__restrict__ t,
                                            const double *__restrict__ yp, const double *__restrict__ dp, size_t S,
                                            Scratch<JR + 2 * JC> &sc, double *mu, double *P)
{
    auto T_ = [&](int n) { return t[n]; };
    auto Y_ = [&](int n) { return STAGED ? yp[n] : yp[(size_t)n * S]; };
    auto D_ = [&](int n) { return STAGED ? dp[n] : dp[(size_t)n * S]; };
    constexpr int J = JR + 2 * JC;
    constexpr int NE = sizeof(Elem<J>) / sizeof(double);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int CH = (N + K9_THREADS - 1) / K9_THREADS;
    const int n0 = min(tid * CH, N), n1 = min(n0 + CH, N);

    // ---- phase 1: composite of this thread's observations -------------------------------------------
    Elem<J> e;
    e.identity();
    if (tid == 0) {                               // the prior: mean 0, covariance P_inf
#pragma unroll
        for (int i = 0; i < J * J; ++i) e.A[i] = 0.0;
        m.stationary(e.C);
    }
    for (int n = n0; n < n1; ++n) {
        const double tn = T_(n);
        const double dt = n > 0 ? tn - T_(n - 1) : 0.0;
        append<JR, JC, DUAL>(m, e, dt, Y_(n) - mean(tn), D_(n) + m.jit);
    }

    // ---- phase 2: inclusive scan over the CTA ---------------------------------------------------------
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        const Elem<J> prev = shfl_up<J>(e, d, 32);
        if (lane >= d) { Elem<J> r; combine<J>(prev, e, r, true); e = r; }
    }
    __syncthreads();                              // scratch free (a previous evaluation may still be reading it)
    if (lane == 31) {
        const double *src = reinterpret_cast<const double *>(&e);
#pragma unroll
        for (int i = 0; i < NE; ++i) sc.warp_tot[wid][i] = src[i];
    }
    __syncthreads();
    if (wid == 0) {
        Elem<J> w;
        w.identity();
        if (lane < K9_WARPS) {
            double *dst = reinterpret_cast<double *>(&w);
#pragma unroll
            for (int i = 0; i < NE; ++i) dst[i] = sc.warp_tot[lane][i];
        }
#pragma unroll 1
        for (int d = 1; d < K9_WARPS; d <<= 1) {
            const Elem<J> prev = shfl_up<J>(w, d, 32);
            if (lane >= d && lane < K9_WARPS) { Elem<J> r; combine<J>(prev, w, r, true); w = r; }
        }
        if (lane < K9_WARPS) {
            const double *src = reinterpret_cast<const double *>(&w);
#pragma unroll
            for (int i = 0; i < NE; ++i) sc.warp_tot[lane][i] = src[i];
        }
    }
    __syncthreads();
    // exclusive prefix of this thread: (warps before) then (lanes before)
    Elem<J> ex = shfl_up<J>(e, 1, 32);
    if (lane == 0) ex.identity();
    if (wid > 0) {
        Elem<J> wp;
        double *dst = reinterpret_cast<double *>(&wp);
#pragma unroll
        for (int i = 0; i < NE; ++i) dst[i] = sc.warp_tot[wid - 1][i];
        Elem<J> r;
        combine<J>(wp, ex, r, false);
#pragma unroll
        for (int i = 0; i < J; ++i) ex.b[i] = r.b[i];
#pragma unroll
        for (int i = 0; i < Elem<J>::NS; ++i) ex.C[i] = r.C[i];
    }

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < J; ++i) mu[i] = 0.0;
        m.stationary(P);
    } else {
#pragma unroll
        for (int i = 0; i < J; ++i) mu[i] = ex.b[i];
#pragma unroll
        for (int i = 0; i < Elem<J>::NS; ++i) P[i] = ex.C[i];
    }
}

// log-likelihood of ONE series by the whole CTA (every thread calls it; every thread gets the result).
// x: free parameter vector (any address space); ds: data-set column.  *st: 0 ok, 1 not positive definite, 2 non-finite.
// STAGED: yp / dp have unit stride (compile time); otherwise stride S (run time).  The pointers are generic: the series in
// global memory or its shared-memory copy (K9 / K10 use the run-time-stride form for both, so the function exists once).
template <int JR, int JC, bool DUAL, bool STAGED>
__device__ __forceinline__ double series_loglike(const Plan &pl, const double *__restrict__ x, int N, const double *__restrict__ t,
                                                 const double *__restrict__ yp, const double *__restrict__ dp, size_t S,
                                                 bool ds_mean, double ds_mean_value, Scratch<JR + 2 * JC> &sc, int *st)
{
    auto T_ = [&](int n) { return t[n]; };
    auto Y_ = [&](int n) { return STAGED ? yp[n] : yp[(size_t)n * S]; };
    auto D_ = [&](int n) { return STAGED ? dp[n] : dp[(size_t)n * S]; };
    constexpr int J = JR + 2 * JC;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int CH = (N + K9_THREADS - 1) / K9_THREADS;
    const int n0 = min(tid * CH, N), n1 = min(n0 + CH, N);

    Model<JR, JC, DUAL> m;
    m.setup(pl, x);
    MeanEval mean = plan_mean(pl, x);
    if (ds_mean) { mean.kind = MTG_MEAN_CONSTANT; mean.p0 = ds_mean_value; }

    double mu[J], P[Elem<J>::NS];
    scan_prefix<JR, JC, DUAL, STAGED>(m, mean, N, t, yp, dp, S, sc, mu, P);

    // ---- phase 3: filter this thread's observations from the prefix -----------------------------------
    double chi2 = 0.0, prod = 1.0; int esum = 0; bool bad = false;
    for (int n = n0; n < n1; ++n) {
        const double tn = T_(n);
        const double dt = n > 0 ? tn - T_(n - 1) : 0.0;
        typename Model<JR, JC, DUAL>::Trans F;
        m.transition(dt, F);
        m.template left<1>(F, mu);
        m.predict_cov(F, P);
        double Ph[J];
        m.Ph(P, Ph);
        const double s = m.hdot(Ph) + D_(n) + m.jit;
        const double v = Y_(n) - mean(tn) - m.hdot(mu);
        bad = bad || !(s > 0.0);
        const double rs = 1.0 / s;
        chi2 = fma(v * v, rs, chi2);
        prod *= s;
        {                                           // log s summed as the log of a renormalised running product
            int ex2; prod = frexp(prod, &ex2); esum += ex2;
        }
#pragma unroll
        for (int r = 0; r < J; ++r) {
            const double K = Ph[r] * rs;
            mu[r] = fma(K, v, mu[r]);
#pragma unroll
            for (int k = r; k < J; ++k) P[sy<J>(r, k)] = fma(-K, Ph[k], P[sy<J>(r, k)]);
        }
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
    *st = sc.status;
    return sc.result;
}

// Out-of-line form for kernels that evaluate at several call sites (K10): one copy of the ~3 k instructions.
template <int JR, int JC, bool DUAL>
__device__ __noinline__ double series_loglike_call(const Plan &pl, const double *x, int N, const double *t, const double *yp, const double *dp,
                                                   size_t S, bool ds_mean, double ds_mean_value, Scratch<JR + 2 * JC> &sc, int *st)
{
    return series_loglike<JR, JC, DUAL, false>(pl, x, N, t, yp, dp, S, ds_mean, ds_mean_value, sc, st);
}

// Series of up to K9_STAGE_MAX_N observations are copied to shared memory first (coalesced, one round of latency): every
// thread then reads its observations twice (phases 1 and 3) -- and K10 once per proposal -- without touching L2 again.
#define K9_STAGE_MAX_N 4096
__device__ __forceinline__ void stage_series(double *sm, int N, const double *__restrict__ t, const double *__restrict__ yp,
                                             const double *__restrict__ dp, size_t S)
{
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        sm[n] = __ldg(t + n); sm[N + n] = __ldg(yp + (size_t)n * S); sm[2 * N + n] = __ldg(dp + (size_t)n * S);
    }
    __syncthreads();
}

}  // namespace k9
