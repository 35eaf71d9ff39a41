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
#include "mtg_common.cuh"
#include "mtg_math.cuh"
#include "mtg_philox.cuh"
#include "mtg_ensemble.cuh"

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

// e_i (earlier) then e_j (later).  FULL = false: only b and C of the result are formed (exclusive prefix for phase 3).
template <int J, bool FULL>
__device__ __forceinline__ void combine(const Elem<J> &ei, const Elem<J> &ej, Elem<J> &out)
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
        for (int k = 0; k < pl.n_jit; ++k) jit += exp(2.0 * plan_param(pl, x, pl.jit_idx[k]));
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
template <int JR, int JC, bool DUAL>
__device__ __forceinline__ void append(const Model<JR, JC, DUAL> &m, Elem<JR + 2 * JC> &e, double dt, double y, double R)
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
}

// CTA-collective scratch of one evaluation
template <int J>
struct Scratch {
    double warp_tot[K9_WARPS][sizeof(Elem<J>) / sizeof(double)];
    double red_ll[K9_WARPS];
    int red_bad[K9_WARPS];
    double result; int status;
};

// log-likelihood of ONE series by the whole CTA (every thread calls it; every thread gets the result).
// x: free parameter vector (any address space); ds: data-set column.  *st: 0 ok, 1 not positive definite, 2 non-finite.
// STAGED: t / yp / dp are shared-memory copies of the series (unit stride); otherwise global memory (yp, dp strided by S).
template <int JR, int JC, bool DUAL, bool STAGED>
__device__ __forceinline__ double series_loglike(const Plan &pl, const double *__restrict__ x, int N, const double *__restrict__ t,
                                                 const double *__restrict__ yp, const double *__restrict__ dp, size_t S,
                                                 bool ds_mean, double ds_mean_value, Scratch<JR + 2 * JC> &sc, int *st)
{
    auto T_ = [&](int n) { return STAGED ? t[n] : __ldg(t + n); };
    auto Y_ = [&](int n) { return STAGED ? yp[n] : __ldg(yp + (size_t)n * S); };
    auto D_ = [&](int n) { return STAGED ? dp[n] : __ldg(dp + (size_t)n * S); };
    constexpr int J = JR + 2 * JC;
    constexpr int NE = sizeof(Elem<J>) / sizeof(double);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int CH = (N + K9_THREADS - 1) / K9_THREADS;
    const int n0 = min(tid * CH, N), n1 = min(n0 + CH, N);

    Model<JR, JC, DUAL> m;
    m.setup(pl, x);
    MeanEval mean = plan_mean(pl, x);
    if (ds_mean) { mean.kind = MTG_MEAN_CONSTANT; mean.p0 = ds_mean_value; }

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
        if (lane >= d) { Elem<J> r; combine<J, true>(prev, e, r); e = r; }
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
            if (lane >= d && lane < K9_WARPS) { Elem<J> r; combine<J, true>(prev, w, r); w = r; }
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
        combine<J, false>(wp, ex, r);
#pragma unroll
        for (int i = 0; i < J; ++i) ex.b[i] = r.b[i];
#pragma unroll
        for (int i = 0; i < Elem<J>::NS; ++i) ex.C[i] = r.C[i];
    }

    // ---- phase 3: filter this thread's observations from the prefix -----------------------------------
    double mu[J], P[Elem<J>::NS];
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
    double ll;
    if (staged) {
        stage_series(k9_series_sm, a.N, a.t, a.y + ds, a.diag + ds, (size_t)a.S);
        ll = series_loglike<JR, JC, DUAL, true>(pl, x, a.N, k9_series_sm, k9_series_sm + a.N, k9_series_sm + 2 * a.N, 1, dsm,
                                                dsm ? a.ds_mean[ds] : 0.0, sc, &st);
    } else
        ll = series_loglike<JR, JC, DUAL, false>(pl, x, a.N, a.t, a.y + ds, a.diag + ds, (size_t)a.S, dsm, dsm ? a.ds_mean[ds] : 0.0, sc, &st);
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

    if (staged) stage_series(k10_series_sm, a.N, a.t, a.y + ds, a.diag + ds, (size_t)a.S);
    auto log_prob = [&](const double *x) -> double {   // CTA-collective; x in shared memory
        if (!plan_in_prior(pl, x)) return -INFINITY;   // uniform over the CTA
        int st;
        if (staged)
            return series_loglike<JR, JC, DUAL, true>(pl, x, a.N, k10_series_sm, k10_series_sm + a.N, k10_series_sm + 2 * a.N, 1, dsm, dsv, sc, &st);
        return series_loglike<JR, JC, DUAL, false>(pl, x, a.N, a.t, a.y + ds, a.diag + ds, (size_t)a.S, dsm, dsv, sc, &st);
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
