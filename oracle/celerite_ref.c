/*
 * Oracle: plain-C restatement of celerite's CholeskySolver for the mind_the_gaps hot path.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never linked into the product library.
 *
 * What it restates (celerite 0.4.2 is un-vendored, pinned in /root/reference/pyproject.toml:16):
 *   solver.compute(jitter, a_real, c_real, a_comp, b_comp, c_comp, d_comp, t, yerr^2)   -> factor()
 *   solver.dot_solve(resid), solver.log_determinant()                                    -> solve_logdet()
 * as reached from /root/reference/mind_the_gaps/gpmodelling.py:54 (compute), :147-152
 * (_log_probability) and :168-169 (_neg_log_like).  Algorithm: Foreman-Mackey et al. 2017, sec. 5
 * (eqs. 43-46), restated in SURVEY.md 3.1.  Like celerite it is a TWO-pass scheme: the factor pass
 * stores D[N], W[N][J], phi[N][J], U[N][J]; the solve pass reads them back.  "parity unpinned" vs
 * the real package (not importable here); pinned against the dense-Cholesky oracle in tests/.
 *
 * Layouts: y, diag are series-major [S][N]; series b uses data set b / group.
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXJ 32

typedef struct {
    int N, J;
    double *D;   /* [N]    */
    double *W;   /* [N][J] */
    double *P;   /* [N][J]  phi_n (row 0 unused) */
    double *U;   /* [N][J] */
    double logdet;
} factor_t;

static int factor(factor_t *F, int N, const double *t, const double *diag, int Jr, int Jc,
                  const double *ar, const double *cr, const double *ac, const double *bc,
                  const double *cc, const double *dc, double jitter)
{
    const int J = Jr + 2 * Jc;
    double S[MAXJ][MAXJ], V[MAXJ], SU[MAXJ];
    double a_sum = jitter;
    for (int j = 0; j < Jr; ++j) a_sum += ar[j];
    for (int j = 0; j < Jc; ++j) a_sum += ac[j];
    memset(S, 0, sizeof S);
    F->N = N; F->J = J; F->logdet = 0.0;
    double Dp = 0.0;
    for (int n = 0; n < N; ++n) {
        double *U = F->U + (size_t)n * J, *W = F->W + (size_t)n * J, *P = F->P + (size_t)n * J;
        for (int j = 0; j < Jr; ++j) { U[j] = ar[j]; V[j] = 1.0; }
        for (int j = 0; j < Jc; ++j) {
            double cd = cos(dc[j] * t[n]), sd = sin(dc[j] * t[n]);
            U[Jr + 2 * j] = ac[j] * cd + bc[j] * sd;  U[Jr + 2 * j + 1] = ac[j] * sd - bc[j] * cd;
            V[Jr + 2 * j] = cd;                        V[Jr + 2 * j + 1] = sd;
        }
        if (n > 0) {
            const double dt = t[n] - t[n - 1];
            const double *Wp = F->W + (size_t)(n - 1) * J;
            for (int j = 0; j < Jr; ++j) P[j] = exp(-cr[j] * dt);
            for (int j = 0; j < Jc; ++j) { double e = exp(-cc[j] * dt); P[Jr + 2 * j] = e; P[Jr + 2 * j + 1] = e; }
            for (int j = 0; j < J; ++j)
                for (int k = 0; k <= j; ++k) {
                    double s = P[j] * P[k] * (S[j][k] + Dp * Wp[j] * Wp[k]);
                    S[j][k] = s; S[k][j] = s;
                }
        }
        double D = diag[n] + a_sum;
        for (int j = 0; j < J; ++j) {
            double acc = 0.0;
            for (int k = 0; k < J; ++k) acc += S[j][k] * U[k];
            SU[j] = acc;
            D -= U[j] * acc;
        }
        if (!(D > 0.0)) return 1;
        for (int j = 0; j < J; ++j) W[j] = (V[j] - SU[j]) / D;
        F->D[n] = D; F->logdet += log(D); Dp = D;
    }
    return 0;
}

/* z = L^-1 r by the forward sweep; returns r^T K^-1 r = sum z_n^2 / D_n */
static double dot_solve(const factor_t *F, const double *r, double mean)
{
    const int J = F->J, N = F->N;
    double f[MAXJ];
    double chi2 = 0.0, zp = 0.0;
    memset(f, 0, sizeof f);
    for (int n = 0; n < N; ++n) {
        const double *U = F->U + (size_t)n * J, *P = F->P + (size_t)n * J;
        double z = r[n] - mean;
        if (n > 0) {
            const double *Wp = F->W + (size_t)(n - 1) * J;
            for (int j = 0; j < J; ++j) { f[j] = P[j] * (f[j] + Wp[j] * zp); z -= U[j] * f[j]; }
        }
        chi2 += z * z / F->D[n];
        zp = z;
    }
    return chi2;
}

/* status: 0 ok, 1 not positive definite, 2 non-finite */
int oracle_loglike_batch(int N, const double *t, const double *y, const double *diag, int S, int group,
                         int B, int Jr, int Jc,
                         const double *ar, const double *cr, const double *ac, const double *bc,
                         const double *cc, const double *dc, const double *jitter, const double *mean,
                         double *out_ll, int *status, int nthreads)
{
    const int J = Jr + 2 * Jc;
    if (J > MAXJ || N < 1 || group < 1) return -1;
    (void)S;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    #pragma omp parallel
    {
        factor_t F;
        F.D = (double *)malloc(sizeof(double) * N);
        F.W = (double *)malloc(sizeof(double) * (size_t)N * (J > 0 ? J : 1));
        F.P = (double *)malloc(sizeof(double) * (size_t)N * (J > 0 ? J : 1));
        F.U = (double *)malloc(sizeof(double) * (size_t)N * (J > 0 ? J : 1));
        #pragma omp for schedule(dynamic, 4)
        for (int b = 0; b < B; ++b) {
            const size_t ds = (size_t)(b / group);
            int st = factor(&F, N, t, diag + ds * N, Jr, Jc, ar + (size_t)b * Jr, cr + (size_t)b * Jr,
                            ac + (size_t)b * Jc, bc + (size_t)b * Jc, cc + (size_t)b * Jc, dc + (size_t)b * Jc,
                            jitter ? jitter[b] : 0.0);
            double ll = -INFINITY;
            if (st == 0) {
                double chi2 = dot_solve(&F, y + ds * N, mean ? mean[b] : 0.0);
                ll = -0.5 * (chi2 + F.logdet + N * 1.8378770664093453);
                if (!isfinite(ll)) { ll = -INFINITY; st = 2; }
            }
            out_ll[b] = ll;
            if (status) status[b] = st;
        }
        free(F.D); free(F.W); free(F.P); free(F.U);
    }
    return 0;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
