// mtg_grad.cu -- K2: log-likelihood AND its gradient in one kernel (forward + reverse sweep).
//
// The reference never forms a gradient: GPModelling.fit hands scipy's L-BFGS-B no `jac`
// (mind_the_gaps/gpmodelling.py:192-193) so every optimiser iteration costs ndim+1 likelihood
// evaluations by finite differences.  Here one thread owns one series and differentiates the celerite
// recursion in reverse mode WITHOUT an O(N) tape in HBM:
//   forward  : the K1 recursion, storing only a checkpoint of the state (S, f, W, D, z) every C steps
//              (time-major [N/C][K][Bc], coalesced; C ~ 6 for J = 3);
//   reverse  : segment by segment from the end -- restore the checkpoint, replay the C steps into
//              per-thread columns of SHARED memory ((C+1) x K doubles per thread, ~100 KB per CTA: this is
//              what the 227 KB of sm_100a shared memory buy), then run the adjoint sweep over the segment
//              out of shared memory, accumulating the adjoints of the celerite coefficients, which are chained
//              to the log-parameters on the device (plan_chain).
// The first version of this kernel taped every step to HBM (K = 14 doubles per step at J = 3, written once and
// read once): it ran at 0.64 of the HBM copy bandwidth and was bound there (profiles/r01_k2_*).  Replaying a
// segment costs one more forward sweep of FP64 work but removes 1 - 1/C of that traffic, which makes K2
// FP64-bound like K1 and removes the HBM-latency stall per step that dominated small (latency-bound) batches.
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
#include "mtg_chain.cuh"

#define K2_THREADS 128

// Record layout (smem slot or global checkpoint): S (NS), f (J), W (J), D, z  -> K doubles.
//   smem slot i of thread tid, component k:   slots[(i * K + k) * T + tid]      (conflict free: consecutive tids)
//   checkpoint of segment g, chunk-local series lb: ckpt[((size_t)g * K + k) * Bc + lb]   (coalesced)
// A checkpoint holds the recursion state BEFORE the first step of its segment with W unnormalised (exactly what
// K1State carries), so the recomputation replays the forward sweep bit for bit; smem slots hold W / D as the
// adjoint formulas want it.
template <int JR, int JC, bool DUAL>
__global__ void __launch_bounds__(K2_THREADS)
k2_loglike_grad(const __grid_constant__ Plan pl, const BatchArgs a, double *__restrict__ ckpt, int b0, int Bc, int C)
{
    constexpr int J = JR + 2 * JC;
    constexpr int NS = J * (J + 1) / 2;
    constexpr int K = NS + 2 * J + 2;
    extern __shared__ double2 k2_smem2[];
    double *k2_smem = reinterpret_cast<double *>(k2_smem2);
    double *tabs_sm = k2_smem;
    const mtgm::Tabs tb = mtgm::tabs_fill(tabs_sm);
    __syncthreads();
    const int T = blockDim.x;
    double *slots = k2_smem + MTG_TABS_DOUBLES + threadIdx.x;
    auto slot = [&](int i, int k) -> double & { return slots[(size_t)(i * K + k) * T]; };
    const int lb = blockIdx.x * T + threadIdx.x;            // index inside the chunk
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
    double *cp = ckpt + lb;
    const size_t Bs = (size_t)Bc;
    const int nseg = (N + C - 1) / C;

    // ---------------- forward sweep: checkpoint the state at every segment start -------------------
    K1State<JR, JC, DUAL> s;
    s.init(__ldg(a.t));
    {
        double t_nx = __ldg(a.t), y_nx = __ldg(yp), d_nx = __ldg(dp);
        int n = 0;
#pragma unroll 1
        for (int g = 0; g < nseg; ++g) {
            double *rec = cp + (size_t)g * K * Bs;
#pragma unroll
            for (int i = 0; i < NS; ++i) __stcs(rec + (size_t)i * Bs, s.S[i]);
#pragma unroll
            for (int j = 0; j < J; ++j) { __stcs(rec + (size_t)(NS + j) * Bs, s.f[j]); __stcs(rec + (size_t)(NS + J + j) * Bs, s.W[j]); }
            __stcs(rec + (size_t)(NS + 2 * J) * Bs, s.Dp);
            __stcs(rec + (size_t)(NS + 2 * J + 1) * Bs, s.zp);
            const int n1 = min(N, n + C);
#pragma unroll 1
            for (; n < n1; ++n) {
                const double tn = t_nx, yn = y_nx, dn = d_nx;
                if (n + 1 < N) { t_nx = __ldg(a.t + n + 1); y_nx = __ldg(yp + (size_t)(n + 1) * S); d_nx = __ldg(dp + (size_t)(n + 1) * S); }
                s.step(c, tb, tn, yn - mean(tn), dn, n);
            }
        }
    }
    int stt;
    double ll = s.loglike(N, &stt);
    if (!inside) { ll = -INFINITY; stt = 3; }
    a.out_ll[b] = ll;
    if (a.status) a.status[b] = stt;
    double *grad = a.out_grad + (size_t)b * pl.n_free;
    for (int k = 0; k < pl.n_free; ++k) grad[k] = 0.0;
    if (stt != 0) return;

    // ---------------- reverse sweep, one segment at a time ------------------------------------------
    double bS[NS > 0 ? NS : 1], bf[J > 0 ? J : 1], bW[J > 0 ? J : 1], bD = 0.0, bz = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) bS[i] = 0.0;
#pragma unroll
    for (int j = 0; j < J; ++j) { bf[j] = 0.0; bW[j] = 0.0; }
    CoefAdj<JR, JC, DUAL> g; g.zero();
    double gm0 = 0.0, gm1 = 0.0, gm2 = 0.0, gm3 = 0.0;       // mean-parameter gradients

#pragma unroll 1
    for (int sg = nseg - 1; sg >= 0; --sg) {
        const int n0 = sg * C, n1 = min(N, n0 + C);
        // ---- restore the checkpoint and replay the segment into the smem slots ----
        {
            const double *rec = cp + (size_t)sg * K * Bs;
#pragma unroll
            for (int i = 0; i < NS; ++i) s.S[i] = __ldcs(rec + (size_t)i * Bs);
#pragma unroll
            for (int j = 0; j < J; ++j) { s.f[j] = __ldcs(rec + (size_t)(NS + j) * Bs); s.W[j] = __ldcs(rec + (size_t)(NS + J + j) * Bs); }
            s.Dp = __ldcs(rec + (size_t)(NS + 2 * J) * Bs);
            s.zp = __ldcs(rec + (size_t)(NS + 2 * J + 1) * Bs);
            s.rDp = n0 > 0 ? mtgm::rcp_pos(s.Dp) : 0.0;
            s.tp = __ldg(a.t + (n0 > 0 ? n0 - 1 : 0));
            s.chi2 = 0.0; s.prod = 1.0; s.esum = 0; s.bad = false;
#pragma unroll
            for (int i = 0; i < NS; ++i) slot(0, i) = s.S[i];
#pragma unroll
            for (int j = 0; j < J; ++j) { slot(0, NS + j) = s.f[j]; slot(0, NS + J + j) = s.W[j] * s.rDp; }
            slot(0, NS + 2 * J) = s.Dp; slot(0, NS + 2 * J + 1) = s.zp;
            double t_nx = __ldg(a.t + n0), y_nx = __ldg(yp + (size_t)n0 * S), d_nx = __ldg(dp + (size_t)n0 * S);
#pragma unroll 1
            for (int n = n0; n < n1; ++n) {
                const double tn = t_nx, yn = y_nx, dn = d_nx;
                if (n + 1 < n1) { t_nx = __ldg(a.t + n + 1); y_nx = __ldg(yp + (size_t)(n + 1) * S); d_nx = __ldg(dp + (size_t)(n + 1) * S); }
                s.step(c, tb, tn, yn - mean(tn), dn, n);
                const int i1 = n - n0 + 1;
#pragma unroll
                for (int i = 0; i < NS; ++i) slot(i1, i) = s.S[i];
#pragma unroll
                for (int j = 0; j < J; ++j) { slot(i1, NS + j) = s.f[j]; slot(i1, NS + J + j) = s.W[j] * s.rDp; }
                slot(i1, NS + 2 * J) = s.Dp; slot(i1, NS + 2 * J + 1) = s.zp;
            }
        }
        // ---- adjoint sweep over the segment ----
        double tn = __ldg(a.t + n1 - 1);
#pragma unroll 1
        for (int n = n1 - 1; n >= n0; --n) {
            const int ic = n - n0 + 1, ip = n - n0;             // slots of step n and of step n-1
            double cS[NS > 0 ? NS : 1], cf[J > 0 ? J : 1], cW[J > 0 ? J : 1];
#pragma unroll
            for (int i = 0; i < NS; ++i) cS[i] = slot(ic, i);
#pragma unroll
            for (int j = 0; j < J; ++j) { cf[j] = slot(ic, NS + j); cW[j] = slot(ic, NS + J + j); }
            const double cD = slot(ic, NS + 2 * J), cz = slot(ic, NS + 2 * J + 1);
            const double tprev = n > 0 ? __ldg(a.t + n - 1) : tn;
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
            // propagate to step n-1:  bg = p o bf ; bA = (p p^T) o bS      (previous record: slot ip)
            double pW[J > 0 ? J : 1];
#pragma unroll
            for (int j = 0; j < J; ++j) pW[j] = slot(ip, NS + J + j);
            const double pD = slot(ip, NS + 2 * J), pz = slot(ip, NS + 2 * J + 1);
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
            for (int j = 0; j < J; ++j) { bf[j] = bg[j]; bW[j] = nbW[j]; }
            bD = nbD; bz = nbz; tn = tprev;
        }
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
// Launch shape: T threads per CTA and C steps per segment such that two CTAs fit the 227 KB of shared memory
// of an SM (the adjoint sweep needs ~190 registers, i.e. two 128-thread CTAs per SM anyway).
struct K2Shape { int T, C; size_t smem; };
static K2Shape k2_shape(const Plan &pl, int N)
{
    const int J = pl.jr + 2 * pl.jc;
    const int K = J * (J + 1) / 2 + 2 * J + 2;
    size_t kb = 110;
    if (const char *e = getenv("MTG_K2_SMEM_KB")) { int v = atoi(e); if (v >= 16 && v <= 224) kb = (size_t)v; }
    const size_t budget = kb * 1024 - MTG_TABS_DOUBLES * sizeof(double);
    K2Shape sh;
    sh.T = K2_THREADS;
    for (;;) {
        const size_t per_slot = (size_t)K * sh.T * sizeof(double);
        int slots = (int)(budget / per_slot);
        if (slots >= 3 || sh.T == 32) { sh.C = slots - 1 < 1 ? 1 : slots - 1; break; }
        sh.T /= 2;
    }
    if (const char *e = getenv("MTG_K2_SEGMENT")) { int v = atoi(e); if (v > 0 && v < sh.C) sh.C = v; }
    if (sh.C > N) sh.C = N;
    sh.smem = (MTG_TABS_DOUBLES + (size_t)(sh.C + 1) * K * sh.T) * sizeof(double);
    return sh;
}

template <int JR, int JC, bool DUAL>
static int launch_k2_one(const Plan &pl, const BatchArgs &a, double *ckpt, int b0, int Bc, cudaStream_t stream)
{
    const K2Shape sh = k2_shape(pl, a.N);
    auto kern = k2_loglike_grad<JR, JC, DUAL>;
    // (re-)stated on every launch: a few hundred ns, and safe with several host threads / devices
    MTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem));
    const dim3 grid((Bc + sh.T - 1) / sh.T), block(sh.T);
    kern<<<grid, block, sh.smem, stream>>>(pl, a, ckpt, b0, Bc, sh.C);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k2_loglike_grad launch");
}

template <int JR, int JC>
static int launch_k2(const Plan &pl, const BatchArgs &a, double *ckpt, int b0, int Bc, cudaStream_t stream)
{
    if (pl.has_sho) return launch_k2_one<JR, JC, true>(pl, a, ckpt, b0, Bc, stream);
    return launch_k2_one<JR, JC, false>(pl, a, ckpt, b0, Bc, stream);
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

// checkpoint bytes for one series (the K2 "tape": one state record per segment)
size_t mtg_k2_tape_bytes_per_series(const Plan &pl, int N)
{
    const int J = pl.jr + 2 * pl.jc;
    const K2Shape sh = k2_shape(pl, N);
    const int nseg = (N + sh.C - 1) / sh.C;
    return sizeof(double) * (size_t)nseg * (size_t)(J * (J + 1) / 2 + 2 * J + 2);
}

// series per chunk such that the checkpoints fit the scratch budget (MTG_TAPE_MB env, default 24 GiB)
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
    BatchArgs a{N, S, group, B, t, y, diag, ds_mean, params, out_ll, out_grad, status, flags, nullptr, nullptr, nullptr};
    // small batches: one CTA per series, forward scan + suffix scan of the adjoint (K11, mtg_scan_grad.cu): tens of
    // microseconds instead of K2's ~1 ms chain of three dependent sweeps
    const bool can_scan = mtg_k11_supports(pl, N);
    if ((flags & MTG_FLAG_SCAN) && !can_scan) { mtg_set_error("MTG_FLAG_SCAN: model / series length not covered by the scan gradient"); return MTG_EUNSUP; }
    if ((flags & MTG_FLAG_SCAN) || (can_scan && !(flags & MTG_FLAG_NO_SCAN) && N >= 64 && (long long)B <= mtg_scan_grad_max_batch()))
        return mtg_k11_dispatch(pl, a, stream);
    if ((rc = mtg_pool_setup())) return rc;
    const int chunk = mtg_k2_chunk(pl, N, B);
    double *tape = nullptr;
    MTG_CUDA(cudaMallocAsync((void **)&tape, mtg_k2_tape_bytes_per_series(pl, N) * (size_t)chunk, stream));
    rc = mtg_k2_run(pl, a, tape, chunk, stream);
    cudaError_t e = cudaFreeAsync(tape, stream);
    if (rc) return rc;
    return mtg_cuda_check(e, "cudaFreeAsync");
}
