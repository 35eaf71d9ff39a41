// mtg_common.cuh -- shared device/host helpers for libmtg_b200 (sm_100a only).
//
// * Plan: the host-compiled form of mtg_model_t.  celerite evaluates
//   Term.get_real_coefficients / get_complex_coefficients in Python per parameter set
//   (reference call sites mind_the_gaps/models/celerite_models.py:9-31, 39-52, 58-66, 77-83);
//   here the term list is compiled once into per-slot "programs" so that every recursion slot has
//   a compile-time register position and the parameter -> coefficient map runs on the device.
// * mbarrier / cp.async.bulk (TMA bulk copy, SASS UBLKCP) wrappers used to stage the shared
//   time / rate / variance arrays through shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/mtg_b200.h"
#define MTG_LOG2E 1.4426950408889634
#define MTG_LN2 0.6931471805599453

#define MTG_MAXR 4          // real recursion slots
#define MTG_MAXC 8          // complex (width-2) recursion slots

// complex-slot program kinds (device side)
enum { CS_COMPLEX3 = 0, CS_COMPLEX4, CS_SHO, CS_MATERN32, CS_LORENTZIAN, CS_COSINUS, CS_BENDINGPL, CS_RAW };

struct Plan {
    int32_t jr, jc;                 // recursion slots actually computed
    int32_t j_celerite;             // width as celerite counts it (incl. degenerate real terms)
    int32_t has_sho;                // any SHO slot (can flip to two real terms when Q < 1/2)
    int32_t b_zero;                 // every complex slot has b == 0 structurally (ComplexTerm(a,c,d), Lorentzian, Cosinus)
    int32_t r_ia[MTG_MAXR], r_ic[MTG_MAXR];     // real slot: a = exp(p[ia]), c = exp(p[ic])
    int32_t r_raw[MTG_MAXR];                    // ... or a = p[ia], c = p[ic] (MTG_TERM_RAW_REAL: coefficients supplied by the caller)
    int32_t c_kind[MTG_MAXC];
    int32_t c_idx[MTG_MAXC][4];
    double  c_extra[MTG_MAXC];
    int32_t n_jit; int32_t jit_idx[MTG_MAX_TERMS]; int32_t jit_raw[MTG_MAX_TERMS];     // jitter = exp(2 p) or p (raw)
    int32_t n_bpl; int32_t bpl_s0[MTG_MAX_TERMS], bpl_q[MTG_MAX_TERMS];   // BendingPowerlaw prior pairs
    int32_t mean_kind; int32_t mean_idx[4];
    int32_t n_full, n_free;
    int32_t src_pos[MTG_MAX_PARAMS];      // full index -> position in the free vector, or -1 (frozen)
    int32_t free_idx[MTG_MAX_PARAMS];     // free slot -> full index
    double  full_default[MTG_MAX_PARAMS];
    double  lo[MTG_MAX_PARAMS], hi[MTG_MAX_PARAMS];
};

// Host: compile an mtg_model_t.  Returns 0 or MTG_E*.
int mtg_build_plan(const mtg_model_t *m, Plan *out);

struct BatchArgs {
    int N, S, group, B;
    const double *t, *y, *diag, *ds_mean, *params;
    double *out_ll; double *out_grad; int32_t *status;
    uint32_t flags;
    const int32_t *index;   // optional active list: series id of work item i (K2 / L-BFGS)
    const int32_t *count;   // optional device-side number of work items (<= launch bound)
    double *per_step;       // optional K2 output [2][N][B]: alpha_n = (K^-1 r)_n and (K^-1)_nn  (GP.predict)
};

// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double plan_param(const Plan &pl, const double *__restrict__ x, int i)
{
    const int pos = pl.src_pos[i];
    return pos >= 0 ? x[pos] : pl.full_default[i];
}

// exp() for the once-per-series parameter maps: ONE out-of-line copy per kernel.  libdevice's exp inlines ~45 instructions
// per call site and the coefficient maps below have ~25 of them behind a switch: 1.4 k instructions of cold straight-line
// code that the latency kernels (K9 / K10 / K11, one evaluation = a few thousand instructions per warp) fetched through the
// instruction cache on every evaluation (ncu: stalled_no_instructions 29 %, profiles/r02a_k11_*).
static __device__ __noinline__ double mtg_exp(double x) { return exp(x); }

// diagonal jitter of term k, amplitude / rate of real slot j
__device__ __forceinline__ double plan_jitter(const Plan &pl, const double *__restrict__ x, int k)
{ const double p = plan_param(pl, x, pl.jit_idx[k]); return pl.jit_raw[k] ? p : mtg_exp(2.0 * p); }
__device__ __forceinline__ double plan_real_a(const Plan &pl, const double *__restrict__ x, int j)
{ const double p = plan_param(pl, x, pl.r_ia[j]); return pl.r_raw[j] ? p : mtg_exp(p); }
__device__ __forceinline__ double plan_real_c(const Plan &pl, const double *__restrict__ x, int j)
{ const double p = plan_param(pl, x, pl.r_ic[j]); return pl.r_raw[j] ? p : mtg_exp(p); }

// celerite Model.log_prior: 0 inside the closed box else -inf; the overrides the reference path can reach:
// BendingPowerlaw (celerite_models.py:85-90) and celerite's ComplexTerm with four parameters.  Returns true when inside.
// (Any other log_prior override is detected on the host, GP.device_prior_supported, and evaluated there.)
__device__ __forceinline__ bool plan_in_prior(const Plan &pl, const double *__restrict__ x)
{
    bool ok = true;
#pragma unroll 1
    for (int k = 0; k < pl.n_free; ++k) {
        const double v = x[k];
        ok = ok && (v >= pl.lo[k]) && (v <= pl.hi[k]);      // NaN -> outside
    }
#pragma unroll 1
    for (int k = 0; k < pl.n_bpl; ++k)
        ok = ok && !(plan_param(pl, x, pl.bpl_s0[k]) < plan_param(pl, x, pl.bpl_q[k]));
    // celerite ComplexTerm.log_prior with a fitted log_b: -inf when a c < b d (the term is then not positive definite)
#pragma unroll 1
    for (int j = 0; j < pl.jc; ++j)
        if (pl.c_kind[j] == CS_COMPLEX4) {
            const int *ix = pl.c_idx[j];
            ok = ok && !(plan_param(pl, x, ix[0]) + plan_param(pl, x, ix[2]) < plan_param(pl, x, ix[1]) + plan_param(pl, x, ix[3]));
        }
    return ok;
}

// Per-series coefficients in registers.  DUAL: SHO slots may hold two real terms (Q < 1/2).
template <int JR, int JC, bool DUAL>
struct Coef {
    double ar[JR > 0 ? JR : 1], ncr[JR > 0 ? JR : 1];                       // a, -c log2(e)
    double ac[JC > 0 ? JC : 1], bc[JC > 0 ? JC : 1], ncc[JC > 0 ? JC : 1], dc[JC > 0 ? JC : 1];
    double a2[(DUAL && JC > 0) ? JC : 1], nc2[(DUAL && JC > 0) ? JC : 1];   // second real term of an overdamped SHO
    bool   pair[(DUAL && JC > 0) ? JC : 1];                                 // true: complex pair
    double a_sum;                                                           // jitter + sum a
    bool   b_zero;                                                          // warp-uniform: all b == 0
};

template <int JR, int JC, bool DUAL>
__device__ __forceinline__ void plan_coefficients(const Plan &pl, const double *__restrict__ x,
                                                  Coef<JR, JC, DUAL> &c)
{
    double a_sum = 0.0;
    for (int k = 0; k < pl.n_jit; ++k) a_sum += plan_jitter(pl, x, k);
#pragma unroll
    for (int j = 0; j < JR; ++j) {
        c.ar[j] = plan_real_a(pl, x, j);
        c.ncr[j] = -plan_real_c(pl, x, j) * MTG_LOG2E;
        a_sum += c.ar[j];
    }
#pragma unroll
    for (int j = 0; j < JC; ++j) {
        const int kind = pl.c_kind[j];
        const double p0 = plan_param(pl, x, pl.c_idx[j][0]), p1 = plan_param(pl, x, pl.c_idx[j][1]);
        const double p2 = plan_param(pl, x, pl.c_idx[j][2]), p3 = plan_param(pl, x, pl.c_idx[j][3]);
        // the exponentials every kind needs, through the single out-of-line exp (Matern-3/2 takes e^{2 log_sigma}, e^{-log_rho})
        double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
        if (kind != CS_RAW) {
            e0 = mtg_exp(kind == CS_MATERN32 ? 2.0 * p0 : p0);
            e1 = mtg_exp(kind == CS_MATERN32 ? -p1 : p1);
            if (kind != CS_COSINUS && kind != CS_MATERN32) e2 = mtg_exp(p2);
            if (kind == CS_COMPLEX4) e3 = mtg_exp(p3);
        }
        double a, b = 0.0, cc, d;
        bool pair = true; double a2 = 0.0, c2 = 0.0;
        switch (kind) {
        case CS_COMPLEX3: a = e0; cc = e1; d = e2; break;
        case CS_COMPLEX4: a = e0; b = e1; cc = e2; d = e3; break;
        case CS_RAW:      a = p0; b = p1; cc = p2; d = p3; break;
        case CS_LORENTZIAN: a = e0; cc = 0.5 * e2 / e1; d = e2; break;
        case CS_COSINUS:  a = e0; cc = 0.0; d = e1; break;
        case CS_BENDINGPL: a = e0; b = e1; cc = e2; d = e2; break;
        case CS_MATERN32: { const double eps = pl.c_extra[j];
                          const double w0 = sqrt(3.0) * e1; const double S0 = e0 / w0;
                          a = w0 * S0; b = w0 * w0 * S0 / eps; cc = w0; d = eps; } break;
        default: /* CS_SHO */ {
                          const double S0 = e0, Q = e1, w0 = e2;
                          if (!DUAL || Q >= 0.5) {
                              const double f = sqrt(4.0 * Q * Q - 1.0);
                              a = S0 * w0 * Q; b = a / f; cc = 0.5 * w0 / Q; d = cc * f;
                          } else {
                              const double f = sqrt(1.0 - 4.0 * Q * Q);
                              pair = false;
                              a = 0.5 * S0 * w0 * Q * (1.0 + 1.0 / f); cc = 0.5 * w0 / Q * (1.0 - f);
                              a2 = 0.5 * S0 * w0 * Q * (1.0 - 1.0 / f); c2 = 0.5 * w0 / Q * (1.0 + f);
                              d = 0.0;
                          } } break;
        }
        c.ac[j] = a; c.bc[j] = b; c.ncc[j] = -cc * MTG_LOG2E; c.dc[j] = d;
        a_sum += a;
        if (DUAL) { c.pair[j] = pair; c.a2[j] = a2; c.nc2[j] = -c2 * MTG_LOG2E; a_sum += a2; }
    }
    c.a_sum = a_sum;
    c.b_zero = pl.b_zero != 0;
}

// mean model value at time tn (gpmodelling.py:62-124; models/mean_models.py:6-31)
struct MeanEval {
    int kind; double p0, p1, p2, p3;
    __device__ __forceinline__ double operator()(double tn) const {
        if (kind == MTG_MEAN_CONSTANT) return p0;
        if (kind == MTG_MEAN_LINEAR) return p0 * tn + p1;
        const double q = (tn - p0);
        return p2 / (2.0 * M_PI * p1) * exp(-q * q / (2.0 * p1 * p1)) + p3;     // mean_models.py:9-10
    }
};
__device__ __forceinline__ MeanEval plan_mean(const Plan &pl, const double *__restrict__ x)
{
    MeanEval m; m.kind = pl.mean_kind; m.p0 = m.p1 = m.p2 = m.p3 = 0.0;
    m.p0 = plan_param(pl, x, pl.mean_idx[0]);
    if (pl.mean_kind != MTG_MEAN_CONSTANT) m.p1 = plan_param(pl, x, pl.mean_idx[1]);
    if (pl.mean_kind == MTG_MEAN_GAUSSIAN) { m.p2 = plan_param(pl, x, pl.mean_idx[2]); m.p3 = plan_param(pl, x, pl.mean_idx[3]); }
    return m;
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_fence_init()
{ asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "MTG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra MTG_DONE;\n"
        "bra MTG_WAIT;\n"
        "MTG_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{ asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#endif  // __CUDACC__

// ---- host-side helpers ---------------------------------------------------------------------
int    mtg_k1_dispatch(const Plan &pl, const BatchArgs &a, cudaStream_t stream);
int    mtg_k9_dispatch(const Plan &pl, const BatchArgs &a, cudaStream_t stream);      // parallel-in-time scan (mtg_scan.cu)
bool   mtg_k9_supports(const Plan &pl);
int    mtg_k11_dispatch(const Plan &pl, const BatchArgs &a, cudaStream_t stream);     // scan gradient (mtg_scan_grad.cu); honours a.index / a.count
bool   mtg_k11_supports(const Plan &pl, int N);
long long mtg_scan_grad_max_batch();
size_t mtg_k2_tape_bytes_per_series(const Plan &pl, int N);
int    mtg_k2_chunk(const Plan &pl, int N, int B);
// a.B work items (series a.index[i] when a.index != NULL, limited to *a.count when a.count != NULL)
int    mtg_k2_run(const Plan &pl, const BatchArgs &a, double *tape, int chunk, cudaStream_t stream);
int  mtg_pool_setup();
void mtg_set_error(const char *fmt, ...);
int  mtg_cuda_check(cudaError_t e, const char *what);
void mtg_count_launch(uint64_t n = 1);
#define MTG_CUDA(call) do { int _rc = mtg_cuda_check((call), #call); if (_rc) return _rc; } while (0)
