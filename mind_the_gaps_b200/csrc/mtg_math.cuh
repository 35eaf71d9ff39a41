// mtg_math.cuh -- FP64 exp / sincos for the celerite recursion inner loop.
//
// Why not libdevice's exp() / sincos(): the ncu source page of K1 (profiles/r01_*) shows 118 FP64
// instructions per step but 246 issued instructions -- ~74 of the extra ones are IMAD.MOV / UMOV
// pairs that re-materialise the 64-bit polynomial coefficients libdevice embeds as immediates
// (DFMA has no 64-bit immediate form), which makes the kernel issue-bound rather than FP64-pipe
// bound.  Here every coefficient is a __constant__ operand (DFMA R, R, c[bank][off], R), the
// special-case branches are reduced to what the recursion can actually produce (exp of a
// non-positive argument; sincos of a finite phase), and accuracy stays at <= 1.3 ulp
// (tests/test_math_gpu.py).
#pragma once
#include <cuda_runtime.h>
#include "mtg_tables.cuh"

namespace mtgm {

// 2^r on |r| <= 1/2: degree-11 interpolant at Chebyshev nodes (truncation 3e-18, 1.1 ulp in Horner form)
static __constant__ double EXP2_C[12] = {
    1.0, 0.6931471805599453, 0.24022650695910158, 0.055504108664821625, 0.009618129107587256,
    0.001333355814640647, 0.00015403530463724353, 1.5252733841556773e-05, 1.3215432535912375e-06,
    1.0178057087733941e-07, 7.074194297288521e-09, 4.4558179083360645e-10};
// [0] 1.5 * 2^52, [1] flush-to-zero threshold
static __constant__ double EXP2_K[2] = {6755399441055744.0, -1020.0};
#ifndef MTG_LOG2E
#define MTG_LOG2E 1.4426950408889634
#define MTG_LN2 0.6931471805599453
#endif
// sin / cos kernels on |r| <= pi/4: minimax coefficients of fdlibm's __kernel_sin / __kernel_cos
static __constant__ double SIN_C[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03,
                                       -1.98412698298579493134e-04, 2.75573137070700676789e-06,
                                       -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double COS_C[6] = {4.16666666666666019037e-02, -1.38888888888741095749e-03,
                                       2.48015872894767294178e-05, -2.75573143513906633035e-07,
                                       2.08757232129817482790e-09, -1.13596475577881948265e-11};
// [0] 2/pi, [1..3] -pi/2 split in three doubles, [4] 1.5 * 2^52, [5] fast-path limit
static __constant__ double TRIG_K[6] = {0.6366197723675814, -1.5707963267948966, -6.123233995736766e-17,
                                        1.4973849048591698e-33, 6755399441055744.0, 1.0e5};

// 2^t for t <= 0.  The recursion needs exp(-c dt): callers keep -c log2(e) per series, so the range
// reduction is exact (t - round(t)) and needs no Cody-Waite constants.  Results below 2^-1020 flush to 0.
__device__ __forceinline__ double exp2_nonpos(double t)
{
    const double kd = t + EXP2_K[0];                        // round(t) in the low mantissa bits
    const int k = __double2loint(kd);
    const double r = t - (kd - EXP2_K[0]);                  // |r| <= 1/2, exact
    double p = EXP2_C[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, r, EXP2_C[i]);
    const int hi = __double2hiint(p) + (k << 20);           // p in [0.70, 1.42]: scale by 2^k in the exponent field
    p = __hiloint2double(hi, __double2loint(p));
    // t < -1020 decided on the high word (integer pipe): for t <= 0 a larger unsigned high word means a more negative t
    return (unsigned)__double2hiint(t) > 0xC08FE000u ? 0.0 : p;
}
// exp(x) for x <= 0 through the same kernel (test hook / occasional use)
__device__ __forceinline__ double exp_nonpos(double x) { return exp2_nonpos(x * MTG_LOG2E); }

// Fast-path limit of the sincos kernels: |x| < 2^40.  The reduction r = x - k s (s = pi/2 or pi/128 split in two
// doubles) is two FMAs: the first is the exactly rounded x - k s_hi (|r| is small, so its rounding error is
// 2^-53 |r|), the second adds k s_lo ~ |x| 2^-53; what is neglected is k s 2^-106 ~ |x| 2^-106, i.e. < 1e-19 up to
// 2^40, and k < 2^48 still sits in the low mantissa bits of k + 1.5 2^52.  (A first version stopped at 1e5 and sent
// larger phases -- omega0 t_n of a fast QPO on a decade-long baseline in seconds -- to libdevice's Payne-Hanek
// path, which stalls the whole warp for one lane: the ensemble sampler ran 2.5x slower once walkers got there.)
#define MTG_TRIG_FAST_LIMIT_HI 0x42700000
// sin(x), cos(x) for finite x; Cody-Waite reduction below 2^40, libdevice beyond
__device__ __forceinline__ void sincos_fast(double x, double *sn, double *cs)
{
    if ((__double2hiint(x) & 0x7fffffff) >= MTG_TRIG_FAST_LIMIT_HI) { sincos(x, sn, cs); return; }      // |x| >= 2^40 (or NaN/inf): libdevice
    const double kd = fma(x, TRIG_K[0], TRIG_K[4]);
    const int q = __double2loint(kd);
    const double kf = kd - TRIG_K[4];
    double r = fma(kf, TRIG_K[1], x);
    r = fma(kf, TRIG_K[2], r);                              // two FMA terms carry 106 bits of pi/2: enough below 1e5
    const double z = r * r;
    double ps = SIN_C[5], pc = COS_C[5];
#pragma unroll
    for (int i = 4; i >= 0; --i) { ps = fma(ps, z, SIN_C[i]); pc = fma(pc, z, COS_C[i]); }
    const double s = fma(r * z, ps, r);
    const double c = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double a = (q & 1) ? c : s, b = (q & 1) ? s : c;
    *sn = (q & 2) ? -a : a;
    *cs = ((q + 1) & 2) ? -b : b;
}


// ---- table-driven variants (the K1/K2 inner loop) ------------------------------------------------
// The polynomial versions above cost 15 (2^x) and ~21 (sincos) FP64-pipe instructions; K1 is bound by
// that pipe (profiles/r01c_*), so the inner loop trades FP64 work for one shared-memory lookup each:
//   2^t   = 2^k * T[j] * 2^r,  t = k + j/256 + r, |r| <= 1/512  -> degree-4 polynomial, 9 FP64 ops
//   cis x = cis(q pi/128) * cis(r), |r| <= pi/256                 -> sin: 2 terms, cos-1: 3 terms, 15 FP64 ops
// Tables are correctly rounded (tools/gen_tables.py) and copied global -> shared once per CTA.
static __constant__ double EXP2T_C[6] = {256.0, -1.0 / 256.0, MTG_EXP2_P1, MTG_EXP2_P2, MTG_EXP2_P3, MTG_EXP2_P4};
// [0] 128/pi [1] -pi/128 hi [2] -pi/128 lo [3] -1/6 [4] 1/120 [5] -1/2 [6] 1/24 [7] -1/720
static __constant__ double TRIGT_C[8] = {MTG_TRIG_INV_STEP, -MTG_TRIG_STEP_HI, -MTG_TRIG_STEP_LO,
                                         -1.0 / 6.0, 1.0 / 120.0, -0.5, 1.0 / 24.0, -1.0 / 720.0};

#define MTG_TABS_DOUBLES (MTG_EXP2_TAB_N + 2 * MTG_TRIG_TAB_N)
struct Tabs { const double *e2; const double *trig; };      // shared-memory copies

// every thread of the CTA calls this; the caller synchronises before the first lookup
__device__ __forceinline__ Tabs tabs_fill(double *smem /* MTG_TABS_DOUBLES, 16-byte aligned */)
{
    for (int i = threadIdx.x; i < MTG_EXP2_TAB_N; i += blockDim.x) smem[i] = EXP2_TAB[i];
    for (int i = threadIdx.x; i < 2 * MTG_TRIG_TAB_N; i += blockDim.x) smem[MTG_EXP2_TAB_N + i] = TRIG_TAB[i];
    Tabs t; t.e2 = smem; t.trig = smem + MTG_EXP2_TAB_N;
    return t;
}

// 2^t for t <= 0 (flushes to 0 below 2^-1020), <= 1.1 ulp
__device__ __forceinline__ double exp2_nonpos_tab(double t, const Tabs &tb)
{
    const double kd = fma(t, EXP2T_C[0], EXP2_K[0]);        // round(256 t) in the low mantissa bits
    const int ki = __double2loint(kd);
    const double r = fma(kd - EXP2_K[0], EXP2T_C[1], t);    // exact, |r| <= 1/512
    const double T = tb.e2[ki & (MTG_EXP2_TAB_N - 1)];
    double q = fma(r, EXP2T_C[5], EXP2T_C[4]);
    q = fma(q, r, EXP2T_C[3]);
    q = fma(q, r, EXP2T_C[2]);
    double p = fma(T, q * r, T);                            // in [0.998, 2.003)
    const int hi = __double2hiint(p) + ((ki >> 8) << 20);
    p = __hiloint2double(hi, __double2loint(p));
    return (unsigned)__double2hiint(t) > 0xC08FE000u ? 0.0 : p;
}

// sin(x), cos(x) for finite x, absolute error < 2e-16; libdevice beyond |x| = 2^40
__device__ __forceinline__ void sincos_tab(double x, const Tabs &tb, double *sn, double *cs)
{
    if ((__double2hiint(x) & 0x7fffffff) >= MTG_TRIG_FAST_LIMIT_HI) { sincos(x, sn, cs); return; }
    const double kd = fma(x, TRIGT_C[0], TRIG_K[4]);
    const int q = __double2loint(kd) & (MTG_TRIG_TAB_N - 1);
    const double kf = kd - TRIG_K[4];
    double r = fma(kf, TRIGT_C[1], x);
    r = fma(kf, TRIGT_C[2], r);                             // |r| <= pi/256
    const double2 CS = *reinterpret_cast<const double2 *>(tb.trig + 2 * q);
    const double z = r * r;
    const double sr = fma(r * z, fma(z, TRIGT_C[4], TRIGT_C[3]), r);                   // sin r
    const double cm = z * fma(z, fma(z, TRIGT_C[7], TRIGT_C[6]), TRIGT_C[5]);          // cos r - 1
    *sn = fma(CS.y, cm, fma(CS.x, sr, CS.y));
    *cs = fma(CS.x, cm, fma(-CS.y, sr, CS.x));
}

// ---- big-table variants (K1) --------------------------------------------------------------------------
// K1 is bound by the FP64 pipe at ~95 % of its issue rate (profiles/r01e_k1_*), so every FP64 instruction removed
// from the step is throughput.  With 2048-entry tables (48 KB of shared memory per CTA, which K1 can afford and K2 /
// K8 cannot) the residuals shrink to |r| <= 1/4096 and |r| <= pi/2048:
//   2^r - 1 = r (c1 + r (c2 + r c3))            truncation (ln2/4096)^4/24 = 3.4e-17     7 FP64 ops (was 8)
//   sin r   = r - r^3/6, cos r - 1 = -z/2 + z^2/24   truncation 7e-17 / 2e-20           13 FP64 ops (was 15)
static __constant__ double EXP2B_C[5] = {2048.0, -1.0 / 2048.0, MTG_EXP2_P1, MTG_EXP2_P2, MTG_EXP2_P3};
// [0] 1024/pi [1] -pi/1024 hi [2] -pi/1024 lo [3] -1/6 [4] -1/2 [5] 1/24
static __constant__ double TRIGB_C[6] = {MTG_TRIGB_INV_STEP, -MTG_TRIGB_STEP_HI, -MTG_TRIGB_STEP_LO, -1.0 / 6.0, -0.5, 1.0 / 24.0};

#define MTG_TABSB_DOUBLES (MTG_EXP2B_TAB_N + 2 * MTG_TRIGB_TAB_N)
struct TabsBig { const double *e2; const double *trig; };

__device__ __forceinline__ TabsBig tabs_fill_big(double *smem /* MTG_TABSB_DOUBLES, 16-byte aligned */)
{
    for (int i = threadIdx.x; i < MTG_EXP2B_TAB_N; i += blockDim.x) smem[i] = EXP2B_TAB[i];
    double2 *tr = reinterpret_cast<double2 *>(smem + MTG_EXP2B_TAB_N);
    const double2 *src = reinterpret_cast<const double2 *>(TRIGB_TAB);
    for (int i = threadIdx.x; i < MTG_TRIGB_TAB_N; i += blockDim.x) tr[i] = src[i];
    TabsBig t; t.e2 = smem; t.trig = smem + MTG_EXP2B_TAB_N;
    return t;
}

__device__ __forceinline__ double exp2_nonpos_tab(double t, const TabsBig &tb)
{
    const double kd = fma(t, EXP2B_C[0], EXP2_K[0]);        // round(2048 t) in the low mantissa bits
    const int ki = __double2loint(kd);
    const double r = fma(kd - EXP2_K[0], EXP2B_C[1], t);    // exact, |r| <= 1/4096
    const double T = tb.e2[ki & (MTG_EXP2B_TAB_N - 1)];
    double q = fma(r, EXP2B_C[4], EXP2B_C[3]);
    q = fma(q, r, EXP2B_C[2]);
    double p = fma(T, q * r, T);
    const int hi = __double2hiint(p) + ((ki >> 11) << 20);
    p = __hiloint2double(hi, __double2loint(p));
    return (unsigned)__double2hiint(t) > 0xC08FE000u ? 0.0 : p;
}

__device__ __forceinline__ void sincos_tab(double x, const TabsBig &tb, double *sn, double *cs)
{
    if ((__double2hiint(x) & 0x7fffffff) >= MTG_TRIG_FAST_LIMIT_HI) { sincos(x, sn, cs); return; }
    const double kd = fma(x, TRIGB_C[0], TRIG_K[4]);
    const int q = __double2loint(kd) & (MTG_TRIGB_TAB_N - 1);
    const double kf = kd - TRIG_K[4];
    double r = fma(kf, TRIGB_C[1], x);
    r = fma(kf, TRIGB_C[2], r);                             // |r| <= pi/2048
    const double2 CS = *reinterpret_cast<const double2 *>(tb.trig + 2 * q);
    const double z = r * r;
    const double sr = fma(r * z, TRIGB_C[3], r);                      // sin r
    const double cm = z * fma(z, TRIGB_C[5], TRIGB_C[4]);             // cos r - 1
    *sn = fma(CS.y, cm, fma(CS.x, sr, CS.y));
    *cs = fma(CS.x, cm, fma(-CS.y, sr, CS.x));
}

// 1/x for a positive normal x: MUFU.RCP64H seed (>= 19 bits) + one cubic Newton step, ~1 ulp.
// IEEE division costs three more FP64 instructions and a slow-path branch; D <= 0 / denormal / NaN
// are flagged by the caller independently of this value.
__device__ __forceinline__ double rcp_pos(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

}  // namespace mtgm
