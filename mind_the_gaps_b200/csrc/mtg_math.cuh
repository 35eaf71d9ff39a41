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

// sin(x), cos(x) for finite x; Cody-Waite three-term reduction below 1e5, libdevice beyond
__device__ __forceinline__ void sincos_fast(double x, double *sn, double *cs)
{
    if ((__double2hiint(x) & 0x7fffffff) >= 0x40F86A00) { sincos(x, sn, cs); return; }      // |x| >= 1e5 (or NaN/inf): libdevice
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

}  // namespace mtgm
