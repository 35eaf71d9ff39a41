// mtg_sim_dev.cuh -- device helpers shared by the simulator translation units (mtg_sim.cu, compiled with
// -fmad=false, and mtg_tk95.cu): Philox streams, the celerite PSD, numpy's pairwise mean.
#pragma once
#include "mtg_sim.cuh"
#include "mtg_philox.cuh"

// two uniforms for (stream `purpose`, item, draw)
__device__ __forceinline__ double2 philox_u2(uint64_t seed, uint32_t purpose, uint64_t item, uint32_t draw)
{
    const uint4 r = philox4x32_10(make_uint4((uint32_t)item, (uint32_t)(item >> 32), draw, purpose),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return make_double2(u01_from_bits(r.x, r.y), u01_from_bits(r.z, r.w));
}
__device__ __forceinline__ double2 philox_normal2(uint64_t seed, uint32_t purpose, uint64_t item, uint32_t draw)
{
    const double2 u = philox_u2(seed, purpose, item, draw);
    const double r = sqrt(-2.0 * log(u.x));
    double s, c;
    sincospi(2.0 * u.y, &s, &c);
    return make_double2(r * c, r * s);
}

__device__ __forceinline__ double psd_eval(const PsdCoef &c, double w)
{
    const double w2 = w * w;
    double p = 0.0;
    for (int j = 0; j < c.nr; ++j) p += c.ar[j] * c.cr[j] / (c.cr[j] * c.cr[j] + w2);
    for (int j = 0; j < c.nc; ++j) {
        const double a = c.ac[j], b = c.bc[j], cc = c.cc[j], d = c.dc[j];
        const double w02 = cc * cc + d * d;
        p += ((a * cc + b * d) * w02 + (a * cc - b * d) * w2) / (w2 * w2 + 2.0 * (cc * cc - d * d) * w2 + w02 * w02);
    }
    return 0.79788456080286535588 * p;          // sqrt(2/pi)
}

// ---- fast helpers for the Philox path of K3 / K12 (statistical equivalence only; the injected-noise path keeps IEEE ops) ----
// Standard normal PAIR from 64 random bits by Box-Muller with FP32 / MUFU transcendentals.  The radius uses 40 bits:
// u1 = 2^-(z+1) * 1.m with z = leading zeros of the field (geometric), m the next 23 bits, so
// -2 ln u1 = 2 ln2 ((z + 1) - log2(1.m)) reaches 7.4 sigma (P = 1.4e-13 per deviate) although every floating-point step is
// single precision; the angle uses 24 bits.  Relative error of a deviate ~1e-6: irrelevant for a random draw.
__device__ __forceinline__ float2 normal_pair_from_bits(uint32_t hi, uint32_t lo)
{
    const uint64_t bits = ((uint64_t)hi << 32) | (lo & 0xff000000u);                    // 40 random bits, left aligned
    const int z = bits ? __clzll((long long)bits) : 40;
    const uint64_t norm = bits << z;                                   // top bit set (or 0)
    const float mant = __uint_as_float(0x3f800000u | (uint32_t)((norm >> 40) & 0x7fffffu));       // [1, 2)
    const float neg_log2_u = (float)(z + 1) - __log2f(mant) + 2.0e-8f;                     // > 0
    const float r2 = 1.3862943611198906f * neg_log2_u;                                      // -2 ln u1
    const float rad = r2 * rsqrtf(r2);
    const float ang = ((float)(lo & 0xffffffu) - 8388607.5f) * 3.7450704e-07f;              // (-pi, pi) in steps of 2 pi 2^-24
    return make_float2(rad * __cosf(ang), rad * __sinf(ang));
}
// The spectrum stream of the simulator: the half-spectrum entry kk = c + P r (column c < P, row r) of lightcurve `lc`
// takes half (r & 1) of the Philox block (lc, (r >> 1) P + c): rows 2j and 2j + 1 of a column share one block.  P is the
// pair stride of the simulator (N2 of the fused geometry, where a thread of K12 owns adjacent rows of a column; 1 otherwise).
__device__ __forceinline__ uint64_t spectrum_block(uint64_t lc, int M, int P, unsigned r, unsigned c)
{ return lc * (uint64_t)(M + P) + (uint64_t)(r >> 1) * (uint64_t)P + c; }
__device__ __forceinline__ void spectrum_normals(uint64_t seed, uint64_t block, float2 *even_row, float2 *odd_row)
{
    const uint4 q = philox4x32_10(make_uint4((uint32_t)block, (uint32_t)(block >> 32), 0u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    *even_row = normal_pair_from_bits(q.x, q.y);
    *odd_row = normal_pair_from_bits(q.z, q.w);
}
// 1/x and sqrt(x) for x in the single-precision range from an FP32 seed + two Newton steps (<= 2 ulp); IEEE otherwise
__device__ __forceinline__ double fast_rcp(double x)
{
    if (!(fabs(x) > 1e-30 && fabs(x) < 1e30)) return 1.0 / x;
    double y = (double)(1.0f / (float)x);
    y = fma(y, fma(-x, y, 1.0), y);
    return fma(y, fma(-x, y, 1.0), y);
}
__device__ __forceinline__ double fast_sqrt(double x)
{
    if (!(x > 1e-30 && x < 1e30)) return sqrt(x);
    double y = (double)rsqrtf((float)x);                 // ~ 1/sqrt(x)
    y = fma(0.5 * y, fma(-x * y, y, 1.0), y);
    y = fma(0.5 * y, fma(-x * y, y, 1.0), y);
    const double g = x * y;                              // sqrt(x), then one correction step
    return fma(0.5 * y, fma(-g, g, x), g);
}
__device__ __forceinline__ double psd_eval_fast(const PsdCoef &c, double w)
{
    const double w2 = w * w;
    double p = 0.0;
    for (int j = 0; j < c.nr; ++j) p += c.ar[j] * c.cr[j] * fast_rcp(c.cr[j] * c.cr[j] + w2);
    for (int j = 0; j < c.nc; ++j) {
        const double a = c.ac[j], b = c.bc[j], cc = c.cc[j], d = c.dc[j];
        const double w02 = cc * cc + d * d;
        p += ((a * cc + b * d) * w02 + (a * cc - b * d) * w2) * fast_rcp(w2 * w2 + 2.0 * (cc * cc - d * d) * w2 + w02 * w02);
    }
    return 0.79788456080286535588 * p;          // sqrt(2/pi)
}

// numpy's pairwise summation (npy pairwise_sum, block 128, 8 accumulators)
static __device__ double np_pairwise_sum(const double *a, int n)
{
    if (n < 8) {
        double res = -0.0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
            r0 += a[i]; r1 += a[i + 1]; r2 += a[i + 2]; r3 += a[i + 3];
            r4 += a[i + 4]; r5 += a[i + 5]; r6 += a[i + 6]; r7 += a[i + 7];
        }
        double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}
// np.mean of the fancy-indexed (k, 1) slice the reference averages (simulator.py:362-364): numpy's
// add.reduce is the pairwise sum over all k elements, then a true divide by k (verified against
// numpy 2.3 for k = 1..1000 in tests/test_simulator_gpu.py)
__device__ __forceinline__ double np_mean(const double *a, int n)
{
    if (n <= 0) return nan("");
    return np_pairwise_sum(a, n) / (double)n;
}

