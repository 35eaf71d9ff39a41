// mtg_tk95.cu -- K12 / K13: Timmer & Koenig synthesis fused with the exposure binning (sm_100a).
//
// Replaces, for Gaussian-PDF simulations whose output is the BINNED lightcurve, the chain
//   k_spectrum -> cuFFT Z2D (n_fft points) -> k_bin
// of mtg_sim.cu, i.e. get_fft + np.fft.irfft + downsample of the reference
// (mind_the_gaps/simulator.py:468-501, :386-393, :340-367).
//
// Why: the reference synthesises extension_factor x the observing baseline at min(exposure) / aliasing_factor
// resolution (n_fft ~ 2.7e6 for config C3 / C4) and then keeps the ~2-3 samples that fall inside each of the N ~ 1000
// exposures of ONE random segment: 99.9 % of the inverse FFT's output is never read.  cuFFT cannot prune, so the
// chain above moves ~75 MB (43 MB of it the spectrum and the series, three to five times over inside the library)
// per lightcurve through HBM.  Here the transform is split as n = N1 x N2 (four-step / Cooley-Tukey,
// k = N2 k1 + k2, m = m1 + N1 m2):
//
//   x[m] = sum_{k2 < N2} w_n^(m k2) A_k2[m mod N1],      A_k2[m1] = sum_{k1 < N1} X[N2 k1 + k2] w_N1^(m1 k1)
//
//   K12 k_tk95_columns  one CTA per (lightcurve, column k2): the N1 = 8192 spectrum entries of the column are
//        GENERATED in shared memory (Philox normals x sqrt(S(w)/2); entries above Nyquist as the conjugate of their
//        mirror, so every half-spectrum entry is drawn exactly once), transformed there by an in-place radix-8
//        decimation-in-frequency FFT (padded layout, twiddle tables in shared memory, last 16 points in registers),
//        and only the entries A_k2[m mod N1] of the <= Qmax samples this lightcurve's segment needs are written:
//        Y[l][k2][q].  The Hermitian symmetry A_(N2-k2) = w_N1^(-m1) conj(A_k2) halves the columns: K2 = N2/2 + 1.
//   K13 k_tk95_rows     one thread per needed sample: x = Re sum_k2 c_k2 w_n^(m k2) Y[l][k2][q]  (a length-K2 dot
//        product with a geometric twiddle; coalesced over q), then the reference's per-bin mean (numpy pairwise order).
//
// HBM traffic per lightcurve: ~2 x Q K2 16 B (6-7 MB written and read back for Q ~ 2500, K2 = 165) instead of ~75 MB;
// the work is bound by the FP64 pipe (FFT butterflies + PSD evaluation) and the integer pipe (Philox).
// Index work (segment, bin membership) stays in mtg_sim.cu (k_segment, k_need; compiled without FMA contraction).
#include "mtg_sim_dev.cuh"
#include <cmath>

#define TK_THREADS 512

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
__device__ __forceinline__ double2 cmuli(double2 a) { return make_double2(-a.y, a.x); }         // a * i

// y_p = sum_q v_q i^(p q)   (inverse-transform sign)
__device__ __forceinline__ void dft4(double2 &v0, double2 &v1, double2 &v2, double2 &v3)
{
    const double2 e0 = cadd(v0, v2), e1 = cadd(v1, v3), o0 = csub(v0, v2), o1 = cmuli(csub(v1, v3));
    v0 = cadd(e0, e1); v2 = csub(e0, e1); v1 = cadd(o0, o1); v3 = csub(o0, o1);
}
// v_p <- sum_q v_q w8^(p q), w8 = exp(+i pi/4); natural order in and out
__device__ __forceinline__ void dft8(double2 (&v)[8])
{
    const double h = 0.70710678118654752440;
    double2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
    double2 b0 = csub(v[0], v[4]);
    const double2 d1 = csub(v[1], v[5]), d2 = csub(v[2], v[6]), d3 = csub(v[3], v[7]);
    double2 b1 = make_double2((d1.x - d1.y) * h, (d1.x + d1.y) * h);
    double2 b2 = cmuli(d2);
    double2 b3 = make_double2((-d3.x - d3.y) * h, (d3.x - d3.y) * h);
    dft4(a0, a1, a2, a3);          // even outputs 0, 2, 4, 6
    dft4(b0, b1, b2, b3);          // odd outputs 1, 3, 5, 7
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// padded shared-memory layout: one 16-byte pad per 16 points, so that the register FFT on 16 CONTIGUOUS points
// (one thread per block of 16) and the strided radix passes are both free of bank conflicts
__device__ __forceinline__ int tk_phys(int i) { return i + (i >> 4); }

// position of frequency m1 after the DIF passes (digit reversal over the pass radices, natural order in the last block)
__device__ __forceinline__ int tk_pos(const TkGeom &g, int m1)
{
    int pos = 0, shift = g.logN1;
    for (int s = 0; s < g.npass; ++s) {
        shift -= g.lograd[s];
        pos |= (m1 & ((1 << g.lograd[s]) - 1)) << shift;
        m1 >>= g.lograd[s];
    }
    return pos | m1;
}

// sqrt(x) for x in the single-precision range: FP32 rsqrt seed, one Newton step on 1/sqrt, one correction of the root
// (error 1e-7 -> 1.5e-14 -> below 1 ulp); IEEE otherwise
__device__ __forceinline__ double tk_sqrt(double x)
{
    if (!(x > 1e-30 && x < 1e30)) return sqrt(x);
    double y = (double)rsqrtf((float)x);
    y = fma(0.5 * y, fma(-x * y, y, 1.0), y);
    const double r = x * y;
    return fma(0.5 * y, fma(-r, r, x), r);
}

// ---- celerite PSD of one lightcurve, pre-scaled for the inner loop of K12 ------------------------------------------
// amp^2(w) = fold^2/2 sqrt(2/pi) [ sum_j a_j c_j / (c_j^2 + w^2) + sum_j ((a c + b d) w0^2 + (a c - b d) w^2) / (w^4 + 2 (c^2 - d^2) w^2 + w0^4) ]
// (psd_eval of mtg_sim_dev.cuh = celerite's get_psd) with every constant product formed once per lightcurve.
struct TkPsd { int nr, nc; double ra[MTG_MAXR + 2 * MTG_MAXC], rc2[MTG_MAXR + 2 * MTG_MAXC], cA1[MTG_MAXC], cA2[MTG_MAXC], cB1[MTG_MAXC], cB0[MTG_MAXC]; };

// 1/x and sqrt(x) for positive normal x from the MUFU.RCP64H / RSQ64H seeds (~2^-20): no conversions, no range tests
__device__ __forceinline__ double tk_rcp_pos(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}
__device__ __forceinline__ double tk_sqrt_pos(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x * y, 0.5 * y, 0.5), y);                   // Newton on 1/sqrt: 2^-20 -> ~1e-12
    const double r = x * y;
    return fma(0.5 * y, fma(-r, r, x), r);                       // correction of the root: below 1 ulp
}
__device__ __forceinline__ double tk_amp(const TkPsd &p, double w2)
{
    double acc = 0.0;
    for (int j = 0; j < p.nr; ++j) acc = fma(p.ra[j], tk_rcp_pos(p.rc2[j] + w2), acc);
    for (int j = 0; j < p.nc; ++j)
        acc = fma(fma(p.cA2[j], w2, p.cA1[j]), tk_rcp_pos(fma(w2, w2 + p.cB1[j], p.cB0[j])), acc);
    return acc > 0.0 ? tk_sqrt_pos(acc) : 0.0;
}

// the same with the first two real and the first complex term held in registers (the kernels of configs C1-C4)
struct TkAmpRegs {
    int nr, nc;
    double ra0, rc0, ra1, rc1, cA1, cA2, cB1, cB0;
    __device__ __forceinline__ void load(const TkPsd &p)
    {
        nr = p.nr; nc = p.nc;
        ra0 = nr > 0 ? p.ra[0] : 0.0; rc0 = nr > 0 ? p.rc2[0] : 1.0;
        ra1 = nr > 1 ? p.ra[1] : 0.0; rc1 = nr > 1 ? p.rc2[1] : 1.0;
        cA1 = nc > 0 ? p.cA1[0] : 0.0; cA2 = nc > 0 ? p.cA2[0] : 0.0; cB1 = nc > 0 ? p.cB1[0] : 0.0; cB0 = nc > 0 ? p.cB0[0] : 1.0;
    }
    __device__ __forceinline__ double amp(const TkPsd &p, double w2) const
    {
        double acc = ra0 * tk_rcp_pos(rc0 + w2);
        if (nr > 1) acc = fma(ra1, tk_rcp_pos(rc1 + w2), acc);
        if (nc > 0) acc = fma(fma(cA2, w2, cA1), tk_rcp_pos(fma(w2, w2 + cB1, cB0)), acc);
        for (int j = 2; j < nr; ++j) acc = fma(p.ra[j], tk_rcp_pos(p.rc2[j] + w2), acc);
        for (int j = 1; j < nc; ++j) acc = fma(fma(p.cA2[j], w2, p.cA1[j]), tk_rcp_pos(fma(w2, w2 + p.cB1[j], p.cB0[j])), acc);
        return acc > 0.0 ? tk_sqrt_pos(acc) : 0.0;
    }
};

// amplitude table of a host-evaluated PSD in the column order of K12: amp_t[k2][k1] = sqrt(S(w_kk) / 2) * fold with
// kk = N2 k1 + k2 folded below Nyquist (0 at kk = 0); built once per generate call, read coalesced by every lightcurve
__global__ void __launch_bounds__(256)
k_tk95_amp_table(const __grid_constant__ TkGeom g, double fold, const double *__restrict__ psd_table, double *__restrict__ amp_t)
{
    const int k1 = blockIdx.x * blockDim.x + threadIdx.x, k2 = blockIdx.y;
    if (k1 >= g.N1) return;
    const unsigned k = (unsigned)g.N2 * (unsigned)k1 + (unsigned)k2;
    const unsigned kk = k > (unsigned)(g.n >> 1) ? (unsigned)g.n - k : k;
    amp_t[(size_t)k2 * g.N1 + k1] = kk > 0 ? tk_sqrt(0.5 * psd_table[kk]) * fold : 0.0;
}

// The transform of one column (natural order in, tk_pos order out) by NT threads, then the gather of the needed entries to
// yo[q].  `sync` is the barrier of the NT participating threads.  STAGED: the column arrives as single-precision pairs in
// `stage` and the first pass runs out of place (stage -> data) with its twiddles read from `tw0` in global memory;
// `after_first` runs once the first pass no longer needs `stage`.  Otherwise the column is in `data` and tw0 = tw + tw_off[0].
template <int NT, bool STAGED, class Sync, class After>
__device__ __forceinline__ void tk_transform_and_keep(const TkGeom &g, double2 *data, const float2 *stage, const double2 *tw0,
                                                      const double2 *tw, const unsigned short *spos, double2 *__restrict__ yo,
                                                      int tid, Sync sync, After after_first)
{
    const int N1 = g.N1;
    auto ld = [&](bool from_stage, int i) -> double2 {
        if (STAGED && from_stage) { const float2 f = stage[i]; return make_double2((double)f.x, (double)f.y); }
        return data[tk_phys(i)];
    };
    // ---- in-place decimation-in-frequency passes ------------------------------------------------------
    int logL = g.logN1;
    for (int s = 0; s < g.npass; ++s) {
        const int lr = g.lograd[s], logLs = logL - lr, Ls = 1 << logLs;
        const bool first = s == 0;
        const double2 *W1 = first ? tw0 : tw + g.tw_off[s], *W2 = W1 + Ls, *W4 = W2 + Ls;
        if (lr == 3) {
            for (int t = tid; t < (N1 >> 3); t += NT) {
                const int j = t & (Ls - 1), base = ((t >> logLs) << logL) + j;
                double2 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = ld(first, base + (q << logLs));
                dft8(v);
                const double2 w1 = W1[j], w2 = W2[j], w4 = W4[j];
                const double2 w3 = cmul(w1, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2);
                const double2 w7 = cmul(w4, w3);
                data[tk_phys(base)] = v[0];
                data[tk_phys(base + (1 << logLs))] = cmul(v[1], w1);
                data[tk_phys(base + (2 << logLs))] = cmul(v[2], w2);
                data[tk_phys(base + (3 << logLs))] = cmul(v[3], w3);
                data[tk_phys(base + (4 << logLs))] = cmul(v[4], w4);
                data[tk_phys(base + (5 << logLs))] = cmul(v[5], w5);
                data[tk_phys(base + (6 << logLs))] = cmul(v[6], w6);
                data[tk_phys(base + (7 << logLs))] = cmul(v[7], w7);
            }
        } else {                                             // radix 4
            for (int t = tid; t < (N1 >> 2); t += NT) {
                const int j = t & (Ls - 1), base = ((t >> logLs) << logL) + j;
                double2 v0 = ld(first, base), v1 = ld(first, base + Ls);
                double2 v2 = ld(first, base + 2 * Ls), v3 = ld(first, base + 3 * Ls);
                dft4(v0, v1, v2, v3);
                const double2 w1 = W1[j], w2 = W2[j];
                data[tk_phys(base)] = v0;
                data[tk_phys(base + Ls)] = cmul(v1, w1);
                data[tk_phys(base + 2 * Ls)] = cmul(v2, w2);
                data[tk_phys(base + 3 * Ls)] = cmul(v3, cmul(w1, w2));
            }
        }
        logL = logLs;
        sync();
        if (first) after_first();
    }
    // ---- last F = 8 | 16 contiguous points per thread, in registers, natural order out ----------------------------
    if (g.F == 16) {
        for (int t = tid; t < (N1 >> 4); t += NT) {
            double2 *blk = data + t * 17;                    // tk_phys(16 t)
            double2 a[8], b[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double2 lo = blk[q], hi = blk[q + 8];
                a[q] = cadd(lo, hi); b[q] = csub(lo, hi);
            }
            // b_q *= w16^q
            const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173, h = 0.70710678118654752440;
            b[1] = cmul(b[1], make_double2(c1, s1));
            b[2] = make_double2((b[2].x - b[2].y) * h, (b[2].x + b[2].y) * h);
            b[3] = cmul(b[3], make_double2(s1, c1));
            b[4] = cmuli(b[4]);
            b[5] = cmul(b[5], make_double2(-s1, c1));
            b[6] = make_double2((-b[6].x - b[6].y) * h, (b[6].x - b[6].y) * h);
            b[7] = cmul(b[7], make_double2(-c1, s1));
            dft8(a); dft8(b);
#pragma unroll
            for (int r = 0; r < 8; ++r) { blk[2 * r] = a[r]; blk[2 * r + 1] = b[r]; }
        }
    } else {
        for (int t = tid; t < (N1 >> 3); t += NT) {
            double2 v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = data[tk_phys(8 * t + q)];
            dft8(v);
#pragma unroll
            for (int q = 0; q < 8; ++q) data[tk_phys(8 * t + q)] = v[q];
        }
    }
    sync();
    // ---- keep what this lightcurve's bins need: Y[l][k2][q] = A_k2[m_q mod N1] ---------------------------
#pragma unroll 4
    for (int q = tid; q < g.Qmax; q += NT) {
        const unsigned pos = spos[q];
        if (pos != 0xffffu) yo[q] = data[pos];
    }
}

// K12.  PARITY: IEEE operations on injected normals (tests); otherwise Philox normals and either the amplitude table
// (host-evaluated PSD) or the celerite PSD of this lightcurve's parameters.
// Persistent CTAs, each a CONTIGUOUS range of work items w -> (l, k2) = (w / K2, w % K2): the twiddle tables are loaded
// once per CTA, the PSD coefficients and the shared-memory positions of the needed samples once per lightcurve.
template <bool PARITY>
__global__ void __launch_bounds__(TK_THREADS, 1)
k_tk95_columns(const __grid_constant__ TkGeom g, int Lc, double val, double fold, const PsdCoef *__restrict__ coefs,
               const double *__restrict__ psd_table, const double *__restrict__ amp_t, const double *__restrict__ normals,
               uint64_t seed, uint64_t lc_offset, const double2 *__restrict__ tw_g, const int *__restrict__ mslot,
               double2 *__restrict__ Y)
{
    extern __shared__ __align__(16) double2 tk_smem[];
    double2 *data = tk_smem;                                     // N1 + N1/16 points
    double2 *tw = tk_smem + (g.N1 + (g.N1 >> 4) + 1);            // twiddle tables of the passes
    unsigned short *spos = reinterpret_cast<unsigned short *>(tw + g.tw_total);       // Qmax positions in `data` (0xffff: unused slot)
    __shared__ PsdCoef sc;
    __shared__ TkPsd ps;
    const int tid = threadIdx.x;
    const int N1 = g.N1, n = g.n, M = g.M;
    const unsigned half = (unsigned)(n >> 1), uN2 = (unsigned)g.N2;
    const double w_scale = (val * 2.0) * 3.141592653589793;
    for (int i = tid; i < g.tw_total; i += TK_THREADS) tw[i] = tw_g[i];
    int l_loaded = -1;
    const long long work = (long long)Lc * g.K2;
    const long long per = (work + gridDim.x - 1) / gridDim.x;
    const long long w0 = (long long)blockIdx.x * per, w1 = w0 + per < work ? w0 + per : work;
    for (long long w = w0; w < w1; ++w) {
        const int l = (int)(w / g.K2), k2 = (int)(w % g.K2);
        __syncthreads();                                         // previous item's gather is done with `data`, `spos` and the PSD
        if (l != l_loaded) {
            if (coefs) {
                if (PARITY) {
                    if (tid < (int)(sizeof(PsdCoef) / sizeof(int)))
                        reinterpret_cast<int *>(&sc)[tid] = reinterpret_cast<const int *>(coefs + l)[tid];
                } else {
                    const PsdCoef &c = coefs[l];
                    const double K = (0.5 * fold * fold) * 0.79788456080286535588;
                    if (tid == 0) { ps.nr = c.nr; ps.nc = c.nc; }
                    if (tid < c.nr) { ps.ra[tid] = K * c.ar[tid] * c.cr[tid]; ps.rc2[tid] = c.cr[tid] * c.cr[tid]; }
                    if (tid >= 32 && tid - 32 < c.nc) {
                        const int j = tid - 32;
                        const double a = c.ac[j], b = c.bc[j], cc = c.cc[j], d = c.dc[j], w02 = cc * cc + d * d;
                        ps.cA1[j] = K * (a * cc + b * d) * w02; ps.cA2[j] = K * (a * cc - b * d);
                        ps.cB1[j] = 2.0 * (cc * cc - d * d); ps.cB0[j] = w02 * w02;
                    }
                }
            }
            const int *ms = mslot + (size_t)l * g.Qmax;
            for (int q = tid; q < g.Qmax; q += TK_THREADS) {
                const int m = ms[q];
                spos[q] = m >= 0 ? (unsigned short)tk_phys(tk_pos(g, m & (N1 - 1))) : (unsigned short)0xffff;
            }
            l_loaded = l;
            if (coefs) __syncthreads();
        }
        // ---- generate the column X[N2 k1 + k2], k1 < N1 (natural order) -------------------------------
        if (PARITY) {
            for (int k1 = tid; k1 < N1; k1 += TK_THREADS) {
                const unsigned k = uN2 * (unsigned)k1 + (unsigned)k2;
                const bool mirror = k > half;
                const unsigned kk = mirror ? (unsigned)n - k : k;
                double2 X = make_double2(0.0, 0.0);
                if (kk > 0) {
                    const double wk = (((double)kk * val) * 2.0) * 3.141592653589793;
                    const double psd = coefs ? psd_eval(sc, wk) : psd_table[kk];
                    const double re = normals[((size_t)l * 2) * M + kk], im = normals[((size_t)l * 2 + 1) * M + kk];
                    const double amp = sqrt(0.5 * psd) * fold;
                    X.x = re * amp; X.y = im * amp;
                    if (kk == half) X.y = 0.0;                   // Nyquist bin is real (simulator.py:499-500)
                    if (mirror) X.y = -X.y;
                }
                data[tk_phys(k1)] = X;
            }
        } else {
            // a thread owns rows k1, k1 + 1 of the column: one Philox block feeds both (spectrum_block), entries above Nyquist
            // are the conjugates of their mirrors in column N2 - k2
            const uint64_t lcg = lc_offset + (uint64_t)l;
            const double *amp_col = amp_t ? amp_t + (size_t)k2 * N1 : nullptr;
            for (int k1 = 2 * tid; k1 < N1; k1 += 2 * TK_THREADS) {
                unsigned kk[2], row[2], col[2]; bool mir[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const unsigned r = (unsigned)(k1 + e), k = uN2 * r + (unsigned)k2;
                    mir[e] = k > half;
                    kk[e] = mir[e] ? (unsigned)n - k : k;
                    row[e] = !mir[e] ? r : (k2 == 0 ? (unsigned)N1 - r : (unsigned)N1 - 1u - r);
                    col[e] = (!mir[e] || k2 == 0) ? (unsigned)k2 : uN2 - (unsigned)k2;
                }
                const uint64_t blk0 = spectrum_block(lcg, M, g.N2, row[0], col[0]), blk1 = spectrum_block(lcg, M, g.N2, row[1], col[1]);
                float2 ze, zo, z[2];
                spectrum_normals(seed, blk0, &ze, &zo);
                z[0] = (row[0] & 1u) ? zo : ze;
                if (blk1 != blk0) spectrum_normals(seed, blk1, &ze, &zo);
                z[1] = (row[1] & 1u) ? zo : ze;
                double amp[2];
                if (amp_col) { const double2 a = *reinterpret_cast<const double2 *>(amp_col + k1); amp[0] = a.x; amp[1] = a.y; }
                else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) { const double wk = (double)kk[e] * w_scale; amp[e] = tk_amp(ps, wk * wk); }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double2 X = make_double2((double)z[e].x * amp[e], (double)z[e].y * amp[e]);
                    if (kk[e] == half) X.y = 0.0;                // Nyquist bin is real (simulator.py:499-500)
                    if (mir[e]) X.y = -X.y;
                    if (kk[e] == 0u) X = make_double2(0.0, 0.0);
                    data[tk_phys(k1 + e)] = X;
                }
            }
        }
        __syncthreads();
        tk_transform_and_keep<TK_THREADS, false>(g, data, nullptr, tw + g.tw_off[0], tw, spos, Y + ((size_t)l * g.K2 + k2) * g.Qmax, tid,
                                                 [] { __syncthreads(); }, [] {});
    }
}

// ---- K12, warp-specialised (the production kernel of the Philox path) -------------------------------------------------
// Generation is integer / FP32 work (Philox, Box-Muller), the transform is FP64 + shared-memory work: run back to back by
// the same warps (k_tk95_columns) each leaves the other's pipes idle.  Here warps 8-15 GENERATE column i + 1 into a
// single-precision staging buffer (a Philox deviate carries no more than single precision) while warps 0-7 TRANSFORM
// column i; the first transform pass runs out of place, staging buffer -> column buffer, and frees the staging buffer
// for the next column.  Hand-over by two named barriers (bar.arrive / bar.sync):
//   FULL (1): the producers filled the staging buffer       FREE (2): the consumers' first pass has read it
// Shared memory: column 136 KB + staging 64 KB + twiddles of passes >= 2 (7 KB; the first pass reads its 48 KB of twiddles
// from global memory, L2-resident) + slot positions.
#define TK_WS_HALF 256
__device__ __forceinline__ void tk_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tk_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__global__ void __launch_bounds__(2 * TK_WS_HALF, 1)
k_tk95_columns_ws(const __grid_constant__ TkGeom g, int Lc, double val, double fold, const PsdCoef *__restrict__ coefs,
                  const double *__restrict__ amp_t, uint64_t seed, uint64_t lc_offset, const double2 *__restrict__ tw_g,
                  const int *__restrict__ mslot, double2 *__restrict__ Y)
{
    extern __shared__ __align__(16) double2 tk_smem[];
    const int N1 = g.N1, n = g.n, M = g.M;
    const int tw_skip = g.npass > 1 ? g.tw_off[1] : g.tw_total;          // the first pass's tables stay in global memory
    double2 *data = tk_smem;
    double2 *tw_s = tk_smem + (N1 + (N1 >> 4) + 1);
    float2 *stage = reinterpret_cast<float2 *>(tw_s + (g.tw_total - tw_skip));
    unsigned short *spos = reinterpret_cast<unsigned short *>(stage + N1);
    __shared__ TkPsd ps;
    for (int i = threadIdx.x; i < g.tw_total - tw_skip; i += 2 * TK_WS_HALF) tw_s[i] = tw_g[tw_skip + i];
    __syncthreads();
    const long long work = (long long)Lc * g.K2;
    const long long per = (work + gridDim.x - 1) / gridDim.x;
    const long long w0 = (long long)blockIdx.x * per, w1 = w0 + per < work ? w0 + per : work;
    if (w0 >= w1) return;
    int l_loaded = -1;
    if (threadIdx.x >= TK_WS_HALF) {
        // ================= producers =================
        const int p = threadIdx.x - TK_WS_HALF;
        const unsigned half = (unsigned)(n >> 1), uN2 = (unsigned)g.N2;
        const double w_scale = (val * 2.0) * 3.141592653589793;
        for (long long w = w0; w < w1; ++w) {
            const int l = (int)(w / g.K2), k2 = (int)(w % g.K2);
            if (coefs && l != l_loaded) {
                tk_bar_sync(4, TK_WS_HALF);                      // every producer is done with the previous lightcurve's PSD
                const PsdCoef &c = coefs[l];
                const double K = (0.5 * fold * fold) * 0.79788456080286535588;
                if (p == 0) { ps.nr = c.nr; ps.nc = c.nc; }
                if (p < c.nr) { ps.ra[p] = K * c.ar[p] * c.cr[p]; ps.rc2[p] = c.cr[p] * c.cr[p]; }
                if (p >= 32 && p - 32 < c.nc) {
                    const int j = p - 32;
                    const double a = c.ac[j], b = c.bc[j], cc = c.cc[j], d = c.dc[j], w02 = cc * cc + d * d;
                    ps.cA1[j] = K * (a * cc + b * d) * w02; ps.cA2[j] = K * (a * cc - b * d);
                    ps.cB1[j] = 2.0 * (cc * cc - d * d); ps.cB0[j] = w02 * w02;
                }
                tk_bar_sync(4, TK_WS_HALF);
            }
            l_loaded = l;
            const uint64_t lcg = lc_offset + (uint64_t)l;
            const double *amp_col = amp_t ? amp_t + (size_t)k2 * N1 : nullptr;
            tk_bar_sync(2, 2 * TK_WS_HALF);                      // FREE: the consumers have read the previous column out of `stage`
            if (k2 != 0 && N1 >= 4 * TK_WS_HALF) {
                // Fast path (every column but k2 = 0).  Rows k1 < N1/2 are below Nyquist: entry kk = N2 k1 + k2, stream row k1
                // of column k2.  Rows k1 >= N1/2 are the conjugates of kk = n - N2 k1 - k2 = N2 (N1 - 1 - k1) + (N2 - k2):
                // stream row N1 - 1 - k1 of column N2 - k2.  A thread's pair (k1, k1 + 1) shares one Philox block either way;
                // two pairs per iteration give the integer pipe two independent Philox chains.
                const uint64_t base_lo = lcg * (uint64_t)(M + g.N2) + (uint64_t)k2, base_hi = lcg * (uint64_t)(M + g.N2) + (uint64_t)(uN2 - (unsigned)k2);
                TkAmpRegs ar;
                if (!amp_col) ar.load(ps);
                auto one_pair = [&](int k1) {
                    const bool hi = k1 >= (N1 >> 1);
                    // stream rows of elements k1 (e = 0) and k1 + 1 (e = 1): low half (k1, k1 + 1), high half (N1 - 1 - k1, N1 - 2 - k1)
                    const unsigned jrow = hi ? (unsigned)(N1 - 2 - k1) >> 1 : (unsigned)k1 >> 1;
                    float2 ze, zo;
                    spectrum_normals(seed, (hi ? base_hi : base_lo) + (uint64_t)jrow * (uint64_t)uN2, &ze, &zo);
                    const unsigned kk0 = hi ? (unsigned)n - (uN2 * (unsigned)k1 + (unsigned)k2) : uN2 * (unsigned)k1 + (unsigned)k2;
                    const unsigned kk1 = hi ? kk0 - uN2 : kk0 + uN2;
                    float a0, a1;
                    if (amp_col) { const double2 a = *reinterpret_cast<const double2 *>(amp_col + k1); a0 = (float)a.x; a1 = (float)a.y; }
                    else {
                        const double wa = (double)kk0 * w_scale, wb = (double)kk1 * w_scale;
                        a0 = (float)ar.amp(ps, wa * wa); a1 = (float)ar.amp(ps, wb * wb);
                    }
                    // low half: element 0 sits on an even stream row (ze), element 1 on the odd one; high half: the other way round, conjugated
                    const float2 z0 = hi ? zo : ze, z1 = hi ? ze : zo;
                    const float sgn = hi ? -1.f : 1.f;
                    *reinterpret_cast<float4 *>(stage + k1) = make_float4(z0.x * a0, sgn * z0.y * a0, z1.x * a1, sgn * z1.y * a1);
                };
                for (int k1 = 2 * p; k1 < N1; k1 += 4 * TK_WS_HALF) { one_pair(k1); one_pair(k1 + 2 * TK_WS_HALF); }
            } else {
#pragma unroll 1
            for (int k1 = 2 * p; k1 < N1; k1 += 2 * TK_WS_HALF) {
                unsigned kk[2], row[2], col[2]; bool mir[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const unsigned r = (unsigned)(k1 + e), k = uN2 * r + (unsigned)k2;
                    mir[e] = k > half;
                    kk[e] = mir[e] ? (unsigned)n - k : k;
                    row[e] = !mir[e] ? r : (k2 == 0 ? (unsigned)N1 - r : (unsigned)N1 - 1u - r);
                    col[e] = (!mir[e] || k2 == 0) ? (unsigned)k2 : uN2 - (unsigned)k2;
                }
                const uint64_t blk0 = spectrum_block(lcg, M, g.N2, row[0], col[0]), blk1 = spectrum_block(lcg, M, g.N2, row[1], col[1]);
                float2 ze, zo, z[2];
                spectrum_normals(seed, blk0, &ze, &zo);
                z[0] = (row[0] & 1u) ? zo : ze;
                if (blk1 != blk0) spectrum_normals(seed, blk1, &ze, &zo);
                z[1] = (row[1] & 1u) ? zo : ze;
                float amp[2];
                if (amp_col) { const double2 a = *reinterpret_cast<const double2 *>(amp_col + k1); amp[0] = (float)a.x; amp[1] = (float)a.y; }
                else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) { const double wk = (double)kk[e] * w_scale; amp[e] = (float)tk_amp(ps, wk * wk); }
                }
                float4 X = make_float4(z[0].x * amp[0], z[0].y * amp[0], z[1].x * amp[1], z[1].y * amp[1]);
                if (kk[0] == half) X.y = 0.f;                    // Nyquist bin is real (simulator.py:499-500)
                if (kk[1] == half) X.w = 0.f;
                if (mir[0]) X.y = -X.y;
                if (mir[1]) X.w = -X.w;
                if (kk[0] == 0u) { X.x = 0.f; X.y = 0.f; }
                if (kk[1] == 0u) { X.z = 0.f; X.w = 0.f; }
                *reinterpret_cast<float4 *>(stage + k1) = X;
            }
            }
            __threadfence_block();
            tk_bar_arrive(1, 2 * TK_WS_HALF);                    // FULL
        }
    } else {
        // ================= consumers =================
        const int c = threadIdx.x;
        tk_bar_arrive(2, 2 * TK_WS_HALF);                        // the staging buffer starts free
        for (long long w = w0; w < w1; ++w) {
            const int l = (int)(w / g.K2), k2 = (int)(w % g.K2);
            if (l != l_loaded) {
                tk_bar_sync(3, TK_WS_HALF);                      // every consumer is done with the previous lightcurve's positions
                const int *ms = mslot + (size_t)l * g.Qmax;
                for (int q = c; q < g.Qmax; q += TK_WS_HALF) {
                    const int m = ms[q];
                    spos[q] = m >= 0 ? (unsigned short)tk_phys(tk_pos(g, m & (N1 - 1))) : (unsigned short)0xffff;
                }
                l_loaded = l;
            }
            tk_bar_sync(1, 2 * TK_WS_HALF);                      // FULL: the column is in the staging buffer
            const bool more = w + 1 < w1;
            tk_transform_and_keep<TK_WS_HALF, true>(g, data, stage, tw_g + g.tw_off[0], tw_s - tw_skip, spos,
                                                    Y + ((size_t)l * g.K2 + k2) * g.Qmax, c, [] { tk_bar_sync(3, TK_WS_HALF); },
                                                    [more] { if (more) tk_bar_arrive(2, 2 * TK_WS_HALF); });
        }
    }
}

// K13: grid (ceil(n_obs / TK_BINS_PER_CTA), Lc).  Dynamic shared memory: xs_max doubles.
__global__ void __launch_bounds__(256, 4)
k_tk95_rows(const __grid_constant__ TkGeom g, int n_obs, const int *__restrict__ q_off, const int *__restrict__ mslot,
            const int *__restrict__ cnt, const double2 *__restrict__ Y, double mean_add, double *__restrict__ out_tm,
            int L_total, int l_out0)
{
    extern __shared__ double tk_xs[];
    const int l = blockIdx.y, b0 = blockIdx.x * TK_BINS_PER_CTA, b1 = min(b0 + TK_BINS_PER_CTA, n_obs);
    const int q0 = q_off[b0], q1 = q_off[b1];
    const int *ms = mslot + (size_t)l * g.Qmax;
    const double2 *yl = Y + (size_t)l * g.K2 * g.Qmax;
    const bool n2_even = (g.N2 & 1) == 0;
    for (int q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
        const int m = ms[q];
        double acc = 0.0;
        if (m >= 0) {
            // twiddle w_n^(m k2), k2 = 0, 1, ...: geometric, ratio exp(2 pi i m / n)
            double sr, cr;
            sincospi(2.0 * ((double)m / (double)g.n), &sr, &cr);
            const double2 ratio = make_double2(cr, sr);
            double2 t;
            const double2 *yp = yl + q;
            const size_t ld = (size_t)g.Qmax;
            const int K2 = g.K2, last = n2_even ? K2 - 1 : K2;          // k2 = 0 (and N2/2 when N2 is even) count once, the rest twice
            {
                const double2 a = yp[0];
                acc = a.x;                                               // t = 1
                t = ratio;
            }
            int k2 = 1;
            for (; k2 + 8 <= last; k2 += 8) {
                double2 a[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] = yp[(size_t)(k2 + u) * ld];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    acc = fma(2.0, fma(t.x, a[u].x, -t.y * a[u].y), acc);
                    t = cmul(t, ratio);
                }
            }
            for (; k2 < last; ++k2) {
                const double2 a = yp[(size_t)k2 * ld];
                acc = fma(2.0, fma(t.x, a.x, -t.y * a.y), acc);
                t = cmul(t, ratio);
            }
            if (n2_even && K2 > 1) {
                const double2 a = yp[(size_t)(K2 - 1) * ld];
                acc += fma(t.x, a.x, -t.y * a.y);
            }
        }
        tk_xs[q - q0] = acc;
    }
    __syncthreads();
    const int b = b0 + threadIdx.x;
    if (threadIdx.x < TK_BINS_PER_CTA && b < b1) {
        const int c = cnt[(size_t)l * n_obs + b];
        const double *src = tk_xs + (q_off[b] - q0);
        double r;
        if (c <= 0) r = nan("");
        else if (mean_add == 0.0) r = np_mean(src, c);
        else {
            // countrate - meanrate + mean is applied per sample before averaging (simulator.py:393), as in k_bin
            double buf[32];
            if (c <= 32) { for (int i = 0; i < c; ++i) buf[i] = src[i] + mean_add; r = np_mean(buf, c); }
            else r = np_mean(src, c) + mean_add;
        }
        out_tm[(size_t)b * L_total + l_out0 + l] = r;
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
#define TK_SMEM_MAX (226u * 1024u)
static size_t tk_columns_smem(const TkGeom &g)
{
    return sizeof(double2) * ((size_t)g.N1 + (g.N1 >> 4) + 1 + (size_t)g.tw_total) + sizeof(unsigned short) * (((size_t)g.Qmax + 7) & ~(size_t)7);
}
static size_t tk_columns_ws_smem(const TkGeom &g)
{
    const int tw_skip = g.npass > 1 ? g.tw_off[1] : g.tw_total;
    return sizeof(double2) * ((size_t)g.N1 + (g.N1 >> 4) + 1 + (size_t)(g.tw_total - tw_skip)) + sizeof(float2) * (size_t)g.N1 +
           sizeof(unsigned short) * (((size_t)g.Qmax + 7) & ~(size_t)7);
}
static void tk_set_column_length(TkGeom *g, int n_fft, int p)
{
    g->n = n_fft; g->logN1 = p; g->N1 = 1 << p; g->N2 = n_fft >> p; g->K2 = g->N2 / 2 + 1; g->M = n_fft / 2 + 1;
    int L = g->N1, off = 0;
    g->npass = 0;
    while (L > 16) {
        const int lr = L >= 64 ? 3 : 2;
        g->lograd[g->npass] = lr; g->tw_off[g->npass] = off;
        L >>= lr;
        off += (lr == 3 ? 3 : 2) * L;
        ++g->npass;
    }
    g->F = L; g->tw_total = off;
}

void tk_geometry(int n_fft, int n_obs, const double *bin_lo, const double *bin_hi, double sim_dt, TkGeom *g, std::vector<int> *q_off)
{
    TkGeom z{};
    *g = z;
    int mode = 1;                       // MTG_SIM_FUSED: 0 never, 1 when it pays (default), 2 whenever the grid factorises (tests)
    if (const char *e = getenv("MTG_SIM_FUSED")) mode = atoi(e);
    if (n_fft < 64 || n_obs < 1) return;
    int p = 0;
    while (p < 13 && ((n_fft >> (p + 1)) << (p + 1)) == n_fft) ++p;
    if (const char *e = getenv("MTG_SIM_FUSED_LOG2N1")) { const int v = atoi(e); if (v >= 3 && v < p) p = v; }      // tests: force a column length
    if (p < 3) return;
    tk_set_column_length(g, n_fft, p);
    // sample slots: a bin of width w holds at most floor(w / dt) + 1 grid points; one more slot absorbs the rounding of the
    // grid times (k_need flags a bin that would need more)
    q_off->assign(n_obs + 1, 0);
    long long tot = 0;
    double expected = 0.0;
    for (int b = 0; b < n_obs; ++b) {
        const double wdt = (bin_hi[b] - bin_lo[b]) / sim_dt;
        if (!(wdt >= 0.0) || wdt > 1e6) return;
        tot += (long long)floor(wdt) + 2;
        expected += wdt + 0.5;
        if (tot > (1 << 24)) return;
        (*q_off)[b + 1] = (int)tot;
    }
    g->Qmax = (int)tot;
    int xs = 0;
    for (int b0 = 0; b0 < n_obs; b0 += TK_BINS_PER_CTA) {
        const int b1 = b0 + TK_BINS_PER_CTA < n_obs ? b0 + TK_BINS_PER_CTA : n_obs;
        xs = std::max(xs, (*q_off)[b1] - (*q_off)[b0]);
    }
    g->xs_max = xs;
    if ((size_t)xs * sizeof(double) > 160 * 1024) return;
    // shared memory of K12: column + twiddle tables + one 16-bit position per slot; shorten the columns until it fits
    while (tk_columns_smem(*g) > TK_SMEM_MAX && g->logN1 > 3) tk_set_column_length(g, n_fft, g->logN1 - 1);
    if (tk_columns_smem(*g) > TK_SMEM_MAX) return;
    // worth it when the entries kept per lightcurve (written once, read once) are fewer bytes than the spectrum + series
    // the cuFFT chain stores (and moves three to five times)
    const double fused_bytes = expected * g->K2 * 16.0, chain_bytes = (double)g->M * 16.0 + (double)n_fft * 8.0;
    if (mode != 2 && fused_bytes > chain_bytes) return;
    if (mode == 0) return;              // geometry kept: N2 is the pair stride of the spectrum stream on either path
    g->ok = 1;
}

void tk_twiddles(const TkGeom &g, std::vector<double2> *tw)
{
    tw->assign(g.tw_total > 0 ? g.tw_total : 1, make_double2(1.0, 0.0));
    int L = g.N1;
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int s = 0; s < g.npass; ++s) {
        const int lr = g.lograd[s], Ls = L >> lr, ntab = lr == 3 ? 3 : 2;
        for (int tab = 0; tab < ntab; ++tab) {
            const int mult = 1 << tab;                                        // w_L^(j), w_L^(2j), w_L^(4j)
            for (int j = 0; j < Ls; ++j) {
                const long double ang = two_pi * (long double)((long long)j * mult) / (long double)L;
                (*tw)[g.tw_off[s] + tab * Ls + j] = make_double2((double)cosl(ang), (double)sinl(ang));
            }
        }
        L = Ls;
    }
}

int tk_launch_amp_table(const TkGeom &g, double fold, const double *psd_table, double *amp_t, cudaStream_t stream)
{
    k_tk95_amp_table<<<dim3((g.N1 + 255) / 256, g.K2), 256, 0, stream>>>(g, fold, psd_table, amp_t);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k_tk95_amp_table launch");
}

int tk_launch_columns(const TkGeom &g, int Lc, double val, double fold, const PsdCoef *coefs, const double *psd_table,
                      const double *amp_t, const double *normals, uint64_t seed, uint64_t lc_offset, const double2 *tw,
                      const int *mslot, double2 *Y, cudaStream_t stream)
{
    const size_t dyn = tk_columns_smem(g);
    static bool opted[64][2] = {}, opted_ws[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int par = normals ? 1 : 0;
    if (dev < 64 && !opted[dev][par]) {
        cudaError_t e = par ? cudaFuncSetAttribute(k_tk95_columns<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TK_SMEM_MAX)
                            : cudaFuncSetAttribute(k_tk95_columns<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TK_SMEM_MAX);
        int rc = mtg_cuda_check(e, "k_tk95_columns shared-memory opt-in");
        if (rc) return rc;
        opted[dev][par] = true;
    }
    static int n_sm[64] = {};
    if (dev < 64 && n_sm[dev] == 0) cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    const int sms = dev < 64 && n_sm[dev] > 0 ? n_sm[dev] : 148;
    const long long work = (long long)Lc * g.K2;
    const int grid = (int)std::min<long long>(work, (long long)sms);
    static int ws_mode = -1;            // MTG_TK95_WS=0: the single-role kernel on the Philox path too (A/B)
    if (ws_mode < 0) { const char *e = getenv("MTG_TK95_WS"); ws_mode = e ? atoi(e) : 1; }
    if (par) k_tk95_columns<true><<<grid, TK_THREADS, dyn, stream>>>(g, Lc, val, fold, coefs, psd_table, nullptr, normals, seed, lc_offset, tw, mslot, Y);
    else if (ws_mode && g.npass >= 1 && (coefs || amp_t) && tk_columns_ws_smem(g) <= TK_SMEM_MAX) {
        if (dev < 64 && !opted_ws[dev]) {
            int rc = mtg_cuda_check(cudaFuncSetAttribute(k_tk95_columns_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TK_SMEM_MAX),
                                    "k_tk95_columns_ws shared-memory opt-in");
            if (rc) return rc;
            opted_ws[dev] = true;
        }
        k_tk95_columns_ws<<<grid, 2 * TK_WS_HALF, tk_columns_ws_smem(g), stream>>>(g, Lc, val, fold, coefs, amp_t, seed, lc_offset, tw, mslot, Y);
    }
    else     k_tk95_columns<false><<<grid, TK_THREADS, dyn, stream>>>(g, Lc, val, fold, coefs, psd_table, amp_t, normals, seed, lc_offset, tw, mslot, Y);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k_tk95_columns launch");
}

int tk_launch_rows(const TkGeom &g, int Lc, int n_obs, const int *q_off, const int *mslot, const int *cnt, const double2 *Y,
                   double mean_add, double *out_tm, int L_total, int l_out0, cudaStream_t stream)
{
    const size_t dyn = sizeof(double) * (size_t)std::max(g.xs_max, 1);
    if (dyn > 48 * 1024) {
        static bool opted[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 64 && !opted[dev]) {
            int rc = mtg_cuda_check(cudaFuncSetAttribute(k_tk95_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                                    "k_tk95_rows shared-memory opt-in");
            if (rc) return rc;
            opted[dev] = true;
        }
    }
    const dim3 grid((n_obs + TK_BINS_PER_CTA - 1) / TK_BINS_PER_CTA, Lc);
    k_tk95_rows<<<grid, 256, dyn, stream>>>(g, n_obs, q_off, mslot, cnt, Y, mean_add, out_tm, L_total, l_out0);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k_tk95_rows launch");
}
