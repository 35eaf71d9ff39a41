// mtg_loglike.cu -- K1: batched celerite log-likelihood, one thread per series (sm_100a).
//
// Replaces celerite CholeskySolver.compute + dot_solve + log_determinant (un-vendored C++), as
// reached from mind_the_gaps/gpmodelling.py:54, :147-152 (_log_probability), :168-169
// (_neg_log_like).  The factor, forward-solve and log-determinant passes of celerite are fused
// into ONE sweep over the N observations: only (S, f, W, D, z) of the previous step are live, so
// the whole state (J(J+1)/2 + 2J + O(1) doubles) stays in registers and nothing O(N) is stored.
//
// Memory plan: the time array (and, when one lightcurve is shared by the batch, rates and
// variances too) is staged through shared memory in 256-step tiles by TMA bulk copies
// (cp.async.bulk + mbarrier, double buffered); every thread then reads the same smem word
// (broadcast, conflict free).  Per-data-set rates/variances are read straight from the
// time-major [N][S] arrays: a warp touches <= 32 consecutive doubles per step (coalesced) and the
// next step's values are prefetched one step ahead.
#include "mtg_recursion.cuh"
#include <algorithm>
#include <cstdlib>

// CTA shape per recursion width J = JR + 2 JC.  Every CTA carries 48 KB of lookup tables + <= 12 KB of tiles, so at
// most 3 CTAs fit the 227 KB of an SM: the narrow recursions (J <= 3, <= 128 registers) run 256 threads x 2 CTAs
// (16 warps per SM, as 4 x 128 before the tables grew), the wide ones keep 128 threads and their natural
// register count (J = 7 needs 255).
#ifndef K1_NARROW_J
#define K1_NARROW_J 3
#endif
#ifndef K1_NARROW_THREADS
#define K1_NARROW_THREADS 256
#define K1_NARROW_BLOCKS 2
#endif
#ifndef K1_WIDE_THREADS
#define K1_WIDE_THREADS 128
#define K1_WIDE_BLOCKS 1
#endif
constexpr int k1_threads(int J) { return J <= K1_NARROW_J ? K1_NARROW_THREADS : K1_WIDE_THREADS; }
constexpr int k1_min_blocks(int J) { return J <= K1_NARROW_J ? K1_NARROW_BLOCKS : K1_WIDE_BLOCKS; }
#ifndef K1_UNROLL
#define K1_UNROLL 2
#endif
#ifndef K1_WIDE_UNROLL
#define K1_WIDE_UNROLL K1_UNROLL
#endif
constexpr int k1_unroll(int J) { return J <= K1_NARROW_J ? K1_UNROLL : K1_WIDE_UNROLL; }
// (A hand-rotated software pipeline -- phi / u / v of step i + 1 computed beside the state update of step i through
//  K1State::prepare / advance -- was measured and dropped: 1.36e8 vs 1.57e8 evaluations/s; the extra live registers
//  cost more than the filled dependency bubbles gain at the 128-register cap.)
#define K1_STR2(x) #x
#define K1_STR(x) K1_STR2(x)
#define K1_TILE 256

// SHARED: S == 1, rates/variances staged through smem with the times.
template <int JR, int JC, bool DUAL, bool SHARED>
__global__ void __launch_bounds__(k1_threads(JR + 2 * JC), k1_min_blocks(JR + 2 * JC))
k1_loglike(const __grid_constant__ Plan pl, const BatchArgs a)
{
    constexpr int K1_THREADS = k1_threads(JR + 2 * JC);
    constexpr int UNR = k1_unroll(JR + 2 * JC);
    constexpr int NARR = SHARED ? 3 : 1;
    __shared__ alignas(128) double sm[2][NARR][K1_TILE];
    __shared__ alignas(8) uint64_t bar[2];
    extern __shared__ __align__(128) double2 k1_dyn_smem[];      // MTG_TABSB_DOUBLES doubles: the big lookup tables
    __shared__ alignas(8) uint64_t bar_tab;
    double *tabs_sm = reinterpret_cast<double *>(k1_dyn_smem);
    mtgm::TabsBig tb; tb.e2 = tabs_sm; tb.trig = tabs_sm + MTG_EXP2B_TAB_N;

    const int tid = threadIdx.x;
    const int b_raw = blockIdx.x * K1_THREADS + tid;
    const bool valid = b_raw < a.B;
    const int b = valid ? b_raw : a.B - 1;
    const int N = a.N;
    const int ntiles = (N + K1_TILE - 1) / K1_TILE;
    const bool aligned = ((((uintptr_t)a.t) & 15) == 0) &&
                         (!SHARED || (((((uintptr_t)a.y) | ((uintptr_t)a.diag)) & 15) == 0));
    auto tile_count = [&](int k) { return min(K1_TILE, N - k * K1_TILE); };
    auto use_bulk = [&](int k) { return aligned && ((tile_count(k) & 1) == 0); };
    auto issue = [&](int k) {      // one thread
        const int st = k & 1, cnt = tile_count(k);
        const uint32_t bytes = (uint32_t)cnt * 8u;
        mbar_expect_tx(&bar[st], bytes * NARR);
        bulk_g2s(&sm[st][0][0], a.t + (size_t)k * K1_TILE, bytes, &bar[st]);
        if (SHARED) {
            bulk_g2s(&sm[st][1][0], a.y + (size_t)k * K1_TILE, bytes, &bar[st]);
            bulk_g2s(&sm[st][2][0], a.diag + (size_t)k * K1_TILE, bytes, &bar[st]);
        }
    };
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar_tab, 1); mbar_fence_init(); }
    __syncthreads();
    if (tid == 0) {
        // the 48 KB of lookup tables arrive by two TMA bulk copies that overlap the per-series set-up below (a
        // thread-strided fill cost a latency-bound 32-series call 70 us)
        mbar_expect_tx(&bar_tab, (uint32_t)(sizeof(double) * MTG_TABSB_DOUBLES));
        bulk_g2s(tabs_sm, mtgm::EXP2B_TAB, (uint32_t)(sizeof(double) * MTG_EXP2B_TAB_N), &bar_tab);
        bulk_g2s(tabs_sm + MTG_EXP2B_TAB_N, mtgm::TRIGB_TAB, (uint32_t)(sizeof(double) * 2 * MTG_TRIGB_TAB_N), &bar_tab);
        if (use_bulk(0)) issue(0);
        if (ntiles > 1 && use_bulk(1)) issue(1);
    }

    // per-series set-up (overlaps the first bulk copies)
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

    K1State<JR, JC, DUAL> s;
    s.init(__ldg(a.t));
    double y_nx = 0.0, d_nx = 0.0;
    if (!SHARED) { y_nx = __ldg(yp); d_nx = __ldg(dp); }

    mbar_wait(&bar_tab, 0u);                   // tables resident
    uint32_t phase[2] = {0u, 0u};
    for (int k = 0; k < ntiles; ++k) {
        const int st = k & 1, cnt = tile_count(k), n0 = k * K1_TILE;
        if (use_bulk(k)) { mbar_wait(&bar[st], phase[st]); phase[st] ^= 1u; }
        else {
            for (int i = tid; i < cnt; i += K1_THREADS) {
                sm[st][0][i] = a.t[n0 + i];
                if (SHARED) { sm[st][1][i] = a.y[n0 + i]; sm[st][2][i] = a.diag[n0 + i]; }
            }
            __syncthreads();
        }
        const double *st_t = &sm[st][0][0];
#pragma unroll UNR
        for (int i = 0; i < cnt; ++i) {
            const int n = n0 + i;
            const double tn = st_t[i];
            double yn, dn;
            if (SHARED) { yn = sm[st][1][i]; dn = sm[st][2][i]; }
            else {
                yn = y_nx; dn = d_nx;
                if (n + 1 < N) { y_nx = __ldg(yp + (size_t)(n + 1) * S); d_nx = __ldg(dp + (size_t)(n + 1) * S); }
            }
            s.step(c, tb, tn, yn - mean(tn), dn, n);
        }
        __syncthreads();                       // every thread is done with stage st
        if (tid == 0 && k + 2 < ntiles && use_bulk(k + 2)) { fence_proxy_async(); issue(k + 2); }
    }

    int stt;
    double ll = s.loglike(N, &stt);
    if (!inside) { ll = -INFINITY; stt = 3; }
    if (valid) {
        a.out_ll[b] = ll;
        if (a.status) a.status[b] = stt;
    }
}

// ---- dispatch ---------------------------------------------------------------------------------
template <int JR, int JC, bool DUAL, bool SHARED>
static int launch_k1_one(const Plan &pl, const BatchArgs &a, cudaStream_t stream)
{
    constexpr size_t dyn = sizeof(double) * MTG_TABSB_DOUBLES;       // 48 KB: above the default limit, opt in once per device
    static bool opted[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !opted[dev]) {
        int rc = mtg_cuda_check(cudaFuncSetAttribute(k1_loglike<JR, JC, DUAL, SHARED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn),
                                "k1_loglike shared-memory opt-in");
        if (rc) return rc;
        opted[dev] = true;
    }
    constexpr int K1_THREADS = k1_threads(JR + 2 * JC);
    const dim3 grid((a.B + K1_THREADS - 1) / K1_THREADS), block(K1_THREADS);
    k1_loglike<JR, JC, DUAL, SHARED><<<grid, block, dyn, stream>>>(pl, a);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "k1_loglike launch");
}

template <int JR, int JC>
static int launch_k1(const Plan &pl, const BatchArgs &a, cudaStream_t stream)
{
    const bool shared = (a.S == 1);
    int rc;
    if (pl.has_sho) rc = shared ? launch_k1_one<JR, JC, true, true>(pl, a, stream) : launch_k1_one<JR, JC, true, false>(pl, a, stream);
    else            rc = shared ? launch_k1_one<JR, JC, false, true>(pl, a, stream) : launch_k1_one<JR, JC, false, false>(pl, a, stream);
    return rc;
}

#define K1_CASE(R, C) if (pl.jr == R && pl.jc == C) return launch_k1<R, C>(pl, a, stream);

int mtg_k1_dispatch(const Plan &pl, const BatchArgs &a, cudaStream_t stream)
{
    K1_CASE(1, 0) K1_CASE(2, 0) K1_CASE(3, 0)
    K1_CASE(0, 1) K1_CASE(1, 1) K1_CASE(2, 1) K1_CASE(3, 1)
    K1_CASE(0, 2) K1_CASE(1, 2) K1_CASE(2, 2)
    K1_CASE(0, 3) K1_CASE(1, 3) K1_CASE(2, 3)
    K1_CASE(0, 4) K1_CASE(1, 4)
    mtg_set_error("unsupported kernel structure: %d real + %d complex recursion slots", pl.jr, pl.jc);
    return MTG_EUNSUP;
}

// largest batch the automatic dispatch hands to the scan kernel (MTG_SCAN_MAX_B overrides)
static long long mtg_scan_max_batch()
{
    static long long v = -1;
    if (v < 0) { const char *e = getenv("MTG_SCAN_MAX_B"); v = e ? atoll(e) : 4096; }
    return v;
}

extern "C" int mtg_loglike_dev(const mtg_model_t *model, int N, int S, int group, int B,
                               const double *t, const double *y, const double *diag, const double *ds_mean,
                               const double *params, double *out_ll, int32_t *status,
                               uint32_t flags, void *stream)
{
    if (!model || N < 1 || S < 1 || group < 1 || B < 0 || !t || !y || !diag || !params || !out_ll) {
        mtg_set_error("mtg_loglike_dev: bad argument"); return MTG_EINVAL;
    }
    if ((flags & MTG_FLAG_DS_MEAN) && !ds_mean) { mtg_set_error("MTG_FLAG_DS_MEAN without ds_mean"); return MTG_EINVAL; }
    if ((long long)(B - 1) / group >= S && B > 0) { mtg_set_error("B/group exceeds the number of data sets S"); return MTG_EINVAL; }
    if (B == 0) return 0;
    Plan pl;
    int rc = mtg_build_plan(model, &pl);
    if (rc) return rc;
    BatchArgs a{N, S, group, B, t, y, diag, ds_mean, params, out_ll, nullptr, status, flags, nullptr, nullptr, nullptr};
    // Small batches of narrow kernels: one CTA per series, the N-step chain parallelised by a scan (K9, mtg_scan.cu).
    // K1 needs >= one thread per series to pay off: its latency is that of the whole chain (0.2 ms at N = 1000) whatever
    // the batch, K9's cost is ~12x the work spread over 256 threads.  Crossover measured at ~2000 series (N = 1000).
    const bool can_scan = mtg_k9_supports(pl);
    if ((flags & MTG_FLAG_SCAN) && !can_scan) { mtg_set_error("MTG_FLAG_SCAN: model not covered by the scan kernel"); return MTG_EUNSUP; }
    // measured crossover (profiles/README.md, tools/k9_probe.py): K9 wins up to ~2.9 N series (N = 1000) / ~3.7 N (N = 200);
    // one wave of CTAs (148) is always a win for N >= 64
    const long long scan_b = std::max(148LL, std::min(2LL * N, mtg_scan_max_batch()));
    const bool scan = (flags & MTG_FLAG_SCAN) || (can_scan && !(flags & MTG_FLAG_NO_SCAN) && N >= 64 && (long long)B <= scan_b);
    if (scan) return mtg_k9_dispatch(pl, a, (cudaStream_t)stream);
    return mtg_k1_dispatch(pl, a, (cudaStream_t)stream);
}
