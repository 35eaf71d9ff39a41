// mtg_scan_grad.cu -- K11: batched log-likelihood + gradient, one CTA per series, parallel in time (mtg_scan_grad.cuh).
//
// The low-latency twin of K2 (mtg_grad.cu) for small batches: a gradient of <= 148 series at N = 1000 costs one scan and one
// suffix scan (tens of microseconds) instead of three dependent sweeps of one thread (1.1 ms).  mtg_loglike_grad_dev routes here
// when the batch is small enough that a few waves of CTAs beat K2's fixed latency (MTG_SCAN_GRAD_MAX_B), the model has
// <= 4 state dimensions and the series + tape fit the shared memory of an SM; MTG_FLAG_SCAN / MTG_FLAG_NO_SCAN force either.
#include "mtg_scan_grad.cuh"
#include <cstdlib>

namespace k9 {

template <int JR, int JC, bool DUAL>
__global__ void __launch_bounds__(K9_THREADS, 1)
k11_loglike_grad_scan(const __grid_constant__ Plan pl, const BatchArgs a, const K11Layout lay, double *__restrict__ scratch)
{
    constexpr int J = JR + 2 * JC;
    __shared__ Scratch<J> sc;
    __shared__ GScratch<JR, JC, DUAL> gs;
    __shared__ double xs[MTG_MAX_PARAMS], gsm[MTG_MAX_PARAMS];
    extern __shared__ double k11_dyn[];               // [3][N] series when staged, then the tape when it fits
    double *series = k11_dyn;
    double *tape = lay.tape_smem ? k11_dyn + (lay.stage ? 3 * (size_t)a.N : 0) : scratch + (size_t)blockIdx.x * lay.scratch_doubles;
    const int nd = pl.n_free, tid = threadIdx.x;
    const int nwork = a.count ? min(a.B, *a.count) : a.B;
    const bool dsm = (a.flags & MTG_FLAG_DS_MEAN) != 0;
    int staged_ds = -1;
    for (int w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int b = a.index ? a.index[w] : w;
        const int ds = b / a.group;
        __syncthreads();                              // previous series done with xs / gsm / the staged data
        if (tid < nd) xs[tid] = a.params[(size_t)b * nd + tid];
        const double *st_t = a.t, *st_y = a.y + ds, *st_d = a.diag + ds;
        size_t stride = (size_t)a.S;
        if (lay.stage) {
            if (ds != staged_ds) { stage_series(series, a.N, a.t, a.y + ds, a.diag + ds, (size_t)a.S); staged_ds = ds; }
            else __syncthreads();
            st_t = series; st_y = series + a.N; st_d = series + 2 * a.N; stride = 1;
        } else __syncthreads();
        const bool inside = !(a.flags & MTG_FLAG_APPLY_PRIOR) || plan_in_prior(pl, xs);
        int st;
        double ll = series_loglike_grad<JR, JC, DUAL>(pl, xs, a.N, st_t, st_y, st_d, stride, dsm, dsm ? a.ds_mean[ds] : 0.0,
                                                      sc, gs, tape, &st, gsm);
        if (!inside) { ll = -INFINITY; st = 3; }
        if (tid == 0) { a.out_ll[b] = ll; if (a.status) a.status[b] = st; }
        if (tid < nd) a.out_grad[(size_t)b * nd + tid] = inside ? gsm[tid] : 0.0;
    }
}

}  // namespace k9

// every model the scan covers (<= 4 state dimensions), any series length: the series is staged in shared memory up to
// K9_STAGE_MAX_N points and the tape lives in shared memory while it fits, in a per-CTA global scratch beyond (mtg_k11_layout)
bool mtg_k11_supports(const Plan &pl, int N)
{
    return mtg_k9_supports(pl) && N >= 1;
}

static int k11_sms()
{
    static int sms[64]; static bool have[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!have[dev]) { cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev); have[dev] = true; }
    return sms[dev];
}

template <int JR, int JC, bool DUAL>
static int launch_k11_one(const Plan &pl, const BatchArgs &a, cudaStream_t stream)
{
    const K11Layout lay = mtg_k11_layout(JR + 2 * JC, a.N);
    auto kern = k9::k11_loglike_grad_scan<JR, JC, DUAL>;
    MTG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, K11_SMEM_LIMIT));
    const int grid = a.B < k11_sms() ? a.B : k11_sms();          // one CTA per SM (255 registers x 256 threads), grid-stride over the series
    double *scratch = nullptr;
    if (lay.scratch_doubles) {
        int rc = mtg_pool_setup();
        if (rc) return rc;
        MTG_CUDA(cudaMallocAsync((void **)&scratch, sizeof(double) * lay.scratch_doubles * (size_t)grid, stream));
    }
    kern<<<grid, K9_THREADS, lay.dyn_bytes, stream>>>(pl, a, lay, scratch);
    mtg_count_launch();
    int rc = mtg_cuda_check(cudaGetLastError(), "k11_loglike_grad_scan launch");
    if (scratch) cudaFreeAsync(scratch, stream);
    return rc;
}
template <int JR, int JC>
static int launch_k11(const Plan &pl, const BatchArgs &a, cudaStream_t stream)
{
    if (pl.has_sho) return launch_k11_one<JR, JC, true>(pl, a, stream);
    return launch_k11_one<JR, JC, false>(pl, a, stream);
}
#define K11_CASE(R, C) if (pl.jr == R && pl.jc == C) return launch_k11<R, C>(pl, a, stream);

int mtg_k11_dispatch(const Plan &pl, const BatchArgs &a, cudaStream_t stream)
{
    K11_CASE(1, 0) K11_CASE(2, 0) K11_CASE(3, 0) K11_CASE(0, 1) K11_CASE(1, 1)
    K11_CASE(2, 1) K11_CASE(0, 2)
    mtg_set_error("scan gradient: unsupported kernel structure: %d real + %d complex slots", pl.jr, pl.jc);
    return MTG_EUNSUP;
}

// largest batch the automatic dispatch hands to K11: ~32 waves of one CTA per SM cost what one K2 launch does
long long mtg_scan_grad_max_batch()
{
    static long long v = -1;
    if (v < 0) { const char *e = getenv("MTG_SCAN_GRAD_MAX_B"); v = e ? atoll(e) : 4096; }
    return v;
}

extern "C" int mtg_scan_grad_supported(const mtg_model_t *model, int N)
{
    Plan pl;
    if (!model || mtg_build_plan(model, &pl)) return 0;
    return mtg_k11_supports(pl, N) ? 1 : 0;
}
