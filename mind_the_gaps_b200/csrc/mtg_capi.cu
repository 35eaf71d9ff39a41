// mtg_capi.cu -- C-ABI glue: model compilation, error reporting, host-buffer entry points,
// measurement probes.  See include/mtg_b200.h for the contract of every symbol.
#include "mtg_common.cuh"
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void mtg_set_error(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
int mtg_cuda_check(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return 0;
    mtg_set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
void mtg_count_launch(uint64_t n) { g_launches += n; }

extern "C" int mtg_version(void) { return MTG_VERSION; }
extern "C" const char *mtg_last_error(void) { return g_err; }
extern "C" uint64_t mtg_launch_count(void) { return g_launches.load(); }
extern "C" int mtg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int mtg_set_device(int device) { return mtg_cuda_check(cudaSetDevice(device), "cudaSetDevice"); }

// Stream-ordered scratch (K2 tape, L-BFGS state) comes from the device's default memory pool; keep
// freed blocks cached in the pool instead of returning them to the driver at every synchronisation,
// otherwise each call pays a multi-GB cudaMalloc.
int mtg_pool_setup()
{
    static std::mutex mu; static bool done[64] = {false};
    int dev = 0;
    MTG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev >= 0 && dev < 64 && done[dev]) return 0;
    cudaMemPool_t pool;
    MTG_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t keep = UINT64_MAX;
    MTG_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    if (dev >= 0 && dev < 64) done[dev] = true;
    return 0;
}

// ---- model -> plan ------------------------------------------------------------------------------
int mtg_build_plan(const mtg_model_t *m, Plan *pl)
{
    memset(pl, 0, sizeof *pl);
    if (m->n_terms < 1 || m->n_terms > MTG_MAX_TERMS || m->n_full < 1 || m->n_full > MTG_MAX_PARAMS ||
        m->n_free < 0 || m->n_free > m->n_full) {
        mtg_set_error("mtg_model_t: n_terms=%d n_full=%d n_free=%d out of range", m->n_terms, m->n_full, m->n_free);
        return MTG_EINVAL;
    }
    auto chk = [&](int i) { return i >= 0 && i < m->n_full; };
    int jr = 0, jc = 0, jcel = 0;
    for (int k = 0; k < m->n_terms; ++k) {
        const int kind = m->kind[k];
        const int *ix = m->pidx[k];
        int np = 0;
        switch (kind) {
        case MTG_TERM_REAL: case MTG_TERM_DRW: case MTG_TERM_RAW_REAL: np = 2;
            if (jr >= MTG_MAXR) goto toowide;
            pl->r_ia[jr] = ix[0]; pl->r_ic[jr] = ix[1]; pl->r_raw[jr] = kind == MTG_TERM_RAW_REAL; ++jr; jcel += 1; break;
        case MTG_TERM_JITTER: case MTG_TERM_RAW_JITTER: np = 1;
            pl->jit_raw[pl->n_jit] = kind == MTG_TERM_RAW_JITTER; pl->jit_idx[pl->n_jit++] = ix[0]; break;
        case MTG_TERM_RAW_COMPLEX:
            if (jc >= MTG_MAXC) goto toowide;
            np = 4; pl->c_kind[jc] = CS_RAW; pl->c_extra[jc] = 0.0;
            for (int i = 0; i < 4; ++i) pl->c_idx[jc][i] = ix[i];
            ++jc; jcel += 2; break;
        case MTG_TERM_COMPLEX3: case MTG_TERM_COMPLEX4: case MTG_TERM_SHO: case MTG_TERM_MATERN32:
        case MTG_TERM_LORENTZIAN: case MTG_TERM_COSINUS: case MTG_TERM_BENDINGPL: {
            if (jc >= MTG_MAXC) goto toowide;
            int ck = CS_COMPLEX3; np = 3;
            if (kind == MTG_TERM_COMPLEX4) { ck = CS_COMPLEX4; np = 4; }
            else if (kind == MTG_TERM_SHO) { ck = CS_SHO; pl->has_sho = 1; }
            else if (kind == MTG_TERM_MATERN32) { ck = CS_MATERN32; np = 2; }
            else if (kind == MTG_TERM_LORENTZIAN) { ck = CS_LORENTZIAN; jcel += 1; /* degenerate real (0,0): celerite_models.py:9-15 */ }
            else if (kind == MTG_TERM_COSINUS) { ck = CS_COSINUS; np = 2; }
            else if (kind == MTG_TERM_BENDINGPL) {
                ck = CS_BENDINGPL; pl->bpl_s0[pl->n_bpl] = ix[0]; pl->bpl_q[pl->n_bpl] = ix[1]; ++pl->n_bpl;
            }
            pl->c_kind[jc] = ck; pl->c_extra[jc] = m->extra[k];
            for (int i = 0; i < 4; ++i) pl->c_idx[jc][i] = i < np ? ix[i] : 0;
            if (ck == CS_MATERN32 && !(m->extra[k] > 0.0)) { mtg_set_error("Matern32 eps must be > 0"); return MTG_EINVAL; }
            ++jc; jcel += 2; } break;
        default: mtg_set_error("unknown term kind %d", kind); return MTG_EINVAL;
        }
        for (int i = 0; i < np; ++i)
            if (!chk(ix[i])) { mtg_set_error("term %d parameter index %d out of range", k, ix[i]); return MTG_EINVAL; }
    }
    pl->jr = jr; pl->jc = jc; pl->j_celerite = jcel;
    pl->b_zero = 1;
    for (int j = 0; j < jc; ++j)
        if (pl->c_kind[j] != CS_COMPLEX3 && pl->c_kind[j] != CS_LORENTZIAN && pl->c_kind[j] != CS_COSINUS) pl->b_zero = 0;
    if (jr + 2 * jc > MTG_MAX_J) goto toowide;
    pl->mean_kind = m->mean_kind;
    {
        const int nm = m->mean_kind == MTG_MEAN_CONSTANT ? 1 : (m->mean_kind == MTG_MEAN_LINEAR ? 2 : 4);
        if (m->mean_kind < 0 || m->mean_kind > MTG_MEAN_GAUSSIAN) { mtg_set_error("unknown mean kind"); return MTG_EINVAL; }
        for (int i = 0; i < 4; ++i) {
            pl->mean_idx[i] = i < nm ? m->mean_pidx[i] : 0;
            if (i < nm && !chk(m->mean_pidx[i])) { mtg_set_error("mean parameter index out of range"); return MTG_EINVAL; }
        }
    }
    pl->n_full = m->n_full; pl->n_free = m->n_free;
    for (int i = 0; i < m->n_full; ++i) { pl->src_pos[i] = -1; pl->full_default[i] = m->full_default[i]; }
    for (int k = 0; k < m->n_free; ++k) {
        if (!chk(m->free_idx[k])) { mtg_set_error("free_idx out of range"); return MTG_EINVAL; }
        pl->free_idx[k] = m->free_idx[k]; pl->src_pos[m->free_idx[k]] = k;
        pl->lo[k] = m->lo[k]; pl->hi[k] = m->hi[k];
    }
    return 0;
toowide:
    mtg_set_error("kernel too wide for this build (max %d real + %d complex slots, J <= %d)", MTG_MAXR, MTG_MAXC, MTG_MAX_J);
    return MTG_EUNSUP;
}

extern "C" int mtg_model_width(const mtg_model_t *model, int *j_celerite, int *j_compute)
{
    Plan pl; int rc = mtg_build_plan(model, &pl);
    if (rc) return rc;
    if (j_celerite) *j_celerite = pl.j_celerite;
    if (j_compute) *j_compute = pl.jr + 2 * pl.jc;
    return 0;
}

// ---- host-buffer context ----------------------------------------------------------------------
// One lazily created context per (process, DEVICE): an internal stream, a small chunk pipeline and grow-only device
// buffers, all created on that device, so repeated host calls (an emcee loop) do not pay cudaMalloc each time and one
// process can drive several GPUs (mtg_set_device(d) then mtg_*_host).  The mutex is per context: host calls aimed
// at different devices run concurrently, calls aimed at the same device share its buffers and serialise.
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { mtg_set_error("cudaMalloc(%zu): %s", want, cudaGetErrorString(e)); return MTG_ENOMEM; }
        cap = want; return 0;
    }
};
struct HostCtx {
    std::mutex mu; cudaStream_t stream = nullptr; int device = -1;
    static constexpr int NPIPE = 3;
    cudaStream_t pipe[NPIPE] = {nullptr, nullptr, nullptr};      // chunk pipeline of large host calls
    cudaEvent_t data_ready = nullptr;
    DevBuf t, y, d, ysn, dsn, mean, par, ll, grad, st, x, nll, it, scratch;
    int ensure(int dev)          // called with mu held and `dev` current
    {
        if (stream) return 0;
        MTG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (int i = 0; i < NPIPE; ++i) MTG_CUDA(cudaStreamCreateWithFlags(&pipe[i], cudaStreamNonBlocking));
        MTG_CUDA(cudaEventCreateWithFlags(&data_ready, cudaEventDisableTiming));
        device = dev; return 0;
    }
};
#define MTG_MAX_DEVICES 64
static HostCtx g_ctx[MTG_MAX_DEVICES];
// the context of the calling thread's current device (NULL + error text when there is none)
static HostCtx *host_ctx()
{
    int dev = -1;
    if (mtg_cuda_check(cudaGetDevice(&dev), "cudaGetDevice")) return nullptr;
    if (dev < 0 || dev >= MTG_MAX_DEVICES) { mtg_set_error("device index %d out of range", dev); return nullptr; }
    return &g_ctx[dev];
}

__global__ void k_transpose_sn_to_ns(const double *__restrict__ in, double *__restrict__ out, int S, int N)
{
    __shared__ double tile[32][33];
    const int n0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int s = s0 + i, n = n0 + threadIdx.x;
        if (s < S && n < N) tile[i][threadIdx.x] = in[(size_t)s * N + n];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int n = n0 + i, s = s0 + threadIdx.x;
        if (s < S && n < N) out[(size_t)n * S + s] = tile[threadIdx.x][i];
    }
}

int mtg_transpose_sn_to_ns(const double *in_sn, double *out_ns, int S, int N, cudaStream_t stream)
{
    if (S == 1) return mtg_cuda_check(cudaMemcpyAsync(out_ns, in_sn, sizeof(double) * N, cudaMemcpyDeviceToDevice, stream), "copy");
    dim3 grid((N + 31) / 32, (S + 31) / 32), block(32, 8);
    k_transpose_sn_to_ns<<<grid, block, 0, stream>>>(in_sn, out_ns, S, N);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "transpose launch");
}

// upload the data arrays of a host call; leaves t / y / diag / ds_mean device pointers in ctx
static int upload_data(HostCtx &c, int N, int S, const double *t, const double *y_sn, const double *diag_sn,
                       const double *ds_mean, uint32_t flags)
{
    const size_t nb = sizeof(double) * (size_t)N, snb = nb * (size_t)S;
    int rc;
    if ((rc = c.t.reserve(nb + 16)) || (rc = c.y.reserve(snb + 16)) || (rc = c.d.reserve(snb + 16))) return rc;
    MTG_CUDA(cudaMemcpyAsync(c.t.p, t, nb, cudaMemcpyHostToDevice, c.stream));
    if (S == 1) {
        MTG_CUDA(cudaMemcpyAsync(c.y.p, y_sn, snb, cudaMemcpyHostToDevice, c.stream));
        MTG_CUDA(cudaMemcpyAsync(c.d.p, diag_sn, snb, cudaMemcpyHostToDevice, c.stream));
    } else {
        if ((rc = c.ysn.reserve(snb)) || (rc = c.dsn.reserve(snb))) return rc;
        MTG_CUDA(cudaMemcpyAsync(c.ysn.p, y_sn, snb, cudaMemcpyHostToDevice, c.stream));
        MTG_CUDA(cudaMemcpyAsync(c.dsn.p, diag_sn, snb, cudaMemcpyHostToDevice, c.stream));
        if ((rc = mtg_transpose_sn_to_ns((const double *)c.ysn.p, (double *)c.y.p, S, N, c.stream))) return rc;
        if ((rc = mtg_transpose_sn_to_ns((const double *)c.dsn.p, (double *)c.d.p, S, N, c.stream))) return rc;
    }
    if (flags & MTG_FLAG_DS_MEAN) {
        if (!ds_mean) { mtg_set_error("MTG_FLAG_DS_MEAN without ds_mean"); return MTG_EINVAL; }
        if ((rc = c.mean.reserve(sizeof(double) * S))) return rc;
        MTG_CUDA(cudaMemcpyAsync(c.mean.p, ds_mean, sizeof(double) * S, cudaMemcpyHostToDevice, c.stream));
    }
    return 0;
}

extern "C" int mtg_loglike_grad_dev(const mtg_model_t *, int, int, int, int, const double *, const double *,
                                    const double *, const double *, const double *, double *, double *,
                                    int32_t *, uint32_t, void *);
extern "C" int mtg_lbfgs_dev(const mtg_model_t *, int, int, int, int, const double *, const double *,
                             const double *, const double *, const double *, const mtg_lbfgs_opts_t *,
                             double *, double *, int32_t *, int32_t *, uint32_t, void *);

static int host_eval(const mtg_model_t *model, int N, int S, int group, int B,
                     const double *t, const double *y_sn, const double *diag_sn, const double *ds_mean,
                     const double *params, double *out_ll, double *out_grad, int32_t *status, uint32_t flags)
{
    if (!model || N < 1 || S < 1 || group < 1 || B < 0 || !t || !y_sn || !diag_sn || !params || !out_ll) {
        mtg_set_error("mtg_loglike_host: bad argument"); return MTG_EINVAL;
    }
    if (B == 0) return 0;
    HostCtx *cp = host_ctx();
    if (!cp) return MTG_ENODEV;
    HostCtx &c = *cp;
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.ensure((int)(cp - g_ctx));
    if (rc) return rc;
    const int nd = model->n_free;
    if ((rc = upload_data(c, N, S, t, y_sn, diag_sn, ds_mean, flags))) return rc;
    if ((rc = c.par.reserve(sizeof(double) * (size_t)B * (nd > 0 ? nd : 1))) || (rc = c.ll.reserve(sizeof(double) * B)) ||
        (rc = c.st.reserve(sizeof(int32_t) * B))) return rc;
    if (out_grad && (rc = c.grad.reserve(sizeof(double) * (size_t)B * (nd > 0 ? nd : 1)))) return rc;
    // One chunk = H2D of its parameter rows, the kernel, D2H of its results, all on one stream.  Large batches are cut
    // into chunks of whole data-set groups (about two full waves of K1) that rotate over NPIPE streams, so the copy
    // engines run under the kernels of the neighbouring chunks instead of in front of / behind one big launch.
    auto run_chunk = [&](size_t b0, int Bc, cudaStream_t s) -> int {
        const size_t ds0 = b0 / (size_t)group;
        const double *yd = (const double *)c.y.p + ds0, *dd = (const double *)c.d.p + ds0;
        const double *md = (flags & MTG_FLAG_DS_MEAN) ? (const double *)c.mean.p + ds0 : (const double *)c.mean.p;
        double *pd = (double *)c.par.p + b0 * (size_t)(nd > 0 ? nd : 1), *ld = (double *)c.ll.p + b0;
        int32_t *sd = (int32_t *)c.st.p + b0;
        if (nd > 0) MTG_CUDA(cudaMemcpyAsync(pd, params + b0 * (size_t)nd, sizeof(double) * (size_t)Bc * nd, cudaMemcpyHostToDevice, s));
        int r;
        if (out_grad) {
            double *gd = (double *)c.grad.p + b0 * (size_t)(nd > 0 ? nd : 1);
            r = mtg_loglike_grad_dev(model, N, S, group, Bc, (const double *)c.t.p, yd, dd, md, pd, ld, gd, sd, flags, s);
            if (r) return r;
            if (nd > 0) MTG_CUDA(cudaMemcpyAsync(out_grad + b0 * (size_t)nd, gd, sizeof(double) * (size_t)Bc * nd, cudaMemcpyDeviceToHost, s));
        } else {
            r = mtg_loglike_dev(model, N, S, group, Bc, (const double *)c.t.p, yd, dd, md, pd, ld, sd, flags, s);
            if (r) return r;
        }
        MTG_CUDA(cudaMemcpyAsync(out_ll + b0, ld, sizeof(double) * (size_t)Bc, cudaMemcpyDeviceToHost, s));
        if (status) MTG_CUDA(cudaMemcpyAsync(status + b0, sd, sizeof(int32_t) * (size_t)Bc, cudaMemcpyDeviceToHost, s));
        return 0;
    };
    const size_t wave2 = 2 * 148 * 512;                               // two K1 waves of 148 SMs x 4 CTAs x 128 threads
    // one shared data set (S == 1, group >= B): any cut is a whole-group cut, every chunk reads data set 0
    size_t chunk = (S == 1) ? wave2 : ((wave2 + group - 1) / group) * (size_t)group;
    if ((size_t)B < 2 * chunk) {
        if ((rc = run_chunk(0, B, c.stream))) return rc;
        MTG_CUDA(cudaStreamSynchronize(c.stream));
        return 0;
    }
    MTG_CUDA(cudaEventRecord(c.data_ready, c.stream));
    for (int i = 0; i < HostCtx::NPIPE; ++i) MTG_CUDA(cudaStreamWaitEvent(c.pipe[i], c.data_ready, 0));
    // The first chunk is ONE wave: nothing overlaps its parameter upload, so the device starts after half the copy time
    // (measured end to end on the C2 bench: see DESIGN.md section 1).
    const size_t first = (S == 1) ? wave2 / 2 : (((wave2 / 2) + group - 1) / group) * (size_t)group;
    int k = 0;
    for (size_t b0 = 0; b0 < (size_t)B; ++k) {
        const int Bc = (int)std::min(k == 0 ? first : chunk, (size_t)B - b0);
        if ((rc = run_chunk(b0, Bc, c.pipe[k % HostCtx::NPIPE]))) break;
        b0 += (size_t)Bc;
    }
    for (int i = 0; i < HostCtx::NPIPE; ++i) {
        cudaError_t e = cudaStreamSynchronize(c.pipe[i]);
        if (e != cudaSuccess && !rc) rc = mtg_cuda_check(e, "pipeline sync");
    }
    return rc;
}

extern "C" int mtg_loglike_host(const mtg_model_t *model, int N, int S, int group, int B,
                                const double *t, const double *y_sn, const double *diag_sn, const double *ds_mean,
                                const double *params, double *out_ll, int32_t *status, uint32_t flags)
{
    return host_eval(model, N, S, group, B, t, y_sn, diag_sn, ds_mean, params, out_ll, nullptr, status, flags);
}

extern "C" int mtg_loglike_grad_host(const mtg_model_t *model, int N, int S, int group, int B,
                                     const double *t, const double *y_sn, const double *diag_sn, const double *ds_mean,
                                     const double *params, double *out_ll, double *out_grad, int32_t *status, uint32_t flags)
{
    if (!out_grad) { mtg_set_error("mtg_loglike_grad_host: out_grad is NULL"); return MTG_EINVAL; }
    return host_eval(model, N, S, group, B, t, y_sn, diag_sn, ds_mean, params, out_ll, out_grad, status, flags);
}

extern "C" int mtg_lbfgs_host(const mtg_model_t *model, int N, int S, int group, int B,
                              const double *t, const double *y_sn, const double *diag_sn, const double *ds_mean,
                              const double *x0, const mtg_lbfgs_opts_t *opts,
                              double *out_x, double *out_nll, int32_t *out_iters, int32_t *status, uint32_t flags)
{
    if (!model || N < 1 || S < 1 || group < 1 || B < 0 || !t || !y_sn || !diag_sn || !x0 || !out_x || !out_nll) {
        mtg_set_error("mtg_lbfgs_host: bad argument"); return MTG_EINVAL;
    }
    if (B == 0) return 0;
    HostCtx *cp = host_ctx();
    if (!cp) return MTG_ENODEV;
    HostCtx &c = *cp;
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = c.ensure((int)(cp - g_ctx));
    if (rc) return rc;
    const int nd = model->n_free;
    if (nd < 1) { mtg_set_error("mtg_lbfgs_host: no free parameters"); return MTG_EINVAL; }
    const size_t xb = sizeof(double) * (size_t)B * nd;
    if ((rc = upload_data(c, N, S, t, y_sn, diag_sn, ds_mean, flags))) return rc;
    if ((rc = c.par.reserve(xb)) || (rc = c.x.reserve(xb)) || (rc = c.nll.reserve(sizeof(double) * B)) ||
        (rc = c.it.reserve(sizeof(int32_t) * B)) || (rc = c.st.reserve(sizeof(int32_t) * B))) return rc;
    MTG_CUDA(cudaMemcpyAsync(c.par.p, x0, xb, cudaMemcpyHostToDevice, c.stream));
    rc = mtg_lbfgs_dev(model, N, S, group, B, (const double *)c.t.p, (const double *)c.y.p, (const double *)c.d.p,
                       (const double *)c.mean.p, (const double *)c.par.p, opts, (double *)c.x.p, (double *)c.nll.p,
                       (int32_t *)c.it.p, (int32_t *)c.st.p, flags, c.stream);
    if (rc) return rc;
    MTG_CUDA(cudaMemcpyAsync(out_x, c.x.p, xb, cudaMemcpyDeviceToHost, c.stream));
    MTG_CUDA(cudaMemcpyAsync(out_nll, c.nll.p, sizeof(double) * B, cudaMemcpyDeviceToHost, c.stream));
    if (out_iters) MTG_CUDA(cudaMemcpyAsync(out_iters, c.it.p, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, c.stream));
    if (status) MTG_CUDA(cudaMemcpyAsync(status, c.st.p, sizeof(int32_t) * B, cudaMemcpyDeviceToHost, c.stream));
    MTG_CUDA(cudaStreamSynchronize(c.stream));
    return 0;
}

// ---- pinned host memory ------------------------------------------------------------------------
extern "C" int mtg_host_alloc(void **ptr, uint64_t bytes) { return mtg_cuda_check(cudaMallocHost(ptr, bytes), "cudaMallocHost"); }
extern "C" int mtg_host_free(void *ptr) { return mtg_cuda_check(cudaFreeHost(ptr), "cudaFreeHost"); }

// ---- measurement probes ------------------------------------------------------------------------
// 8 independent DFMA chains per thread, 1024 threads/SM resident: measures the FP64 FMA issue
// peak of the device (the roofline denominator for K1/K2; MEASURED_PEAKS.json has no FP64 entry).
__global__ void __launch_bounds__(256) k_fp64_probe(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.9999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;       // never true; keeps the chains alive
}

extern "C" int mtg_fp64_peak_probe(double *out_tflops, double *out_ms)
{
    int dev = 0, sms = 0;
    MTG_CUDA(cudaGetDevice(&dev));
    MTG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *d = nullptr;
    MTG_CUDA(cudaMalloc(&d, 64));
    cudaEvent_t e0, e1;
    MTG_CUDA(cudaEventCreate(&e0)); MTG_CUDA(cudaEventCreate(&e1));
    const int iters = 1 << 15, blocks = sms * 8, threads = 256;
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        MTG_CUDA(cudaEventRecord(e0, 0));
        k_fp64_probe<<<blocks, threads>>>(d, iters, 1.0);
        MTG_CUDA(cudaEventRecord(e1, 0));
        MTG_CUDA(cudaEventSynchronize(e1));
        float ms = 0; MTG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    mtg_count_launch(5);
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    if (out_tflops) *out_tflops = flops / (best * 1e-3) / 1e12;
    if (out_ms) *out_ms = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return 0;
}

__global__ void k_fill(double *p, size_t n, double v)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
extern "C" int mtg_l2_flush(void *scratch_dev, uint64_t bytes, void *stream)
{
    if (!scratch_dev || bytes < 8) { mtg_set_error("mtg_l2_flush: bad argument"); return MTG_EINVAL; }
    k_fill<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((double *)scratch_dev, bytes / 8, 0.0);
    mtg_count_launch();
    return mtg_cuda_check(cudaGetLastError(), "l2 flush launch");
}

// ---- test hook for the device math of the inner loop (tests/test_math_gpu.py) --------------------------
#include "mtg_math.cuh"
__global__ void k_test_math(int n, int variant, const double *__restrict__ x, double *__restrict__ e, double *__restrict__ s,
                            double *__restrict__ c, double *__restrict__ rc)
{
    __shared__ alignas(16) double tabs_sm[MTG_TABS_DOUBLES];
    extern __shared__ double2 big_sm[];                    // MTG_TABSB_DOUBLES doubles when variant == 2
    const mtgm::Tabs tb = mtgm::tabs_fill(tabs_sm);
    mtgm::TabsBig tbb{nullptr, nullptr};
    if (variant == 2) tbb = mtgm::tabs_fill_big(reinterpret_cast<double *>(big_sm));
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double sn, cs;
    if (variant == 0) { e[i] = mtgm::exp2_nonpos(x[i]); mtgm::sincos_fast(x[i], &sn, &cs); }
    else if (variant == 2) { e[i] = mtgm::exp2_nonpos_tab(x[i], tbb); mtgm::sincos_tab(x[i], tbb, &sn, &cs); }
    else { e[i] = mtgm::exp2_nonpos_tab(x[i], tb); mtgm::sincos_tab(x[i], tb, &sn, &cs); }
    s[i] = sn; c[i] = cs;
    rc[i] = mtgm::rcp_pos(fabs(x[i]));
}
extern "C" int mtg_test_math_host(int n, int variant, const double *x, double *out_exp, double *out_sin, double *out_cos,
                                  double *out_rcp)
{
    if (n < 1 || !x || !out_exp || !out_sin || !out_cos || !out_rcp) { mtg_set_error("mtg_test_math_host: bad argument"); return MTG_EINVAL; }
    double *d = nullptr;
    const size_t m = (size_t)n;
    MTG_CUDA(cudaMalloc((void **)&d, sizeof(double) * 5 * m));
    MTG_CUDA(cudaMemcpy(d, x, sizeof(double) * m, cudaMemcpyHostToDevice));
    const size_t dyn = variant == 2 ? sizeof(double) * MTG_TABSB_DOUBLES : 0;
    if (dyn) MTG_CUDA(cudaFuncSetAttribute(k_test_math, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    k_test_math<<<(n + 255) / 256, 256, dyn>>>(n, variant, d, d + m, d + 2 * m, d + 3 * m, d + 4 * m);
    mtg_count_launch();
    int rc = mtg_cuda_check(cudaMemcpy(out_exp, d + m, sizeof(double) * m, cudaMemcpyDeviceToHost), "d2h");
    if (!rc) rc = mtg_cuda_check(cudaMemcpy(out_sin, d + 2 * m, sizeof(double) * m, cudaMemcpyDeviceToHost), "d2h");
    if (!rc) rc = mtg_cuda_check(cudaMemcpy(out_cos, d + 3 * m, sizeof(double) * m, cudaMemcpyDeviceToHost), "d2h");
    if (!rc) rc = mtg_cuda_check(cudaMemcpy(out_rcp, d + 4 * m, sizeof(double) * m, cudaMemcpyDeviceToHost), "d2h");
    cudaFree(d);
    return rc;
}
