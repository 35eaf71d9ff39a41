// mtg_ensemble.cuh -- arguments shared by the two ensemble-sampler kernels (K8: mtg_ensemble.cu, K10: mtg_scan.cu)
#pragma once
#include "mtg_common.cuh"

// Test hook: the random draws of emcee's RedBlueMove.propose / StretchMove.get_proposal supplied by the caller instead of
// the Philox streams, so that a run can be compared step by step with the NumPy restatement of emcee (oracle/sampler.py).
struct EnsembleInject {
    const int32_t *split;   // [E][nsteps][W]    the shuffled `inds` (0 / 1, ceil(W/2) zeros) of RedBlueMove.propose
    const double  *uz;      // [E][nsteps][2][H] random.rand(Ns) of StretchMove.get_proposal, H = ceil(W/2)
    const int32_t *rint;    // [E][nsteps][2][H] random.randint(Nc, size=Ns)
    const double  *uacc;    // [E][nsteps][2][H] the model.random.rand() of each accept test, in walker order
};

struct EnsembleArgs {
    int N, S, E, W, nsteps;
    const double *t, *y, *diag, *ds_mean;
    double *pos, *lp;                 // [E][W][nd], [E][W]   in / out
    double *chain, *chain_lp;         // [E][nsteps][W][nd], [E][nsteps][W]   (may be NULL)
    int32_t *n_accept;                // [E][W] accumulated (may be NULL)
    double *best_lp, *best_pos;       // [E], [E][nd]: running maximum of the log-probability over the run (may be NULL)
    double a;                         // stretch scale
    uint64_t seed, step0;
    uint32_t flags; int init_lp;
    int epc;                          // ensembles per CTA (small ensembles share a warp)
    EnsembleInject inj;               // all NULL: Philox streams
};

// The stretch proposal and its acceptance ratio, written with explicitly rounded operations (no FMA contraction) in the
// order emcee evaluates them -- zz = ((a - 1) u + 1)^2 / a;  q = c - (c - s) zz;  lnpdiff = (ndim - 1) log zz + lq - lp --
// so that a run on injected draws reproduces the NumPy restatement bit for bit.
__device__ __forceinline__ double stretch_zz(double a, double u)
{ const double g = __dadd_rn(__dmul_rn(a - 1.0, u), 1.0); return __dmul_rn(g, g) / a; }
__device__ __forceinline__ double stretch_q(double pc, double ps, double zz)
{ return __dsub_rn(pc, __dmul_rn(__dsub_rn(pc, ps), zz)); }
__device__ __forceinline__ double stretch_lnpdiff(int nd, double zz, double lq, double lp)
{ return __dsub_rn(__dadd_rn(__dmul_rn((double)(nd - 1), log(zz)), lq), lp); }
// perm for an injected split: the walkers with inds == 0 in ascending order, then those with inds == 1 (emcee's
// all_inds[S1] / sets[j] ordering).  Rank of walker w.
__device__ __forceinline__ int injected_rank(const int32_t *sp, int w, int n0)
{
    const int32_t my = sp[w];
    int r = my ? n0 : 0;
    for (int k = 0; k < w; ++k) r += (sp[k] == my);
    return r;
}

// K10 (scan sampler): covered model, and all E * ceil(W/2) CTAs co-resident on the device
bool mtg_k10_fits(const Plan &pl, int N, int E, int W);
int  mtg_k10_dispatch(const Plan &pl, const EnsembleArgs &a, cudaStream_t stream);
