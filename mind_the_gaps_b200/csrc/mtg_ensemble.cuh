// mtg_ensemble.cuh -- arguments shared by the two ensemble-sampler kernels (K8: mtg_ensemble.cu, K10: mtg_scan.cu)
#pragma once
#include "mtg_common.cuh"

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
};

// K10 (scan sampler): covered model, and all E * ceil(W/2) CTAs co-resident on the device
bool mtg_k10_fits(const Plan &pl, int N, int E, int W);
int  mtg_k10_dispatch(const Plan &pl, const EnsembleArgs &a, cudaStream_t stream);
