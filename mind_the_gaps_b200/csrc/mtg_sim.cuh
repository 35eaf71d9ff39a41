// mtg_sim.cuh -- simulator object shared by mtg_sim.cu (TK95, binning, noise) and mtg_e13.cu.
#pragma once
#include "mtg_common.cuh"
#include <cufft.h>
#include <map>
#include <vector>

struct PsdCoef {                 // celerite coefficients of one parameter vector (runtime width)
    int nr, nc;
    double ar[MTG_MAXR + 2 * MTG_MAXC], cr[MTG_MAXR + 2 * MTG_MAXC];
    double ac[MTG_MAXC], bc[MTG_MAXC], cc[MTG_MAXC], dc[MTG_MAXC];
};

struct SegInfo { int i0, i1; double time_shift; };      // segment = grid[i0:i1] + time_shift

struct Sim {
    mtg_sim_config_t cfg;        // host pointers nulled after upload
    int M = 0;                   // n_fft / 2 + 1
    double first_bin_start = 0.0;
    double *d_grid = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_exp = nullptr, *d_bkg = nullptr, *d_err = nullptr;
    cudaStream_t stream = nullptr;
    int cap = 0;                 // lightcurves the chunk buffers hold
    double2 *d_spec = nullptr; double *d_lc = nullptr; PsdCoef *d_coef = nullptr; SegInfo *d_seg = nullptr;
    std::map<long long, cufftHandle> plans, plans_d2z;      // (n << 20 | batch) -> plan
};

int sim_get_plan(Sim *s, std::map<long long, cufftHandle> &cache, cufftType type, int n, int batch, cufftHandle *out);
// E13 PDF adjustment of the segments of the current chunk (in place in s->d_lc)
int sim_e13_chunk(Sim *s, int Lc, int l0, uint64_t seed, const double *d_pdf_variates, int32_t *d_iters, cudaStream_t stream);
