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

// Geometry of the fused TK95 path (mtg_tk95.cu): n_fft = N1 * N2 with N1 = 2^logN1 <= 8192 the length of the
// shared-memory column FFTs, K2 = N2/2 + 1 columns per lightcurve, Qmax sample slots per lightcurve.
struct TkGeom {
    int ok;                      // 0: this window / grid runs on cuFFT
    int n, N1, N2, K2, M, logN1;
    int npass, lograd[5], F;     // radix-8 / radix-4 DIF passes in shared memory, then a register FFT on F = 8 | 16 contiguous points
    int tw_off[5], tw_total;     // twiddle tables of the passes (double2 offsets)
    int Qmax, xs_max;            // sample slots per lightcurve; most slots in any group of TK_BINS_PER_CTA bins
};
#define TK_BINS_PER_CTA 64

struct Sim {
    mtg_sim_config_t cfg;        // host pointers nulled after upload
    int M = 0;                   // n_fft / 2 + 1
    double first_bin_start = 0.0;
    double *d_grid = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_exp = nullptr, *d_bkg = nullptr, *d_err = nullptr;
    cudaStream_t stream = nullptr;
    int cap = 0;                 // lightcurves the chunk buffers hold
    double2 *d_spec = nullptr; double *d_lc = nullptr; PsdCoef *d_coef = nullptr; SegInfo *d_seg = nullptr;
    std::map<long long, cufftHandle> plans, plans_d2z;      // (n << 20 | batch) -> plan
    // fused TK95 path
    TkGeom tk{};
    double2 *d_tw = nullptr; int *d_qoff = nullptr;          // per simulator (window / grid constants)
    int capF = 0;                                             // lightcurves the fused chunk buffers hold
    double2 *d_Y = nullptr; int *d_mslot = nullptr, *d_cnt = nullptr; int *d_flag = nullptr;
    PsdCoef *d_coefF = nullptr; SegInfo *d_segF = nullptr; double *d_ampt = nullptr;
    void *e13_cache = nullptr;                                // E13Work of mtg_e13.cu, kept between chunks
    void *e13_large = nullptr;                                // E13Large of mtg_e13.cu (slots of the long-series path)
};

int sim_get_plan(Sim *s, std::map<long long, cufftHandle> &cache, cufftType type, int n, int batch, cufftHandle *out);
// E13 PDF adjustment of the segments of the current chunk (in place in s->d_lc)
int sim_e13_chunk(Sim *s, int Lc, int l0, uint64_t seed, const double *d_pdf_variates, int32_t *d_iters, cudaStream_t stream);
void sim_e13_release(Sim *s);

// ---- fused TK95 path (mtg_tk95.cu) --------------------------------------------------------------------
// host: geometry for a grid / window (ok = 0 when the path does not apply), twiddle tables
void tk_geometry(int n_fft, int n_obs, const double *bin_lo, const double *bin_hi, double sim_dt, TkGeom *g, std::vector<int> *q_off);
void tk_twiddles(const TkGeom &g, std::vector<double2> *tw);
// column FFTs of a chunk: Y[l][k2][q] for the sample slots mslot[l][q] (absolute sample index or -1)
int tk_launch_columns(const TkGeom &g, int Lc, double val, double fold, const PsdCoef *coefs, const double *psd_table,
                      const double *amp_t, const double *normals, uint64_t seed, uint64_t lc_offset, const double2 *tw,
                      const int *mslot, double2 *Y, cudaStream_t stream);
// amplitude table [K2][N1] of a host-evaluated PSD in column order (Philox path)
int tk_launch_amp_table(const TkGeom &g, double fold, const double *psd_table, double *amp_t, cudaStream_t stream);
// row sums + exposure binning: out_tm[n][l_out0 + l]
int tk_launch_rows(const TkGeom &g, int Lc, int n_obs, const int *q_off, const int *mslot, const int *cnt, const double2 *Y,
                   double mean_add, double *out_tm, int L_total, int l_out0, cudaStream_t stream);
