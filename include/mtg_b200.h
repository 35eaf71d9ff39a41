/*
 * mtg_b200.h -- C ABI of libmtg_b200.so: the B200 (sm_100a) implementation of the
 * mind_the_gaps hot path (celerite semiseparable GP log-likelihood + gradient, batched
 * L-BFGS MLE, lightcurve simulator, LRT statistics).
 *
 * The reference (andresgur/mind_the_gaps) is pure Python and has no FFI layer; the seam this
 * library replaces is the Python object protocol between GPModelling / Simulator and their
 * native third-party dependencies.  Each entry point cites the reference
 * interface it stands in for.  INTEGRATION.md shows the ctypes stubs a reference maintainer
 * would add.
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary.
 *   - every function returns 0 on success, a negative MTG_E* code for argument errors and a
 *     positive cudaError_t value for CUDA failures (mtg_last_error() gives the text).
 *   - *_dev functions take DEVICE pointers and enqueue on `stream` (a cudaStream_t passed as
 *     void*; NULL = the legacy default stream) without synchronising; they are re-entrant per
 *     stream and keep no global state.
 *   - *_host functions take HOST pointers, do their own H2D / D2H copies on an internal
 *     stream and return after the result is in the caller's buffer.
 *   - all floating point is IEEE binary64.
 *   - per-series `status`: 0 ok, 1 covariance not positive definite (celerite LinAlgError),
 *     2 non-finite result, 3 outside the prior box (only when bounds are applied).
 *
 * Data layout in HBM
 *   t      [N]        observation times, shared by every series of a call (ascending)
 *   y,diag [N][S]     TIME-MAJOR: element (n, s) at n*S + s, S = number of data sets.
 *                     Series b reads data set  b / group  (group = series per data set:
 *                     walkers of one lightcurve, multi-starts of one simulated lightcurve, or
 *                     B when one lightcurve is shared by the whole batch, S = 1).
 *   params [B][ndim]  row-major free (un-frozen) parameters, celerite parameter-vector order.
 */
#ifndef MTG_B200_H
#define MTG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTG_VERSION 100
#define MTG_MAX_TERMS 8
#define MTG_MAX_PARAMS 24
#define MTG_MAX_J 16

/* error codes (negative) */
#define MTG_EINVAL   (-1)   /* bad argument                                  */
#define MTG_EUNSUP   (-2)   /* kernel structure not supported (J too large)  */
#define MTG_ENODEV   (-3)   /* no CUDA device                                */
#define MTG_ENOMEM   (-4)

/* Kernel terms: log-parameter -> celerite coefficient maps.
 * Cites: mind_the_gaps/models/celerite_models.py:7-34 (Lorentzian), :36-52 (Cosinus),
 * :55-68 (DampedRandomWalk), :71-90 (BendingPowerlaw); celerite.terms RealTerm, ComplexTerm,
 * SHOTerm, Matern32Term, JitterTerm (un-vendored dependency, celerite>=0.4.2). */
enum mtg_term_kind {
    MTG_TERM_REAL = 0,        /* log_a, log_c                       -> real (a, c)                     */
    MTG_TERM_COMPLEX3 = 1,    /* log_a, log_c, log_d                -> complex (a, 0, c, d)            */
    MTG_TERM_COMPLEX4 = 2,    /* log_a, log_b, log_c, log_d         -> complex (a, b, c, d)            */
    MTG_TERM_SHO = 3,         /* log_S0, log_Q, log_omega0          -> complex (Q>=1/2) or 2 real      */
    MTG_TERM_MATERN32 = 4,    /* log_sigma, log_rho ; extra = eps   -> complex                         */
    MTG_TERM_JITTER = 5,      /* log_sigma                          -> diagonal e^{2 log_sigma}        */
    MTG_TERM_DRW = 6,         /* log_S0, log_omega0                 -> real (S0, w0)                   */
    MTG_TERM_LORENTZIAN = 7,  /* log_S0, log_Q, log_omega0          -> complex (S0,0,w0/2Q,w0) + real (0,0) */
    MTG_TERM_COSINUS = 8,     /* log_S0, log_omega0                 -> complex (S0, 0, 0, w0)          */
    MTG_TERM_BENDINGPL = 9,   /* log_S0, log_Q, log_omega0          -> complex (S0, Q, w0, w0); prior -inf if log_S0 < log_Q */
    /* RAW terms: the "parameters" ARE celerite coefficients, evaluated by the caller.  This is the entry for user-defined
     * celerite.terms.Term subclasses (arbitrary Python get_real_coefficients / get_complex_coefficients -- the reference's
     * own kernels are such subclasses, mind_the_gaps/models/celerite_models.py:7-90): the host evaluates the coefficient
     * map once per parameter set and hands the library params[B][n] = the concatenated coefficients.
     * Gradients come back with respect to the coefficients. */
    MTG_TERM_RAW_REAL = 10,   /* a, c                               -> real (a, c)                      */
    MTG_TERM_RAW_COMPLEX = 11,/* a, b, c, d                         -> complex (a, b, c, d)             */
    MTG_TERM_RAW_JITTER = 12  /* sigma^2                            -> diagonal                         */
};

/* Mean models (mind_the_gaps/gpmodelling.py:62-124, models/mean_models.py:6-31). */
enum mtg_mean_kind {
    MTG_MEAN_CONSTANT = 0,    /* value                                        */
    MTG_MEAN_LINEAR = 1,      /* slope, intercept                             */
    MTG_MEAN_GAUSSIAN = 2     /* mean, sigma, amplitude, constant             */
};

/* A GP model: the kernel term list, the mean model and the frozen/free parameter split of
 * celerite.GP (kernel parameters first, then mean parameters).  Passed BY POINTER from host
 * memory; the library copies it into kernel arguments. */
typedef struct mtg_model {
    int32_t n_terms;
    int32_t kind[MTG_MAX_TERMS];
    int32_t pidx[MTG_MAX_TERMS][4];     /* index of each term parameter in the FULL vector   */
    double  extra[MTG_MAX_TERMS];       /* Matern32 eps                                       */
    int32_t mean_kind;
    int32_t mean_pidx[4];               /* indices of the mean parameters in the FULL vector  */
    int32_t n_full;                     /* kernel + mean parameters, frozen included          */
    double  full_default[MTG_MAX_PARAMS];/* value of every parameter (used for frozen ones)   */
    int32_t n_free;                     /* ndim of the sampler / optimiser                    */
    int32_t free_idx[MTG_MAX_PARAMS];   /* free slot k -> index in the FULL vector            */
    double  lo[MTG_MAX_PARAMS];         /* box prior on free slot k (-inf / +inf = None)      */
    double  hi[MTG_MAX_PARAMS];
} mtg_model_t;

/* flags for mtg_loglike_* / mtg_grad_* */
#define MTG_FLAG_APPLY_PRIOR   1u  /* add celerite log_prior (box, BendingPowerlaw): out = -inf outside  (gpmodelling.py:147-152) */
#define MTG_FLAG_DS_MEAN       2u  /* residual = y - ds_mean[data set] (mean frozen at mean(y), gpmodelling.py:83-87)            */
#define MTG_FLAG_SCAN          4u  /* force the parallel-in-time kernels: K9 (mtg_loglike_*), K11 (mtg_loglike_grad_*), K10 (mtg_ensemble_*); MTG_EUNSUP if not covered */
#define MTG_FLAG_NO_SCAN       8u  /* force the one-thread-per-series kernels: K1, K2, K8                                           */

/* ---- library / device ------------------------------------------------------------------ */
int         mtg_version(void);
const char *mtg_last_error(void);
int         mtg_device_count(void);
int         mtg_set_device(int device);
/* Width of the celerite recursion the library will run for `model`:
 * J = Jr + 2 Jc as celerite counts it (j_celerite) and after dropping zero-amplitude real
 * terms that cannot influence the result (j_compute). */
int mtg_model_width(const mtg_model_t *model, int *j_celerite, int *j_compute);

/* ---- K1 / K9: batched log-likelihood ----------------------------------------------------
 * Replaces, for B parameter sets at once, celerite.GP.set_parameter_vector + log_prior +
 * log_likelihood as called from GPModelling._log_probability (gpmodelling.py:127-152) and
 * GPModelling._neg_log_like (:155-169); i.e. celerite CholeskySolver.compute + dot_solve +
 * log_determinant fused into one pass.  Two kernels compute the same number: K1 (one thread per
 * series: large batches) and K9 (one CTA per series, the N-step chain parallelised by an associative
 * scan: <= max(148, 2N) series of a model with <= 4 state dimensions); the library picks,
 * MTG_FLAG_SCAN / MTG_FLAG_NO_SCAN force either.
 *   out_ll[B]   log-likelihood (+ log-prior with MTG_FLAG_APPLY_PRIOR); -inf when status != 0
 *   status[B]   may be NULL
 *   ds_mean[S]  only read with MTG_FLAG_DS_MEAN                                             */
int mtg_loglike_dev(const mtg_model_t *model, int N, int S, int group, int B,
                    const double *t, const double *y, const double *diag, const double *ds_mean,
                    const double *params, double *out_ll, int32_t *status,
                    uint32_t flags, void *stream);
/* Host-buffer form.  y/diag are the natural numpy layout here: SERIES-major [S][N]; the
 * library transposes on the device. */
int mtg_loglike_host(const mtg_model_t *model, int N, int S, int group, int B,
                     const double *t, const double *y_sn, const double *diag_sn, const double *ds_mean,
                     const double *params, double *out_ll, int32_t *status, uint32_t flags);

/* ---- K2: log-likelihood + gradient (fused forward + reverse sweep) ----------------------
 * New capability (the reference differentiates by scipy finite differences inside
 * scipy.optimize.minimize(method="L-BFGS-B"), gpmodelling.py:192-193).
 *   out_grad[B][ndim]  d loglike / d free parameter                                        */
int mtg_loglike_grad_dev(const mtg_model_t *model, int N, int S, int group, int B,
                         const double *t, const double *y, const double *diag, const double *ds_mean,
                         const double *params, double *out_ll, double *out_grad, int32_t *status,
                         uint32_t flags, void *stream);
int mtg_loglike_grad_host(const mtg_model_t *model, int N, int S, int group, int B,
                          const double *t, const double *y_sn, const double *diag_sn, const double *ds_mean,
                          const double *params, double *out_ll, double *out_grad, int32_t *status,
                          uint32_t flags);

/* 1 when the parallel-in-time gradient kernel K11 (one CTA per series: forward scan + suffix scan of the adjoint, ~30 us
 * for <= 148 series at N = 1000 against K2's ~1.1 ms) covers `model` on series of N points -- <= 4 state dimensions and
 * series + tape within the shared memory of an SM -- i.e. when mtg_loglike_grad_* takes it for small batches
 * (MTG_FLAG_SCAN / MTG_FLAG_NO_SCAN force either kernel) and mtg_lbfgs_* runs the tail of a batch, or a whole small
 * batch, in its persistent kernel.  0 otherwise (K2 serves). */
int mtg_scan_grad_supported(const mtg_model_t *model, int N);

/* ---- K4: batched box-constrained L-BFGS maximum-likelihood fits -------------------------
 * Replaces scipy.optimize.minimize(_neg_log_like, x0, method="L-BFGS-B", bounds=...) of
 * GPModelling.fit (gpmodelling.py:172-194) for B independent problems in lock step.
 * Bounds come from model->lo / model->hi.
 *   x0[B][ndim] start points; out_x[B][ndim]; out_nll[B] = -loglike at out_x;
 *   out_iters[B] (may be NULL); status[B]: 0 converged, 1 iteration limit, 2 failed.       */
typedef struct mtg_lbfgs_opts {
    int32_t max_iter;     /* scipy default 15000; typical 200                */
    int32_t history;      /* m, scipy default 10; <= 10                      */
    double  ftol;         /* relative f decrease, scipy 2.2e-9               */
    double  gtol;         /* projected-gradient inf-norm, scipy 1e-5         */
    int32_t max_ls;       /* line-search evaluations per iteration, scipy 20 */
    int32_t reserved;
} mtg_lbfgs_opts_t;
int mtg_lbfgs_dev(const mtg_model_t *model, int N, int S, int group, int B,
                  const double *t, const double *y, const double *diag, const double *ds_mean,
                  const double *x0, const mtg_lbfgs_opts_t *opts,
                  double *out_x, double *out_nll, int32_t *out_iters, int32_t *status,
                  uint32_t flags, void *stream);
int mtg_lbfgs_host(const mtg_model_t *model, int N, int S, int group, int B,
                   const double *t, const double *y_sn, const double *diag_sn, const double *ds_mean,
                   const double *x0, const mtg_lbfgs_opts_t *opts,
                   double *out_x, double *out_nll, int32_t *out_iters, int32_t *status, uint32_t flags);


/* ---- simulator (K3 spectrum synthesis, cuFFT Z2D, K5 segment + exposure binning, K6 E13, K7 noise)
 * Replaces Simulator.generate_lightcurve + add_noise (mind_the_gaps/simulator.py:369-420, 300-338;
 * noise_models.py) for L lightcurves at once.  A simulator object owns the observing window
 * (Simulator.__init__, simulator.py:195-258): the regular grid, the bin edges ("strategy"), the
 * exposures / backgrounds, cuFFT plans and chunk buffers. */
enum mtg_pdf_kind   { MTG_PDF_GAUSSIAN = 0, MTG_PDF_LOGNORMAL = 1, MTG_PDF_UNIFORM = 2 };
enum mtg_noise_kind { MTG_NOISE_POISSON = 0, MTG_NOISE_KRAFT = 1, MTG_NOISE_GAUSSIAN = 2 };

/* mtg_sim_config_t.flags.  MTG_SIM_E13_PADDED (set with the lengthened "fast" grid, never on the reference's own grid): the E13
 * PDF adjustment may run on the segment lengthened to the next 7-smooth length (it needs an FFT pair of the segment's
 * length per iteration; the window's own length is usually Bluestein territory).  The extra samples come from the same
 * synthesised series; only the first `duration` of the adjusted segment is binned. */
#define MTG_SIM_E13_PADDED 1

typedef struct mtg_sim_config {
    int32_t n_obs;              /* N observation bins                                              */
    int32_t n_fft;              /* len(sim_timestamps)            (simulator.py:234-238)           */
    double  sim_dt;             /* min(exposures)/aliasing_factor (simulator.py:217)               */
    double  sim_duration;       /* length of the random segment   (simulator.py:226-228)           */
    double  mean;               /* target mean count rate                                          */
    int32_t pdf;                /* mtg_pdf_kind: Gaussian -> TK95, else E13 (simulator.py:207-214) */
    int32_t max_iter;           /* E13 iterations (default 400)                                    */
    int32_t noise;              /* mtg_noise_kind                 (simulator.py:245-251)           */
    int32_t flags;              /* MTG_SIM_* bits below                                           */
    double  sigma_noise;        /* Gaussian noise sigma                                            */
    double  kraft_counts;       /* Kraft threshold, 15            (noise_models.py:87)             */
    const double *grid;         /* host [n_fft]  sim_timestamps exactly as np.arange produced them */
    const double *bin_lo;       /* host [n_obs]  strategy[i][0]   (simulator.py:254-255)           */
    const double *bin_hi;       /* host [n_obs]  strategy[i][1]                                    */
    const double *exposures;    /* host [n_obs]                                                    */
    const double *bkg_counts;   /* host [n_obs] or NULL (zeros): bkg_rate * exposures              */
    const double *bkg_rate_err; /* host [n_obs] or NULL (zeros)                                    */
} mtg_sim_config_t;

int mtg_sim_create(const mtg_sim_config_t *cfg, void **sim);
int mtg_sim_destroy(void *sim);
/* Destroyed simulators park their chunk buffers and cuFFT plans (gigabytes; Bluestein plans for the reference's
 * un-padded grid lengths) for the next simulator of the same FFT length on the same device; at most MTG_SIM_POOL
 * (environment, default 4, 0 disables) workspaces are kept.  mtg_sim_trim releases them. */
int mtg_sim_trim(void);
/* 1 when this simulator's binned Gaussian-PDF lightcurves come from the fused synthesis + binning kernels
 * (csrc/mtg_tk95.cu: n_fft = 2^p x N2 and a window sparse enough that only a fraction of the series is needed), 0 when
 * they come from the spectrum -> cuFFT -> binning chain.  Environment MTG_SIM_FUSED = 0 | 1 (default) | 2 (whenever the
 * grid factorises), read by mtg_sim_create. */
int mtg_sim_is_fused(void *sim);

/* Generate L noiseless binned lightcurves.
 * PSD source: `model` + params[L][n_free] (celerite kernel PSD evaluated on the device, the
 * generate_from_posteriors path gpmodelling.py:533-536) OR psd_table[n_fft/2+1] = psd(rfftfreq*2pi)
 * evaluated by the caller for an arbitrary callable (one PSD for all L).
 * Random inputs: Philox4x32-10 streams keyed by `seed` unless the caller injects them
 * (parity tests): normals[L][2][n_fft/2+1] (the two rows of simulator.py:492), u_segment[L] (the
 * random_sample() behind np.random.uniform, simulator.py:538), pdf_normals[L][n_seg_max]
 * (standard normal / uniform(0,1) variates turned into pdf.rvs draws, simulator.py:90).
 *   out_rates[L][n_obs]   (host form)  /  out_rates_tm[n_obs][L] time-major (device form)
 *   out_seg[L][2]         first index and length of the random segment (may be NULL)            */
int mtg_sim_generate_host(void *sim, int L, const mtg_model_t *model, const double *params,
                          const double *psd_table, uint64_t seed,
                          const double *normals, const double *u_segment, const double *pdf_variates,
                          double *out_rates, int32_t *out_seg, int32_t *out_e13_iters);
int mtg_sim_generate_dev(void *sim, int L, const mtg_model_t *model, const double *params_dev,
                         const double *psd_table_dev, uint64_t seed,
                         double *out_rates_tm, void *stream);
/* Add noise and estimate uncertainties (noise_models.py).  counts_in[L][n_obs] (the
 * np.random.poisson draws) / normals_in[L][n_obs] inject the randomness for parity tests.
 * status[L][n_obs]: 0 ok, 1 negative expected counts (numpy raises ValueError("lam < 0")).     */
int mtg_sim_add_noise_host(void *sim, int L, const double *rates, uint64_t seed,
                           const int64_t *counts_in, const double *normals_in,
                           double *out_rates, double *out_dy, int32_t *status);
int mtg_sim_add_noise_dev(void *sim, int L, const double *rates_tm, uint64_t seed,
                          double *out_rates_tm, double *out_dy_tm, void *stream);

/* PoissonNoise / KraftNoise / GaussianNoise.add_noise (noise_models.py:49-184) on explicit host arrays:
 * rates[L][n_obs]; exposures / bkg_counts / bkg_rate_err [n_obs] (the last two may be NULL = zeros). */
int mtg_noise_host(int kind, int L, int n_obs, const double *rates, const double *exposures,
                   const double *bkg_counts, const double *bkg_rate_err, double sigma, double kraft_counts,
                   uint64_t seed, const int64_t *counts_in, const double *normals_in,
                   double *out_rates, double *out_dy, int32_t *status);
/* Simulator.simulate_regularly_sampled (simulator.py:369-394): the full-resolution TK95 count rates
 * out_countrate[L][n_fft] (mean included), same PSD sources / injected normals as generate. */
int mtg_sim_regular_host(void *sim, int L, const mtg_model_t *model, const double *params,
                         const double *psd_table, uint64_t seed, const double *normals, double *out_countrate);

/* Simulator.downsample (simulator.py:340-367) on explicit arrays: per-bin np.mean of countrate over
 * {j : time[j] >= lo && time[j] < hi}, summed in numpy's pairwise order (bit-exact), plus the index
 * range of every bin.  time must be non-decreasing.                                              */
int mtg_downsample_host(int n_seg, const double *time, const double *countrate, int n_obs,
                        const double *bin_lo, const double *bin_hi,
                        double *out_rates, int32_t *out_first, int32_t *out_count);
/* E13Simulator.adjust_lightcurve_pdf (simulator.py:65-131) for one lightcurve on explicit draws. */
int mtg_e13_adjust_host(int n, const double *countrate, const double *xsim0, int max_iter,
                        double *out_x, int32_t *out_iters);

/* ---- LRT pipeline helpers (device-resident data, time-major [N][S]) -------------------------
 * diag = (dy + 1e-12)^2 (gpmodelling.py:54) and mean_out[s] = mean_n y[n][s] (gpmodelling.py:83-87). */
int mtg_prepare_data_dev(int N, int S, const double *y_tm, const double *dy_tm, double *diag_tm,
                         double *mean_out, void *stream);
/* best[s] = min over the `group` consecutive problems of data set s (multi-start), arg[s] = its index */
int mtg_group_best_dev(int S, int group, const double *f, double *best, int32_t *arg, void *stream);
/* celerite GP.predict(y, return_var=True, return_cov=False) at the training times
 * (GPModelling.standarized_residuals, gpmodelling.py:353-370) for ONE parameter vector. */
int mtg_predict_host(const mtg_model_t *model, int N, const double *t, const double *y, const double *diag,
                     const double *params, double *out_mu, double *out_var);

/* ---- K8: device-resident affine-invariant ensemble sampler ------------------------------------
 * Replaces the emcee 3.1.4 loop GPModelling.derive_posteriors drives (gpmodelling.py:247-264):
 * EnsembleSampler.sample -> RedBlueMove (two random halves) -> StretchMove (scale a) -> log-probability of the
 * half ensemble = GPModelling._log_probability (:127-152) -> accept / reject, for `nsteps` steps and E independent
 * ensembles in ONE launch.  Two kernels walk the same chain from the same Philox streams: K8 (one CTA per ensemble, one
 * thread per proposal, positions in shared memory: thousands of ensembles) and K10 (one CTA per PROPOSAL with the
 * parallel-in-time log-probability, a barrier across the CTAs of an ensemble per half step, cooperative launch: the
 * one-lightcurve case, taken when all E * ceil(W/2) CTAs fit the device and the model has <= 4 state dimensions;
 * MTG_FLAG_SCAN / MTG_FLAG_NO_SCAN force either).
 *   S        1: every ensemble samples the same lightcurve; E: ensemble e samples data set e (y, diag [N][S])
 *   pos      [E][W][ndim] walker positions, updated in place;  lp [E][W] their log-probabilities
 *   init_lp  non-zero: lp is computed from pos first (start of a run); zero: lp is read (continuation)
 *   step0    absolute index of the first step (keeps the Philox stream independent of how a run is cut)
 *   chain    [E][nsteps][W][ndim] (emcee's (steps, walkers, ndim) per ensemble), chain_lp [E][nsteps][W]; may be NULL
 *   n_accept [E][W] accepted moves, ACCUMULATED; may be NULL
 *   best_lp  [E] largest log-probability any walker of the ensemble visited during the run (what the reference's LRT
 *            keeps: GPModelling.max_loglikelihood, gpmodelling.py:405-410), best_pos [E][ndim] where; may be NULL
 * The box prior is always applied (outside: log-probability -inf, as _log_probability).            */
int mtg_ensemble_dev(const mtg_model_t *model, int N, int S, int E, int W,
                     const double *t, const double *y, const double *diag, const double *ds_mean,
                     double *pos, double *lp, int init_lp, int nsteps, double a, uint64_t seed, uint64_t step0,
                     double *chain, double *chain_lp, int32_t *n_accept, double *best_lp, double *best_pos,
                     uint32_t flags, void *stream);
/* host buffers; here y, diag are [S][N] (one lightcurve per row) */
int mtg_ensemble_host(const mtg_model_t *model, int N, int S, int E, int W,
                      const double *t, const double *y_rows, const double *diag_rows, const double *ds_mean,
                      double *pos, double *lp, int init_lp, int nsteps, double a, uint64_t seed, uint64_t step0,
                      double *chain, double *chain_lp, int32_t *n_accept, uint32_t flags);

/* Test hook for sampler parity: the same run with the random draws of emcee 3.1.4 SUPPLIED by the caller instead of the
 * Philox streams -- split[E][nsteps][W] = the shuffled `inds` of RedBlueMove.propose (0 / 1, ceil(W/2) zeros; walkers of a
 * split are taken in ascending index order, as emcee does), and per step and half h, for the i-th walker of split h:
 * uz = random.rand(Ns)[i] and rint = random.randint(Nc, size=Ns)[i] of StretchMove.get_proposal, uacc = the
 * model.random.rand() of its accept test; all [E][nsteps][2][ceil(W/2)].  Proposal arithmetic is evaluated in emcee's
 * order without FMA contraction, so the chain equals oracle/sampler.py's bit for bit (tests/test_ensemble_gpu.py).
 * Log-probabilities of the start positions are always computed (init_lp = 1). */
int mtg_ensemble_injected_host(const mtg_model_t *model, int N, int S, int E, int W,
                               const double *t, const double *y_rows, const double *diag_rows, const double *ds_mean,
                               double *pos, double *lp, int nsteps, double a,
                               const int32_t *split, const double *uz, const int32_t *rint, const double *uacc,
                               double *chain, double *chain_lp, int32_t *n_accept, uint32_t flags);

/* ---- batched periodograms of an ensemble ---------------------------------
 * mtg_power_spectrum_host: |rfft(rate)|^2 of L regularly sampled lightcurves rates[L][n] -> out_power[L][n/2+1]
 *   (replaces the per-lightcurve numpy.fft.rfft loops of tests/simulator_test.py:24-35 and
 *   docs/notebooks/lomb_scargle_biases.ipynb).
 * mtg_lombscargle_host: generalised Lomb-Scargle periodogram of L lightcurves y[L][N] sharing the times t[N], at the M
 *   frequencies freqs[M] (cycles per unit of t) -> out_power[L][M]; fit_mean / center_data / normalization as
 *   astropy.timeseries.LombScargle(t, y, fit_mean=, center_data=, normalization=).power(freqs) with unit errors, the call
 *   the reference's notebook makes once per lightcurve. */
#define MTG_LS_STANDARD 0
#define MTG_LS_MODEL    1
#define MTG_LS_LOG      2
#define MTG_LS_PSD      3
int mtg_power_spectrum_host(int L, int n, const double *rates, double *out_power);
/* Normalised autocorrelation functions of L series x[L][n_t] -> out_acf[L][n_t] (emcee autocorr.function_1d: mean removed,
 * zero padded to 2 * 2^k, FFT, |.|^2, inverse FFT, divided by lag 0): the convergence diagnostic of
 * GPModelling.derive_posteriors (gpmodelling.py:251-263 -> emcee get_autocorr_time) for all walkers and dimensions at once. */
int mtg_autocorr_host(int L, int n_t, const double *x, double *out_acf);
int mtg_lombscargle_host(int N, int L, int M, const double *t, const double *y, const double *freqs,
                         int fit_mean, int center_data, int normalization, double *out_power);

/* ---- pinned host memory -----------------------------------------------------------------
 * Page-locked buffers for the *_host entry points (cudaMallocHost / cudaFreeHost). */
int mtg_host_alloc(void **ptr, uint64_t bytes);
int mtg_host_free(void *ptr);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* Dependent-free DFMA microbenchmark: returns achieved FP64 TFLOP/s (2 flop per FMA) on the
 * current device, timed with CUDA events.  Used by bench.py as the FP64 roofline denominator. */
int mtg_fp64_peak_probe(double *out_tflops, double *out_ms);
/* Writes `bytes` of device memory (L2 flush between timed iterations). */
int mtg_l2_flush(void *scratch_dev, uint64_t bytes, void *stream);
/* Test hook: the inner loop's own 2^x (x <= 0), sincos and 1/|x| evaluated on x[n].
 * variant 0 = polynomial kernels, 1 = the table-driven kernels K1/K2 run (csrc/mtg_math.cuh). */
int mtg_test_math_host(int n, int variant, const double *x, double *out_exp, double *out_sin, double *out_cos,
                       double *out_rcp);
/* number of kernel launches issued by this library since load (bench.py gpu_launches claim) */
uint64_t mtg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MTG_B200_H */
