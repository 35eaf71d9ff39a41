"""GPU: the null distribution of the LRT statistic is pinned on the CPU reference procedure.

The device pipeline (simulate -> batched multi-start L-BFGS for both models -> T = -2 (L_null - L_alt)) replaces the
notebook loop of the reference (docs/notebooks/tutorial_ppp.ipynb cells 11-15).  Here the SAME simulated lightcurves are
refitted on the CPU exactly as GPModelling.fit does -- scipy's L-BFGS-B with finite differences on the C restatement of
celerite (oracle/celerite_ref.c) -- from the SAME start points, and the two sets of statistics must agree simulation by
simulation wherever the likelihood has one relevant maximum, and statistically over the wide frequency band.  The last
test shows where the level of the distribution comes from: T is the maximum of the likelihood ratio over the QPO
frequency range, so a more intense search (more starts, or the reference's MCMC refinement) raises it (look-elsewhere
effect), and the single-start fit reproduces the lower, chi^2-like level."""
import numpy as np
import pytest
from scipy.optimize import minimize
from scipy.stats import ks_2samp

from mind_the_gaps_b200 import lrt
from mind_the_gaps_b200.celerite import GP
from mind_the_gaps_b200.fitting import multistart_points
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from oracle import cref

from test_pipeline_gpu import DAY, make_lc

pytestmark = pytest.mark.gpu

B_DRW = dict(log_S0=(-15, 5), log_omega0=(-20, -5))
B_LOR = dict(log_S0=(-20, 5), log_Q=(0.5, 7), log_omega0=(-16, -9))


def _problem():
    lc, truth = make_lc(200, 4)
    null_kernel = DampedRandomWalk(truth[0], truth[1], bounds=B_DRW)
    alt_kernel = Lorentzian(truth[0] - 2, np.log(20.0), np.log(2 * np.pi / (10 * DAY)), bounds=B_LOR) + \
        DampedRandomWalk(truth[0], truth[1], bounds=B_DRW)
    return lc, truth, null_kernel, alt_kernel


def _cpu_best(spec, starts, bounds, t, y, dy):
    best = np.inf
    for x0 in starts:
        f = lambda p: (lambda v: v if np.isfinite(v) else 1e25)(-cref.loglike_params(spec, p[None, :], t, y, dy, nthreads=1)[0])
        sol = minimize(f, x0, method="L-BFGS-B", bounds=bounds)          # no jac: 2-point differences, as gpmodelling.py:192-193
        best = min(best, sol.fun)
    return -best


def _device_vs_cpu(alt_bounds_lor, n_sims, seed, ns1):
    """Run the device LRT, refit the same lightcurves on the CPU from the same starts; returns everything needed."""
    ns0 = 2
    lc, truth = make_lc(200, 4)
    null_kernel = DampedRandomWalk(truth[0], truth[1], bounds=B_DRW)
    alt_kernel = Lorentzian(truth[0] - 2, np.log(20.0), np.log(2 * np.pi / (10 * DAY)), bounds=alt_bounds_lor) + \
        DampedRandomWalk(truth[0], truth[1], bounds=B_DRW)
    samples = truth[None, :] + np.random.default_rng(5).normal(0, 0.15, (n_sims, 2))
    res = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, n_starts_null=ns0, n_starts_alt=ns1, seed=seed,
                               return_lightcurves=True, chunk=n_sims, pipeline=1)
    Y, DY = res.lightcurves
    assert np.all(np.isfinite(res.T)) and np.all(res.T > -1e-4)
    # the start points of the device run (lrt.simulate_and_fit draws them from default_rng([seed, rank, chunk offset]))
    gp1 = GP(alt_kernel, mean=0.0)
    b0 = [B_DRW["log_S0"], B_DRW["log_omega0"]]
    b1 = [alt_bounds_lor["log_S0"], alt_bounds_lor["log_Q"], alt_bounds_lor["log_omega0"]] + b0
    rng = np.random.default_rng([seed, 0, 0])
    x0 = multistart_points(samples, b0, ns0, rng).reshape(n_sims, ns0, 2)
    lo1, hi1 = lrt._bounds_arrays(gp1)
    slots = lrt.infer_null_slots(null_kernel, alt_kernel)
    x1 = lrt.alternative_starts(gp1.get_parameter_vector().copy(), res.params_null, slots, lo1, hi1, ns1, rng)
    l0 = np.array([_cpu_best([("drw",)], x0[i], b0, lc.times, Y[i], DY[i]) for i in range(n_sims)])
    l1 = np.array([_cpu_best([("lorentzian",), ("drw",)], x1[i], b1, lc.times, Y[i], DY[i]) for i in range(n_sims)])
    return lc, res, Y, DY, l0, l1, b1


def test_null_fits_and_single_band_alternative_match_cpu_refits_per_simulation():
    """Where the likelihood surface has ONE relevant maximum the two procedures must agree simulation by simulation:
    the null model always, and the alternative once the QPO frequency is confined to about one resolution element
    (the simulations contain no QPO, so over a wide band the Lorentzian's likelihood has a local maximum per
    independent frequency: next test)."""
    w0 = np.log(2 * np.pi / (10 * DAY))
    band = dict(log_S0=(-20, 5), log_Q=(0.5, 7), log_omega0=(w0 - 0.02, w0 + 0.02))
    lc, res, Y, DY, l0, l1, b1 = _device_vs_cpu(band, n_sims=200, seed=11, ns1=8)
    np.testing.assert_allclose(res.loglike_null, l0, rtol=1e-9, atol=1e-6)          # null model: unimodal
    T_cpu = -2.0 * (l0 - l1)
    d = np.abs(res.T - T_cpu)
    frac = np.mean(d < 1e-3)
    # measured: 89 % agree to 1e-3 with a median |dT| of 5e-9; the rest sit in different local maxima of the (S0, Q) plane,
    # in both directions
    assert frac >= 0.85 and np.median(d) < 1e-6, "only %.1f %% of the simulations agree to 1e-3 (median |dT| %.2e, worst %.3f)" % (100 * frac, np.median(d), d.max())
    assert np.mean(res.T >= T_cpu - 1e-3) >= 0.92
    assert ks_2samp(res.T, T_cpu).pvalue > 0.01
    assert abs(np.median(res.T) - np.median(T_cpu)) < 0.05


def test_wide_band_null_distribution_matches_cpu_refits_statistically():
    """The bench / tutorial configuration: log_omega0 free over 7 e-folds (~10^3 independent frequencies for this window).
    Two local optimisers started from the same 8 points end in different local maxima for many simulations -- in both
    directions -- so agreement is (i) exact for the null fits, (ii) statistical for T: same distribution (KS), same
    median, the device never the systematically worse; and (iii) the device optima are genuine local maxima: scipy
    restarted there cannot improve them (but for the rare plateau stop of the relative-reduction test both share)."""
    lc, res, Y, DY, l0, l1, b1 = _device_vs_cpu(B_LOR, n_sims=200, seed=11, ns1=8)
    np.testing.assert_allclose(res.loglike_null, l0, rtol=1e-9, atol=1e-6)
    T_cpu = -2.0 * (l0 - l1)
    assert ks_2samp(res.T, T_cpu).pvalue > 0.01
    assert abs(np.median(res.T) - np.median(T_cpu)) < 0.8
    better_dev, better_cpu = np.mean(res.T > T_cpu + 1e-3), np.mean(T_cpu > res.T + 1e-3)
    # measured: device better 32-39 %, CPU better 19-29 % (it moves with the rounding of the optimiser's arithmetic): the
    # device procedure must not be the systematically worse one
    assert better_cpu < better_dev + 0.15, (better_dev, better_cpu)
    spec = [("lorentzian",), ("drw",)]
    gain = np.empty(40)
    for k, i in enumerate(range(0, 200, 5)):
        f = lambda p: (lambda v: v if np.isfinite(v) else 1e25)(-cref.loglike_params(spec, p[None, :], lc.times, Y[i], DY[i], nthreads=1)[0])
        np.testing.assert_allclose(-f(res.params_alt[i]), res.loglike_alt[i], rtol=1e-10)
        gain[k] = res.loglike_alt[i] + minimize(f, res.params_alt[i], method="L-BFGS-B", bounds=b1).fun
    # the relative-reduction stop (scipy's own factr test) can fire on a plateau: a restart without history then creeps on.
    # Measured: 0-1 of these 40 optima can be improved by more than 1e-3 (by 0.35 in log-likelihood at worst)
    assert np.mean(gain > -1e-3) >= 0.95 and np.median(gain) > -1e-6, "scipy improved %d of 40 device optima, the worst by %.3g" % (np.sum(gain <= -1e-3), -gain.min())


def test_level_of_the_null_distribution_is_a_look_elsewhere_effect():
    n_sims = 256
    lc, truth, null_kernel, alt_kernel = _problem()
    samples = truth[None, :] + np.random.default_rng(6).normal(0, 0.15, (n_sims, 2))
    kw = dict(n_starts_null=2, seed=3, chunk=n_sims, pipeline=1)
    one = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, n_starts_alt=1, **kw)      # the nested start only
    eight = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, n_starts_alt=8, **kw)
    many = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, n_starts_alt=24, **kw)
    # same simulations and same null fits (same seeds): more starts can only find a larger maximum
    np.testing.assert_allclose(one.loglike_null, eight.loglike_null, rtol=1e-12)
    # (start 0, the nested one, is common to all three runs; the random starts of the 24-start run are other draws than
    #  those of the 8-start run, so "more starts" is only better on average there)
    assert np.all(eight.T >= one.T - 1e-6) and np.mean(many.T >= eight.T - 1e-3) > 0.7
    m1, m8, m24 = np.median(one.T), np.median(eight.T), np.median(many.T)
    print("median T: 1 start %.2f, 8 starts %.2f, 24 starts %.2f" % (m1, m8, m24))
    # a single local fit from the nested start: little gained over the null model (the reference's notebook, whose fits are
    # followed by only a short MCMC, shows a median of ~4 for these 3 extra parameters); the maximum over the frequency range
    # (7 e-folds of log_omega0, i.e. many independent frequencies) is higher, and saturates once the range is covered
    assert m1 < m8 - 1.0 and m24 > m8 - 0.3, (m1, m8, m24)
    assert 0.0 <= m1 < 6.0 and 3.0 < m8 < 12.0 and m24 < 16.0
