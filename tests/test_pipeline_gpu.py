"""GPU: the user-level workflow end to end -- GPModelling.fit / derive_posteriors /
generate_from_posteriors / standarized_residuals and the batched LRT driver."""
import warnings

import numpy as np
import pytest

from mind_the_gaps_b200 import lrt
from mind_the_gaps_b200.celerite import GP, terms
from mind_the_gaps_b200.gpmodelling import GPModelling
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from oracle import gp as ogp

from helpers import window

pytestmark = pytest.mark.gpu
DAY = 86400.0


def make_lc(N=200, seed=1):
    """Config C1: 200-point irregular window, DRW truth, Poisson noise."""
    t, exposures = window(N, seed)
    mean = 0.5
    var = (0.1 * mean) ** 2
    w = 2 * np.pi / (20 * DAY)
    rng = np.random.default_rng(seed)
    y = np.empty(N); y[0] = rng.normal(0, np.sqrt(var))
    for n in range(1, N):
        rho = np.exp(-w * (t[n] - t[n - 1]))
        y[n] = rho * y[n - 1] + rng.normal(0, np.sqrt(var * (1 - rho * rho)))
    rate = mean + y
    counts = rng.poisson(rate * exposures)
    return GappyLightcurve(t, counts / exposures, np.sqrt(counts) / exposures, exposures), np.array([np.log(var), np.log(w)])


def test_predict_and_standardised_residuals_match_dense_oracle():
    lc, truth = make_lc(150, 2)
    kernel = DampedRandomWalk(truth[0], truth[1]) + terms.JitterTerm(np.log(0.01))
    gpm = GPModelling(lc, kernel)
    mu, var = gpm.gp.predict(lc.y, return_var=True, return_cov=False)
    from oracle import terms as oterms
    spec = [("drw",), ("jitter",)]
    co = oterms.coefficients(spec, gpm.gp.get_parameter_vector())
    diag = (lc.dy + 1e-12) ** 2
    rmu, rvar = ogp.predict_dense(lc.times, lc.y - lc.mean, diag, *co[:6], jitter=co[6])
    np.testing.assert_allclose(mu, lc.mean + rmu, rtol=1e-9)
    np.testing.assert_allclose(var, rvar, rtol=1e-7, atol=1e-12)
    res = gpm.standarized_residuals()
    np.testing.assert_allclose(res, (lc.y - (lc.mean + rmu)) / np.sqrt(rvar + co[6]), rtol=1e-6)


def test_gpmodelling_fit_and_posteriors_c1():
    """configs[0]: null-model log-likelihood + fit + emcee-style sampling on one 200-point lightcurve."""
    np.random.seed(3)
    lc, truth = make_lc(200, 1)
    bounds = dict(log_S0=(-15, 5), log_omega0=(-20, -5))
    gpm = GPModelling(lc, DampedRandomWalk(truth[0], truth[1], bounds=bounds))
    sol = gpm.fit()
    assert sol.success and np.all(np.abs(sol.x - truth) < 1.5)
    ref = ogp.loglike_params([("drw",)], sol.x[None, :], lc.times, lc.y, lc.dy)[0]
    assert abs(-sol.fun - ref) < 1e-9 * abs(ref)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gpm.derive_posteriors(max_steps=1500, convergence_steps=500, walkers=16, cores=1, progress=False)
    assert gpm.mcmc_samples.shape[1] == 2 and len(gpm.loglikelihoods) == len(gpm.mcmc_samples)
    assert gpm.max_loglikelihood <= -sol.fun + 1e-6 and gpm.max_loglikelihood > -sol.fun - 3.0
    assert np.all(np.abs(gpm.median_parameters - sol.x) < 1.0)
    assert gpm.tau.shape == (2,) and gpm.get_rstat().shape[1] == 2
    # every stored log-probability is the oracle's log-likelihood at that sample (flat prior inside the box)
    idx = np.random.default_rng(0).choice(len(gpm.mcmc_samples), 20)
    ref = ogp.loglike_params([("drw",)], gpm.mcmc_samples[idx], lc.times, lc.y, lc.dy)
    np.testing.assert_allclose(gpm.loglikelihoods[idx], ref, rtol=1e-10)
    lcs = gpm.generate_from_posteriors(nsims=5, cpus=2)
    assert len(lcs) == 5 and all(len(l.y) == lc.n and np.all(l.dy > 0) for l in lcs)
    assert abs(np.mean([l.mean for l in lcs]) - lc.mean) < 0.1


def test_generate_from_posteriors_uses_each_drawn_sample():
    """a9 (gpmodelling.py:478-539): every simulated lightcurve is generated from ITS OWN posterior sample.  Two tight clusters
    of posterior samples whose variances differ by e^2 and bend frequencies by e^2: refitting the simulated lightcurves
    (batched device L-BFGS) must recover, cluster by cluster, the parameters each lightcurve was drawn with, and the
    lightcurves must carry the data's window, mean and a positive error column."""
    from mind_the_gaps_b200 import fitting
    lc, truth = make_lc(400, 7)
    bounds = dict(log_S0=(-15, 5), log_omega0=(-20, -5))
    gpm = GPModelling(lc, DampedRandomWalk(truth[0], truth[1], bounds=bounds))
    rng = np.random.default_rng(1)
    centres = np.array([[truth[0] + 1.0, truth[1] + 1.0], [truth[0] - 1.0, truth[1] - 1.0]])
    samples = np.concatenate([c + rng.normal(0, 0.01, (50, 2)) for c in centres])
    gpm._mcmc_samples = samples
    gpm._loglikelihoods = np.zeros(len(samples))
    nsims = 300
    np.random.seed(12)
    idx = np.random.randint(len(samples), size=nsims)            # the draw generate_from_posteriors makes first
    np.random.seed(12)
    lcs = gpm.generate_from_posteriors(nsims=nsims)
    assert len(lcs) == nsims
    Y = np.array([l.y for l in lcs]); DY = np.array([l.dy for l in lcs])
    assert all(np.array_equal(l.times, lc.times) for l in lcs[:5]) and np.all(DY > 0) and np.all(np.isfinite(Y))
    assert abs(Y.mean() - lc.mean) < 0.02 * lc.mean
    res = fitting.fit_lightcurves(gpm.gp, lc.times, Y, DY, x0=samples[idx])
    xhat = np.array([r.x for r in res])
    for k, c in enumerate(centres):
        sel = idx // 50 == k
        assert sel.sum() > 100
        med = np.median(xhat[sel], axis=0)
        assert np.all(np.abs(med - c) < 0.35), (k, med, c)
    # the two clusters are told apart lightcurve by lightcurve in variance (2 in log) for the large majority
    assert np.mean((xhat[:, 0] > truth[0]) == (idx // 50 == 0)) > 0.9


def test_lrt_driver_null_distribution():
    lc, truth = make_lc(200, 4)
    b_drw = dict(log_S0=(-15, 5), log_omega0=(-20, -5))
    null_kernel = DampedRandomWalk(truth[0], truth[1], bounds=b_drw)
    alt_kernel = Lorentzian(truth[0] - 2, np.log(20.0), np.log(2 * np.pi / (10 * DAY)),
                            bounds=dict(log_S0=(-20, 5), log_Q=(0.5, 7), log_omega0=(-16, -9))) + \
        DampedRandomWalk(truth[0], truth[1], bounds=b_drw)
    rng = np.random.default_rng(5)
    samples = truth[None, :] + rng.normal(0, 0.15, (96, 2))
    res = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, n_starts_null=2, n_starts_alt=6, seed=11,
                               return_lightcurves=True)
    assert res.T.shape == (96,) and np.all(np.isfinite(res.T))
    assert np.all(res.T > -1e-4)                      # nested start: the alternative never fits worse
    # level pinned in tests/test_lrt_null_gpu.py (CPU refits from the same starts; look-elsewhere effect over log_omega0)
    assert 3.0 < np.median(res.T) < 11.0 and np.percentile(res.T, 95) < 30.0
    # the recorded likelihoods are the oracle's at the recorded optima, on the simulated data
    Y, DY = res.lightcurves
    for i in (0, 17, 95):
        l0 = ogp.loglike_params([("drw",)], res.params_null[i][None, :], lc.times, Y[i], DY[i])[0]
        l1 = ogp.loglike_params([("lorentzian",), ("drw",)], res.params_alt[i][None, :], lc.times, Y[i], DY[i])[0]
        assert abs(l0 - res.loglike_null[i]) < 1e-9 * abs(l0) and abs(l1 - res.loglike_alt[i]) < 1e-9 * abs(l1)
    assert 0.0 <= res.pvalue(5.0) <= 1.0
    # same seed -> same simulations and same statistics
    res2 = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, n_starts_null=2, n_starts_alt=6, seed=11)
    np.testing.assert_allclose(res2.loglike_null, res.loglike_null, rtol=1e-12)


def test_lrt_with_reference_style_mcmc_refinement():
    """tutorial_ppp.ipynb cells 11-13: each fit followed by a short ensemble run whose best log-likelihood is kept"""
    lc, truth = make_lc(200, 4)
    b_drw = dict(log_S0=(-15, 5), log_omega0=(-20, -5))
    null_kernel = DampedRandomWalk(truth[0], truth[1], bounds=b_drw)
    alt_kernel = Lorentzian(truth[0] - 2, np.log(20.0), np.log(2 * np.pi / (10 * DAY)),
                            bounds=dict(log_S0=(-20, 5), log_Q=(0.5, 7), log_omega0=(-16, -9))) + \
        DampedRandomWalk(truth[0], truth[1], bounds=b_drw)
    samples = truth[None, :] + np.random.default_rng(5).normal(0, 0.15, (40, 2))
    kw = dict(n_starts_null=2, n_starts_alt=6, seed=11, return_lightcurves=True)
    plain = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, **kw)
    ref = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, refine_walkers=30, refine_steps=60, **kw)
    # same simulations; the refined likelihoods can only improve on the optimiser's
    np.testing.assert_array_equal(ref.lightcurves[0], plain.lightcurves[0])
    assert np.all(ref.loglike_null >= plain.loglike_null - 1e-9)
    # (the alternative's nested start moves with the refined null optimum, so its fits are not identical run to run)
    assert np.mean(ref.loglike_alt - plain.loglike_alt) > -0.05 and np.mean(ref.loglike_alt >= plain.loglike_alt - 1e-3) > 0.8
    Y, DY = ref.lightcurves
    for i in (0, 13, 39):
        l0 = ogp.loglike_params([("drw",)], ref.params_null[i][None, :], lc.times, Y[i], DY[i])[0]
        l1 = ogp.loglike_params([("lorentzian",), ("drw",)], ref.params_alt[i][None, :], lc.times, Y[i], DY[i])[0]
        assert abs(l0 - ref.loglike_null[i]) < 1e-9 * abs(l0) and abs(l1 - ref.loglike_alt[i]) < 1e-9 * abs(l1)
    assert np.all(np.isfinite(ref.T)) and abs(np.median(ref.T) - np.median(plain.T)) < 3.0


def test_tutorial_ppp_example_runs_end_to_end():
    """examples/tutorial_ppp.py: the reference's notebook, cell by cell, at reduced size -- without and with a QPO in the
    input lightcurve.  No QPO: an unremarkable p-value; a coherent QPO: T_obs far in the tail of the null distribution."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "tutorial_ppp.py")
    spec = importlib.util.spec_from_file_location("tutorial_ppp", path)
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    out = mod.main(["--nsims", "200", "--max-steps", "3000", "--loop", "2", "--points", "400"])
    assert len(out["T_dist"]) == 200 and np.all(np.isfinite(out["T_dist"])) and np.all(out["T_dist"] >= -1e-6)
    assert np.all(np.isfinite(out["T_loop"])) and len(out["T_loop"]) == 2
    assert out["p"] > 0.001
    # T of the null simulations follows roughly a chi^2 with 3 extra parameters (median ~ 4, 95 % ~ 10-14)
    assert 0.5 < np.median(out["T_dist"]) < 9 and 4 < np.percentile(out["T_dist"], 95) < 25
    out_q = mod.main(["--nsims", "200", "--max-steps", "3000", "--loop", "0", "--points", "400", "--qpo"])
    assert out_q["T_obs"] > np.percentile(out_q["T_dist"], 99) and out_q["p"] < 0.02


def test_lrt_with_the_tutorial_realterm_complexterm_kernels_is_nested():
    """tutorial_ppp.ipynb cells 7 / 9: null = RealTerm, alternative = ComplexTerm + RealTerm.  The stripped names of the
    ComplexTerm's first parameters equal the RealTerm's; the nested start must still land in the RealTerm's slots."""
    lc, truth = make_lc(200, 4)
    null_kernel = terms.RealTerm(truth[0], truth[1], bounds=dict(log_a=(-15, 5), log_c=(-20, -5)))
    alt_kernel = terms.ComplexTerm(truth[0] - 2, np.log(2 * np.pi / (10 * DAY) / 40.0), np.log(2 * np.pi / (10 * DAY)),
                                   bounds=dict(log_a=(-20, 5), log_c=(-25, -8), log_d=(-16, -9))) + \
        terms.RealTerm(truth[0], truth[1], bounds=dict(log_a=(-15, 5), log_c=(-20, -5)))
    assert lrt.infer_null_slots(null_kernel, alt_kernel) == [3, 4]
    samples = truth[None, :] + np.random.default_rng(5).normal(0, 0.15, (64, 2))
    res = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, n_starts_null=2, n_starts_alt=6, seed=11)
    assert np.all(np.isfinite(res.T)) and np.all(res.T > -1e-4)
    assert res.n_used == 64 and res.n_invalid == 0


def test_lrt_flags_simulations_with_negative_expected_counts():
    """Gaussian-PDF simulation of a faint, highly variable source: some bins get rate * exposure < 0, where the reference's
    np.random.poisson raises ValueError('lam < 0') (noise_models.py:72).  The device pipeline flags those simulations."""
    lc, truth = make_lc(200, 4)
    faint = GappyLightcurve(lc.times, lc.y * 0.2, lc.dy, lc.exposures)          # mean 0.1, sigma of the process 0.035: ~3 sigma from zero
    big = np.array([np.log(0.035 ** 2), truth[1]])
    b_drw = dict(log_S0=(-15, 5), log_omega0=(-20, -5))
    null_kernel = DampedRandomWalk(big[0], big[1], bounds=b_drw)
    alt_kernel = Lorentzian(big[0] - 2, np.log(20.0), np.log(2 * np.pi / (10 * DAY)),
                            bounds=dict(log_S0=(-20, 5), log_Q=(0.5, 7), log_omega0=(-16, -9))) + DampedRandomWalk(big[0], big[1], bounds=b_drw)
    samples = np.tile(big, (48, 1))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        res = lrt.simulate_and_fit(faint, null_kernel, alt_kernel, samples, n_starts_null=1, n_starts_alt=2, seed=2, max_iter=30)
    assert res.n_invalid > 0 and res.n_used + res.n_invalid == 48
    assert any("negative expected counts" in str(x.message) for x in w)
    bad = ~np.isfinite(res.T)
    assert np.all(res.status_null[bad] == 3) and np.all(res.status_alt[bad] == 3)
    p, n = res.pvalue(3.0, return_n=True)
    assert n == res.n_used and 0.0 <= p <= 1.0


def test_complexterm_prior_is_applied_on_the_device():
    """celerite ComplexTerm with a fitted log_b: log_prior = -inf when a c < b d.  The batched / sampler
    log-probability must agree with GPModelling._log_probability row by row."""
    lc, truth = make_lc(120, 6)
    k = terms.ComplexTerm(truth[0], truth[0] - 1.0, truth[1], truth[1] + 1.0,
                          bounds=dict(log_a=(-15, 5), log_b=(-15, 5), log_c=(-20, -5), log_d=(-20, -5))) + \
        DampedRandomWalk(truth[0], truth[1], bounds=dict(log_S0=(-15, 5), log_omega0=(-20, -5)))
    gpm = GPModelling(lc, k)
    p0 = gpm.gp.get_parameter_vector()
    rng = np.random.default_rng(1)
    P = p0[None, :] + rng.normal(0, 1.5, (64, len(p0)))
    batch = gpm._log_probability(P)
    single = np.array([gpm._log_probability(row) for row in P])
    viol = (P[:, 0] + P[:, 2]) < (P[:, 1] + P[:, 3])
    assert viol.any() and (~viol).any()
    assert np.all(np.isneginf(batch[viol]))
    np.testing.assert_allclose(batch[~viol & np.isfinite(single)], single[~viol & np.isfinite(single)], rtol=1e-10)
    assert np.array_equal(np.isfinite(batch), np.isfinite(single))


def test_user_log_prior_override_is_evaluated_on_the_host():
    lc, truth = make_lc(120, 6)

    class Picky(DampedRandomWalk):              # inherits the device coefficient map, overrides the prior
        def log_prior(self):
            return -np.inf if self.log_S0 > -5.0 else super(Picky, self).log_prior()

    gpm = GPModelling(lc, Picky(truth[0], truth[1], bounds=dict(log_S0=(-15, 5), log_omega0=(-20, -5))))
    assert not gpm.gp.device_prior_supported()
    P = np.array([[truth[0], truth[1]], [-4.0, truth[1]], [-6.0, truth[1]]])
    out = gpm._log_probability(P)
    assert np.isfinite(out[0]) == (truth[0] <= -5.0) and np.isneginf(out[1]) and np.isfinite(out[2])
    np.testing.assert_allclose(out[2], gpm._log_probability(P[2]), rtol=1e-12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(0)
        gpm.derive_posteriors(fit=False, max_steps=60, convergence_steps=30, walkers=8, progress=False)
    assert np.all(gpm.mcmc_samples[:, 0] <= -5.0)
