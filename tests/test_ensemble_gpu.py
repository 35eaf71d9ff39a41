"""GPU: the device-resident ensemble sampler (K8) -- its in-kernel log-probability equals K1 / the oracle, the chain is
reproducible and independent of how the run is cut into launches, and the sampled posterior agrees with direct
numerical integration of the posterior on a grid and with the host stretch-move loop (statistical parity with
emcee's algorithm, gpmodelling.py:247-264)."""
import numpy as np
import pytest

from mind_the_gaps_b200.celerite import GP
from mind_the_gaps_b200.gpmodelling import GPModelling
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from mind_the_gaps_b200.sampler import DeviceEnsembleSampler, EnsembleSampler
from oracle import gp as ogp

from helpers import fake_lightcurve

pytestmark = pytest.mark.gpu

B_DRW = dict(log_S0=(-12.0, 2.0), log_omega0=(-18.0, -8.0))


def _drw_problem(n=150, seed=4):
    t, y, dy, ex = fake_lightcurve(n, seed)
    kernel = DampedRandomWalk(-6.0, -13.5, bounds=B_DRW)
    gp = GP(kernel, mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    return t, y, dy, ex, gp


def _start(gp, W, rng, scale=0.05):
    p0 = gp.get_parameter_vector()
    return p0[None, :] + rng.normal(0, scale, (W, len(p0)))


def test_chain_log_prob_equals_k1_and_oracle():
    t, y, dy, ex, gp = _drw_problem()
    W = 16
    s = DeviceEnsembleSampler(W, gp, y, seed=11, block=40, kernel="k8")
    s.run_mcmc(_start(gp, W, np.random.default_rng(0)), 80)
    chain = s.get_chain(); lp = s.get_log_prob()
    assert chain.shape == (80, W, 2) and lp.shape == (80, W)
    flat = chain.reshape(-1, 2)
    k1 = gp.log_probability_batch(y, flat, kernel="k1")
    np.testing.assert_allclose(lp.reshape(-1), k1, rtol=1e-13)
    ref = np.real(ogp.loglike_params([("drw",)], flat[::37], t, y, dy))
    np.testing.assert_allclose(lp.reshape(-1)[::37], ref, rtol=1e-10)
    lo = np.array([B_DRW["log_S0"][0], B_DRW["log_omega0"][0]]); hi = np.array([B_DRW["log_S0"][1], B_DRW["log_omega0"][1]])
    assert np.all(flat >= lo) and np.all(flat <= hi)
    assert 0.15 < s.acceptance_fraction.mean() < 0.9


def test_reproducible_and_independent_of_block_size():
    t, y, dy, ex, gp = _drw_problem()
    W = 12
    p0 = _start(gp, W, np.random.default_rng(1))
    chains = []
    for block in (7, 50, 200):
        np.random.seed(5)
        s = DeviceEnsembleSampler(W, gp, y, seed=3, block=block, kernel="k8")
        s.run_mcmc(p0, 120)
        chains.append((s.get_chain().copy(), s.get_log_prob().copy()))
    for c, l in chains[1:]:
        assert np.array_equal(c, chains[0][0]) and np.array_equal(l, chains[0][1])


def test_posterior_matches_grid_integration_and_host_sampler():
    t, y, dy, ex, gp = _drw_problem()
    # posterior moments by brute-force integration on a grid (uniform prior inside the box)
    g0 = np.linspace(-12.0, 2.0, 281); g1 = np.linspace(-18.0, -8.0, 201)          # the whole prior box, step 0.05
    G = np.array([[a, b] for a in g0 for b in g1])
    lp = gp.log_probability_batch(y, G)
    w = np.exp(lp - lp.max()); w /= w.sum()
    mean_grid = (w[:, None] * G).sum(axis=0)
    std_grid = np.sqrt((w[:, None] * (G - mean_grid) ** 2).sum(axis=0))
    W, steps, burn = 32, 3000, 500
    dev = DeviceEnsembleSampler(W, gp, y, seed=2, block=500)
    dev.run_mcmc(_start(gp, W, np.random.default_rng(3), 0.2), steps)
    tau = dev.get_autocorr_time(tol=0)
    sam = dev.get_chain(discard=burn, flat=True)
    n_eff = sam.shape[0] / np.max(tau)
    assert n_eff > 1500
    tol = 5.0 * std_grid / np.sqrt(n_eff)
    assert np.all(np.abs(sam.mean(axis=0) - mean_grid) < tol + 0.01 * std_grid)
    assert np.all(np.abs(sam.std(axis=0) / std_grid - 1.0) < 0.1)
    # host loop (one K1 launch per half step): same target, same algorithm
    host = EnsembleSampler(W, 2, lambda p: gp.log_probability_batch(y, p), vectorize=True, seed=9)
    host.run_mcmc(_start(gp, W, np.random.default_rng(4), 0.2), 1500)
    hs = host.get_chain(discard=burn, flat=True)
    assert np.all(np.abs(hs.mean(axis=0) - sam.mean(axis=0)) < 0.15 * std_grid)
    assert abs(dev.acceptance_fraction.mean() - host.acceptance_fraction.mean()) < 0.05


def test_many_ensembles_each_on_its_own_lightcurve():
    E, N, W = 6, 120, 16
    ts = None; Y = []; DY = []
    for e in range(E):
        t, y, dy, ex = fake_lightcurve(N, 7)          # same window, different realisations
        rng = np.random.default_rng(100 + e)
        Y.append(y + rng.normal(0, 0.02, N)); DY.append(dy)
        ts = t
    Y = np.array(Y); DY = np.array(DY)
    kernel = DampedRandomWalk(-6.0, -13.5, bounds=B_DRW) + Lorentzian(-8.0, 3.0, -12.0, bounds=dict(log_S0=(-14, 0), log_Q=(0.5, 7), log_omega0=(-15, -9)))
    gp = GP(kernel, mean=0.0)
    gp.compute(ts, DY[0] + 1e-12)
    s = DeviceEnsembleSampler(W, gp, Y, yerr=DY + 1e-12, seed=8, block=30, frozen_mean=Y.mean(axis=1))
    p0 = np.array([_start(gp, W, np.random.default_rng(e)) for e in range(E)])
    s.run_mcmc(p0, 60)
    chain = s.get_chain(); lp = s.get_log_prob()
    assert chain.shape == (E, 60, W, 5) and lp.shape == (E, 60, W)
    for e in (0, 3, 5):
        ref = np.real(ogp.loglike_params([("drw",), ("lorentzian",)], chain[e, -1], ts, Y[e], DY[e]))
        np.testing.assert_allclose(lp[e, -1], ref, rtol=1e-10)
    assert s.get_chain(flat=True).shape == (E, 60 * W, 5)


def test_gpmodelling_derive_posteriors_uses_the_device_sampler():
    t, y, dy, ex = fake_lightcurve(100, 2)
    lc = GappyLightcurve(t, y, dy, ex)
    gpm = GPModelling(lc, DampedRandomWalk(-6.0, -13.5, bounds=B_DRW))
    np.random.seed(1)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gpm.derive_posteriors(fit=True, max_steps=600, convergence_steps=200, walkers=12, progress=False)
    assert isinstance(gpm.sampler, DeviceEnsembleSampler)
    assert gpm.mcmc_samples.shape[1] == 2 and np.isfinite(gpm.max_loglikelihood)
    assert len(gpm.autocorr) == 3
    # the maximum of the chain cannot be far below the MLE and never above it by more than rounding
    fit = gpm.fit(gpm.median_parameters)
    assert gpm.max_loglikelihood <= -fit.fun + 1e-6 and gpm.max_loglikelihood > -fit.fun - 5.0
