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


@pytest.mark.parametrize("kernel_flag", ["k8", "scan"])
@pytest.mark.parametrize("W", [12, 11])
def test_chain_on_injected_draws_equals_the_emcee_restatement(kernel_flag, W):
    """Sampler parity proper: K8 and K10 fed the draws emcee 3.1.4 would take from its RandomState (split, stretch
    variates, partner indices, accept uniforms -- oracle.sampler.draws_like_emcee) walk the SAME chain as the NumPy
    restatement of RedBlueMove / StretchMove driving the oracle's log-probability: positions bit for bit, stored
    log-probabilities to rounding.  Two ensembles on two lightcurves, even and odd ensemble sizes."""
    import ctypes
    from mind_the_gaps_b200 import _native
    from oracle import sampler as osamp
    N, E, nsteps = 120, 2, 60
    t, y0, dy0, _ = fake_lightcurve(N, 4)
    _, y1, dy1, _ = fake_lightcurve(N, 9)
    Y = np.ascontiguousarray(np.stack([y0, y1])); DY = np.stack([dy0, dy1])
    kernel = DampedRandomWalk(-6.0, -13.5, bounds=B_DRW) + Lorentzian(-7.0, 2.0, -12.5, bounds=dict(log_S0=(-14, 2), log_Q=(0.0, 6.0), log_omega0=(-16, -9)))
    gp = GP(kernel, mean=0.0)
    model = gp.device_model()
    nd = model.n_free
    bounds = gp.get_parameter_bounds()
    spec = [("drw",), ("lorentzian",)]
    means = np.ascontiguousarray(Y.mean(axis=1))
    rs = np.random.RandomState(123)
    p0 = gp.get_parameter_vector()[None, None, :] + rs.normal(0, 0.05, (E, W, nd))
    H = (W + 1) // 2
    draws = [osamp.draws_like_emcee(rs, W, nsteps) for _ in range(E)]
    split = np.ascontiguousarray(np.stack([d[0] for d in draws]), dtype=np.int32)
    uz = np.ascontiguousarray(np.stack([d[1] for d in draws])); rint = np.ascontiguousarray(np.stack([d[2] for d in draws]), dtype=np.int32)
    uacc = np.ascontiguousarray(np.stack([d[3] for d in draws]))
    assert uz.shape == (E, nsteps, 2, H)
    pos = np.ascontiguousarray(p0.copy()); lp = np.empty((E, W))
    chain = np.empty((E, nsteps, W, nd)); chain_lp = np.empty((E, nsteps, W)); nacc = np.zeros((E, W), dtype=np.int32)
    diag = np.ascontiguousarray((DY + 1e-12) ** 2)
    flag = {"k8": _native.FLAG_NO_SCAN, "scan": _native.FLAG_SCAN}[kernel_flag] | _native.FLAG_DS_MEAN
    _native.check(_native.lib().mtg_ensemble_injected_host(
        ctypes.byref(model), N, E, E, W, _native.ptr(np.ascontiguousarray(t)), _native.ptr(Y), _native.ptr(diag), _native.ptr(means),
        _native.ptr(pos), _native.ptr(lp), nsteps, 2.0, _native.ptr(split), _native.ptr(uz), _native.ptr(rint), _native.ptr(uacc),
        _native.ptr(chain), _native.ptr(chain_lp), _native.ptr(nacc), flag))
    for e in range(E):
        def logp(x, e=e):
            out = ogp.log_prior(x, bounds) + np.real(ogp.loglike_params(spec, x, t, Y[e], DY[e]))
            return np.where(np.isfinite(out), out, -np.inf)
        ref_chain, ref_lp, ref_acc = osamp.stretch_move_run(logp, p0[e], nsteps, *draws[e])
        np.testing.assert_array_equal(chain[e], ref_chain)
        np.testing.assert_allclose(chain_lp[e], ref_lp, rtol=1e-11)
        np.testing.assert_array_equal(nacc[e], ref_acc)
        np.testing.assert_array_equal(pos[e], ref_chain[-1])
        assert 0.1 < ref_acc.mean() / nsteps < 0.9
