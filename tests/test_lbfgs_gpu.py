"""GPU: batched device L-BFGS (K4) vs scipy's L-BFGS-B run on the oracle, as GPModelling.fit does."""
import numpy as np
import pytest
from scipy.optimize import minimize

from mind_the_gaps_b200 import fitting
from mind_the_gaps_b200.celerite import GP
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from oracle import gp as ogp

from helpers import fake_lightcurve

pytestmark = pytest.mark.gpu

DAY = 86400.0


def simulate_drw(t, var, w, rng, sigma):
    """Exact DRW draw (AR(1) on irregular times) + white noise."""
    y = np.empty(len(t))
    y[0] = rng.normal(0, np.sqrt(var))
    for n in range(1, len(t)):
        rho = np.exp(-w * (t[n] - t[n - 1]))
        y[n] = rho * y[n - 1] + rng.normal(0, np.sqrt(var * (1 - rho * rho)))
    return 0.5 + y + rng.normal(0, sigma, len(t))


def make_set(S, N, seed):
    t, _, _, _ = fake_lightcurve(N, seed)
    rng = np.random.default_rng(seed)
    var, w = 0.05 ** 2, 2 * np.pi / (20 * DAY)
    Y = np.stack([simulate_drw(t, var, w, rng, 0.01) for _ in range(S)])
    DY = np.full((S, N), 0.01)
    return t, Y, DY, np.array([np.log(var), np.log(w)])


def test_null_model_fits_match_scipy():
    S, N = 12, 300
    t, Y, DY, truth = make_set(S, N, 41)
    bounds = [(-15.0, 5.0), (-20.0, -5.0)]
    gp = GP(DampedRandomWalk(truth[0] + 0.5, truth[1] - 0.5, bounds=bounds), mean=0.0)
    res = fitting.fit_lightcurves(gp, t, Y, DY)
    for s in range(S):
        f = lambda p: -ogp.loglike_params([("drw",)], p[None, :], t, Y[s], DY[s])[0]
        ref = minimize(f, gp.get_parameter_vector(), method="L-BFGS-B", bounds=bounds)
        assert res[s].success
        assert res[s].fun <= ref.fun + 1e-5 * abs(ref.fun) + 1e-6
        # the optimum itself is reproduced by the oracle at the returned point
        np.testing.assert_allclose(f(res[s].x), res[s].fun, rtol=1e-10)
        np.testing.assert_allclose(res[s].x, ref.x, atol=5e-3)


def test_alternative_model_multistart_not_worse_than_scipy():
    S, N = 6, 300
    t, Y, DY, truth = make_set(S, N, 42)
    bounds = [(-15.0, 5.0), (-20.0, -5.0), (-15.0, 5.0), (-2.0, 8.0), (-20.0, -5.0)]
    x0 = np.array([truth[0], truth[1], truth[0] - 2.0, np.log(10.0), np.log(2 * np.pi / (10 * DAY))])
    kernel = DampedRandomWalk(x0[0], x0[1], bounds=bounds[:2]) + Lorentzian(x0[2], x0[3], x0[4], bounds=bounds[2:])
    gp = GP(kernel, mean=0.0)
    res = fitting.fit_lightcurves(gp, t, Y, DY, n_starts=8, rng=np.random.default_rng(0), spread=[0.5, 0.5, 1.0, 1.0, 0.5])
    spec = [("drw",), ("lorentzian",)]
    lo = np.array([b[0] for b in bounds]); hi = np.array([b[1] for b in bounds])
    n_match = 0
    for s in range(S):
        f = lambda p: -ogp.loglike_params(spec, p[None, :], t, Y[s], DY[s])[0]
        g = lambda p: -ogp.grad_complex_step(spec, p[None, :], t, Y[s], DY[s])[0]
        ref = minimize(f, x0, method="L-BFGS-B", bounds=bounds)
        n_match += res[s].fun <= ref.fun + 1e-4 * abs(ref.fun) + 1e-5
        assert np.all(res[s].x >= lo) and np.all(res[s].x <= hi)
        np.testing.assert_allclose(f(res[s].x), res[s].fun, rtol=1e-10)
        # the likelihood is multimodal in the QPO parameters, so a different start can end in a
        # different basin; but every returned point must be a genuine local optimum: scipy restarted
        # there (with the exact gradient) cannot improve it
        cont = minimize(f, res[s].x, jac=g, method="L-BFGS-B", bounds=bounds)
        assert res[s].fun - cont.fun < 1e-3, (s, res[s].fun, cont.fun)
    assert n_match >= S - 2, "device L-BFGS reached scipy's optimum for only %d of %d lightcurves" % (n_match, S)


def test_bounds_are_respected_when_optimum_is_outside():
    S, N = 4, 200
    t, Y, DY, truth = make_set(S, N, 43)
    bounds = [(truth[0] + 1.0, 5.0), (-20.0, -5.0)]           # true variance excluded by the lower bound
    gp = GP(DampedRandomWalk(truth[0] + 2.0, truth[1], bounds=bounds), mean=0.0)
    res = fitting.fit_lightcurves(gp, t, Y, DY)
    for s in range(S):
        assert abs(res[s].x[0] - bounds[0][0]) < 1e-12
        f = lambda p: -ogp.loglike_params([("drw",)], p[None, :], t, Y[s], DY[s])[0]
        ref = minimize(f, gp.get_parameter_vector(), method="L-BFGS-B", bounds=bounds)
        assert res[s].fun <= ref.fun + 1e-5 * abs(ref.fun) + 1e-6


def test_scipy_driven_fit_follows_scipy_on_the_oracle():
    """GPModelling.fit default path: scipy's L-BFGS-B with 2-point differences (the reference's own call), the ndim + 1
    evaluations of each request batched in one launch -> the same iterates as scipy differentiating the oracle itself."""
    from mind_the_gaps_b200.gpmodelling import GPModelling
    from mind_the_gaps_b200.lightcurves import GappyLightcurve
    S, N = 3, 300
    t, Y, DY, truth = make_set(S, N, 43)
    bounds = [(-15.0, 5.0), (-20.0, -5.0)]
    for s in range(S):
        lc = GappyLightcurve(t, Y[s], DY[s])
        gpm = GPModelling(lc, DampedRandomWalk(truth[0] + 0.5, truth[1] - 0.5, bounds=bounds))
        sol = gpm.fit(optimizer="scipy")
        f = lambda p: -ogp.loglike_params([("drw",)], p[None, :], t, Y[s], DY[s])[0]
        ref = minimize(f, gpm.initial_params, method="L-BFGS-B", bounds=bounds)
        assert sol.success == ref.success
        assert abs(sol.fun - ref.fun) <= 1e-7 * abs(ref.fun)
        np.testing.assert_allclose(sol.x, ref.x, atol=2e-4)
        np.testing.assert_allclose(gpm.gp.get_parameter_vector(), sol.x)
        # the default: the device optimiser (one launch of the persistent L-BFGS kernel), same optimum
        dev = gpm.fit(gpm.initial_params)
        assert dev.success and dev.fun <= ref.fun + 1e-6 * abs(ref.fun)
        np.testing.assert_allclose(dev.x, ref.x, atol=5e-3)
        np.testing.assert_allclose(gpm.gp.get_parameter_vector(), dev.x)
        np.testing.assert_allclose(f(dev.x), dev.fun, rtol=1e-10)
    # a start on the upper bound differentiates backwards and stays inside the box
    gpm = GPModelling(GappyLightcurve(t, Y[0], DY[0]), DampedRandomWalk(5.0, -5.0, bounds=bounds))
    sol = gpm.fit(optimizer="scipy")
    assert np.all(sol.x >= [-15.0, -20.0]) and np.all(sol.x <= [5.0, -5.0]) and np.isfinite(sol.fun)


@pytest.mark.parametrize("switch", ["0", "4096"])
def test_persistent_tail_reaches_the_same_optima_as_the_round_driver(monkeypatch, switch):
    """mtg_lbfgs_* with the persistent kernel (gradient K11, whole L-BFGS loop of a problem inside one CTA) and with the
    round driver only (K2 + update kernel per round, MTG_LBFGS_SWITCH=0): the same state machine, so the same optima up
    to the rounding of the two gradient kernels; every optimum is verified against the oracle."""
    import importlib
    S, N = 24, 400
    t, Y, DY, truth = make_set(S, N, 44)
    bounds = [(-15.0, 5.0), (-20.0, -5.0), (-15.0, 5.0), (-2.0, 8.0), (-20.0, -5.0)]
    x0 = np.array([truth[0], truth[1], truth[0] - 2.0, np.log(10.0), np.log(2 * np.pi / (10 * DAY))])
    kernel = DampedRandomWalk(x0[0], x0[1], bounds=bounds[:2]) + Lorentzian(x0[2], x0[3], x0[4], bounds=bounds[2:])
    gp = GP(kernel, mean=0.0)
    # the switch count is read once per process: this test runs the library in a child process per setting
    import subprocess, sys, os, json
    code = (
        "import sys, json, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from test_lbfgs_gpu import make_set, DAY\n"
        "from mind_the_gaps_b200 import fitting\n"
        "from mind_the_gaps_b200.celerite import GP\n"
        "from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian\n"
        "t, Y, DY, truth = make_set(24, 400, 44)\n"
        "bounds = [(-15.0, 5.0), (-20.0, -5.0), (-15.0, 5.0), (-2.0, 8.0), (-20.0, -5.0)]\n"
        "x0 = np.array([truth[0], truth[1], truth[0] - 2.0, np.log(10.0), np.log(2 * np.pi / (10 * DAY))])\n"
        "k = DampedRandomWalk(x0[0], x0[1], bounds=bounds[:2]) + Lorentzian(x0[2], x0[3], x0[4], bounds=bounds[2:])\n"
        "res = fitting.fit_lightcurves(GP(k, mean=0.0), t, Y, DY, n_starts=4, rng=np.random.default_rng(0), spread=[0.5, 0.5, 1.0, 1.0, 0.5])\n"
        "print(json.dumps({'fun': [r.fun for r in res], 'x': [list(r.x) for r in res], 'ok': [bool(r.success) for r in res]}))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MTG_LBFGS_SWITCH=switch)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    got = json.loads(out.stdout.strip().splitlines()[-1])
    spec = [("drw",), ("lorentzian",)]
    fun = np.array(got["fun"]); X = np.array(got["x"])
    assert np.mean(got["ok"]) > 0.9
    for s in range(S):
        ref = -ogp.loglike_params(spec, X[s][None, :], t, Y[s], DY[s])[0]
        np.testing.assert_allclose(fun[s], ref, rtol=1e-10)
    # both settings find the same best optimum for (nearly) every lightcurve: compare with an in-process run (default switch)
    res = fitting.fit_lightcurves(gp, t, Y, DY, n_starts=4, rng=np.random.default_rng(0), spread=[0.5, 0.5, 1.0, 1.0, 0.5])
    here = np.array([r.fun for r in res])
    assert np.mean(np.abs(here - fun) < 1e-4 * np.abs(here)) > 0.9
