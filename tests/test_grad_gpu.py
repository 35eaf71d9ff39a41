"""GPU parity: K2 (fused forward + reverse-mode gradient) vs complex-step differentiation of the
oracle recursion.  north_star tolerance: 1e-9 relative; asserted 1e-9 relative to the
largest gradient component of each series (components near zero are compared on that scale)."""
import ctypes

import numpy as np
import pytest

from mind_the_gaps_b200 import _native
from mind_the_gaps_b200.celerite import GP, terms
from mind_the_gaps_b200.celerite.modeling import ConstantModel
from mind_the_gaps_b200.models import DampedRandomWalk, GaussianModel, LinearModel
from oracle import gp as ogp

from helpers import c2_truth, fake_lightcurve
from test_loglike_gpu import KINDS, build

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def assert_grad_close(g, ref, rtol=RTOL):
    scale = np.max(np.abs(ref), axis=1, keepdims=True)
    err = np.abs(g - ref) / scale
    assert np.max(err) < rtol, "max scaled gradient error %g" % np.max(err)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("N", [1, 2, 300])
def test_gradient_matches_complex_step(kind, N):
    t, y, dy, _ = fake_lightcurve(max(N, 2), 31)
    t, y, dy = t[:N], y[:N], dy[:N]
    kernel, spec = build(kind)
    gp = GP(kernel, mean=ConstantModel(float(np.mean(y))), fit_mean=False)
    gp.compute(t, dy + 1e-12)
    p0 = gp.get_parameter_vector()
    p = p0[None, :] + np.random.default_rng(N + 1).normal(0, 0.05, (40, len(p0)))
    if kind == "sho_overdamped+jitter":
        p[:, 1] = np.log(0.2) + np.random.default_rng(1).normal(0, 0.05, 40)
    ll, g, status = gp.grad_log_likelihood_batch(y, p, return_status=True)
    assert np.all(status == 0)
    np.testing.assert_allclose(ll, ogp.loglike_params(spec, p, t, y, dy), rtol=1e-10)
    assert_grad_close(g, ogp.grad_complex_step(spec, p, t, y, dy))


def test_gradient_with_fitted_constant_mean():
    t, y, dy, _ = fake_lightcurve(400, 32)
    gp = GP(DampedRandomWalk(np.log(0.01), np.log(1e-6)), mean=ConstantModel(float(np.mean(y))), fit_mean=True)
    gp.compute(t, dy + 1e-12)
    p = np.tile(gp.get_parameter_vector(), (20, 1)) + np.random.default_rng(2).normal(0, 0.02, (20, 3))
    ll, g = gp.grad_log_likelihood_batch(y, p)
    assert_grad_close(g, ogp.grad_complex_step([("drw",)], p, t, y, dy, fit_mean=True))


def _fd_grad(gp, y, p, eps):
    g = np.empty_like(p)
    for k in range(p.shape[1]):
        pp = p.copy(); pp[:, k] += eps[k]
        pm = p.copy(); pm[:, k] -= eps[k]
        g[:, k] = (gp.log_likelihood_batch(y, pp) - gp.log_likelihood_batch(y, pm)) / (2 * eps[k])
    return g


@pytest.mark.parametrize("meankind", ["linear", "gaussian"])
def test_gradient_with_linear_and_gaussian_mean(meankind):
    """mean-model parameter gradients, checked against central differences of K1 itself."""
    t, y, dy, _ = fake_lightcurve(300, 33)
    if meankind == "linear":
        mean = LinearModel(1e-10, 0.5)
        eps = [1e-5, 1e-5, 1e-14, 1e-6]
    else:
        mean = GaussianModel(float(t[150]), float(t[-1] - t[0]) / 4, 1e6, 0.5)
        eps = [1e-5, 1e-5, 1e3, 1e3, 1.0, 1e-6]
    gp = GP(DampedRandomWalk(np.log(0.01), np.log(1e-6)), mean=mean, fit_mean=True)
    gp.compute(t, dy + 1e-12)
    p = gp.get_parameter_vector()[None, :].copy()
    ll, g = gp.grad_log_likelihood_batch(y, p)
    fd = _fd_grad(gp, y, p, eps)
    np.testing.assert_allclose(g, fd, rtol=2e-5, atol=1e-8 * np.abs(fd).max())


def test_gradient_per_dataset_layout():
    N, S, group = 257, 7, 9
    t, _, _, _ = fake_lightcurve(N, 34)
    rng = np.random.default_rng(5)
    Y = 0.5 + 0.05 * rng.normal(size=(S, N)); DY = rng.uniform(0.01, 0.03, (S, N))
    kernel, spec = build("drw+lor")
    model = GP(kernel, mean=0.0).device_model()
    B = S * group
    p = c2_truth()[None, :] + rng.normal(0, 0.1, (B, 5))
    ll = np.empty(B); g = np.empty((B, 5)); st = np.empty(B, dtype=np.int32)
    diag = np.ascontiguousarray((DY + 1e-12) ** 2); means = np.ascontiguousarray(Y.mean(axis=1))
    _native.check(_native.lib().mtg_loglike_grad_host(
        ctypes.byref(model), N, S, group, B, _native.ptr(np.ascontiguousarray(t)), _native.ptr(Y), _native.ptr(diag),
        _native.ptr(means), _native.ptr(p), _native.ptr(ll), _native.ptr(g), _native.ptr(st), _native.FLAG_DS_MEAN))
    ref = np.concatenate([ogp.grad_complex_step(spec, p[s * group:(s + 1) * group], t, Y[s], DY[s]) for s in range(S)])
    assert_grad_close(g, ref)


def test_gradient_is_zero_outside_prior_or_not_pd():
    t, y, dy, _ = fake_lightcurve(100, 35)
    kernel, _ = build("drw+lor")
    gp = GP(kernel, mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    p = np.tile(gp.get_parameter_vector(), (2, 1)); p[1, 0] = 31.0
    ll, g, st = gp.grad_log_likelihood_batch(y, p, apply_prior=True, return_status=True)
    assert st[1] == 3 and ll[1] == -np.inf and np.all(g[1] == 0.0) and np.all(np.isfinite(g[0]))


def test_chunked_tape_gives_identical_results(monkeypatch):
    """A tiny tape budget forces several chunks; results must be bit-identical."""
    t, y, dy, _ = fake_lightcurve(500, 36)
    kernel, _ = build("drw+lor")
    gp = GP(kernel, mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    p = c2_truth()[None, :] + np.random.default_rng(6).normal(0, 0.1, (1000, 5))
    ll0, g0 = gp.grad_log_likelihood_batch(y, p)
    monkeypatch.setenv("MTG_TAPE_MB", "8")
    ll1, g1 = gp.grad_log_likelihood_batch(y, p)
    assert np.array_equal(ll0, ll1) and np.array_equal(g0, g1)


@pytest.mark.parametrize("segment", [None, "3"])
def test_gradient_at_c2_size(monkeypatch, segment):
    """config C2's own shape: N = 1000, DRW + Lorentzian, against complex-step; the reverse sweep's last segment is a
    partial one both with the default segment length (1000 % 6 = 4) and with a forced 3 (1000 % 3 = 1)."""
    if segment:
        monkeypatch.setenv("MTG_K2_SEGMENT", segment)
    t, y, dy, _ = fake_lightcurve(1000, 2)
    kernel, spec = build("drw+lor")
    gp = GP(kernel, mean=ConstantModel(float(np.mean(y))), fit_mean=False)
    gp.compute(t, dy + 1e-12)
    p = c2_truth()[None, :] + np.random.default_rng(12).normal(0, 0.1, (24, 5))
    ll, g, status = gp.grad_log_likelihood_batch(y, p, return_status=True)
    assert np.all(status == 0)
    np.testing.assert_allclose(ll, ogp.loglike_params(spec, p, t, y, dy), rtol=1e-10)
    assert_grad_close(g, ogp.grad_complex_step(spec, p, t, y, dy))


@pytest.mark.parametrize("mean_model", ["gaussian", "linear", "constant"])
def test_gpmodelling_fitted_mean_models_against_oracle(mean_model):
    """GPModelling(lc, kernel, mean_model=...) (gpmodelling.py:62-124): log-probability and gradient of the model
    GPModelling builds -- kernel parameters followed by the mean model's -- against the oracle with the same mean."""
    from mind_the_gaps_b200.gpmodelling import GPModelling
    from mind_the_gaps_b200.lightcurves import GappyLightcurve
    t, y, dy, exposures = fake_lightcurve(300, 41)
    bump = 0.2 * np.exp(-0.5 * ((t - t[150]) / (0.15 * (t[-1] - t[0]))) ** 2)
    lc = GappyLightcurve(t, y + bump, dy, exposures)
    kernel = DampedRandomWalk(np.log(0.002), np.log(2 * np.pi / (20 * 86400.0)), bounds=[(-15, 5), (-20, -5)])
    gpm = GPModelling(lc, kernel, mean_model=mean_model)
    p0 = gpm.gp.get_parameter_vector()
    nm = {"gaussian": 4, "linear": 2, "constant": 1}[mean_model]
    assert len(p0) == 2 + nm
    rng = np.random.default_rng(3)
    P = np.tile(p0, (12, 1))
    P[:, :2] += rng.normal(0, 0.1, (12, 2))
    if mean_model == "gaussian":
        P[:, 2] += rng.normal(0, 0.02, 12) * (t[-1] - t[0]); P[:, 3] *= rng.uniform(0.3, 0.9, 12)
        P[:, 4] *= rng.uniform(1.0, 2.0, 12); P[:, 5] += rng.normal(0, 0.01, 12)
    elif mean_model == "linear":
        P[:, 2] = rng.normal(0, 1e-9, 12); P[:, 3] = np.mean(lc.y) + rng.normal(0, 0.02, 12)
    else:
        P[:, 2] += rng.normal(0, 0.01, 12)
    fit_mean = True if mean_model == "constant" else mean_model
    ref = ogp.loglike_params([("drw",)], P, t, lc.y, dy, fit_mean=fit_mean)
    got = gpm._log_probability(P)
    inside = np.isfinite(got)
    assert inside.sum() >= 6
    np.testing.assert_allclose(got[inside], ref[inside], rtol=1e-10)
    np.testing.assert_allclose([gpm._log_probability(row) for row in P[inside][:3]], ref[inside][:3], rtol=1e-10)
    ll, g = gpm.gp.grad_log_likelihood_batch(lc.y, P[inside])
    gref = ogp.grad_complex_step([("drw",)], P[inside], t, lc.y, dy, fit_mean=fit_mean)
    # parameters of very different magnitude (times ~1e8 s, amplitudes ~1e7): compare each component on its own scale
    err = np.abs(g - gref) / np.max(np.abs(gref), axis=0, keepdims=True)
    assert np.max(err) < 1e-8, "scaled gradient error per parameter: %s" % np.max(err, axis=0)
