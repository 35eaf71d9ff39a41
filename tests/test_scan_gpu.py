"""GPU parity: K9, the parallel-in-time (associative scan) log-likelihood kernel for small batches, against the
oracle and against K1, through the C ABI.  Tolerance 1e-10 relative (north_star: 1e-9)."""
import ctypes

import numpy as np
import pytest

from mind_the_gaps_b200 import _native
from mind_the_gaps_b200.celerite import GP, terms
from mind_the_gaps_b200.celerite.modeling import ConstantModel
from mind_the_gaps_b200.models import DampedRandomWalk
from mind_the_gaps_b200.models.mean_models import LinearModel
from oracle import gp as ogp

from helpers import c2_truth, fake_lightcurve
from test_loglike_gpu import build

pytestmark = pytest.mark.gpu
RTOL = 1e-10
# kernel structures with <= 4 state dimensions: what K9 covers
SCAN_KINDS = ["drw", "drw+lor", "complex+real", "sho_overdamped+jitter", "complex4"]


def extra(kind):
    if kind == "sho_under":
        return terms.SHOTerm(-6.0, np.log(3.0), -12.0), [("sho",)]
    if kind == "sho+real":
        return terms.SHOTerm(-6.0, 0.0, -12.0) + terms.RealTerm(-6.0, -13.0), [("sho",), ("real",)]
    if kind == "3real":
        return (terms.RealTerm(-6.0, -13.0) + terms.RealTerm(-7.0, -11.0) + terms.RealTerm(-5.0, -15.0),
                [("real",), ("real",), ("real",)])
    if kind == "2sho":                   # J = 4, both Q regimes in the second term across the batch
        return (terms.SHOTerm(-6.0, np.log(3.0), -12.0) + terms.SHOTerm(-7.0, 0.0, -11.0), [("sho",), ("sho",)])
    if kind == "2real+complex4":        # J = 4
        return (terms.RealTerm(-6.0, -13.0) + terms.RealTerm(-7.0, -11.0) + terms.ComplexTerm(-5.0, -8.0, -12.0, -12.5),
                [("real",), ("real",), ("complex4",)])
    if kind == "bpl":
        from mind_the_gaps_b200.models import BendingPowerlaw
        return BendingPowerlaw(-5.0, -7.0, -12.0), [("bendingpl",)]
    if kind == "cos+real":
        from mind_the_gaps_b200.models import Cosinus
        return Cosinus(-8.0, -11.0) + terms.RealTerm(-6.0, -13.0), [("cosinus",), ("real",)]
    return build(kind)


ALL_KINDS = SCAN_KINDS + ["sho_under", "sho+real", "3real", "bpl", "cos+real", "matern+real", "2sho", "2real+complex4"]


@pytest.mark.parametrize("kind", ALL_KINDS)
@pytest.mark.parametrize("N", [1, 2, 31, 255, 256, 257, 1000, 4099])
def test_scan_matches_oracle_and_k1(kind, N):
    t, y, dy, _ = fake_lightcurve(max(N, 2), 21)
    t, y, dy = t[:N], y[:N], dy[:N]
    kernel, spec = extra(kind)
    gp = GP(kernel, mean=ConstantModel(float(np.mean(y))), fit_mean=False)
    gp.compute(t, dy + 1e-12)
    p0 = gp.get_parameter_vector()
    B = 37
    p = p0[None, :] + np.random.default_rng(N).normal(0, 0.05, (B, len(p0)))
    if kind == "sho_overdamped+jitter":
        p[:, 1] = np.log(0.2) + np.random.default_rng(1).normal(0, 0.05, B)
    if kind == "sho+real":
        p[:, 1] = np.log(np.linspace(0.1, 3.0, B))          # both SHO regimes in one launch
    if kind == "2sho":
        p[:, 4] = np.log(np.linspace(0.1, 3.0, B))
    ll, status = gp.log_likelihood_batch(y, p, return_status=True, kernel="scan")
    # the oracle refuses batches straddling Q = 1/2: row by row there
    ref = (np.concatenate([ogp.loglike_params(spec, p[i:i + 1], t, y, dy) for i in range(B)]) if kind in ("sho+real", "2sho")
           else ogp.loglike_params(spec, p, t, y, dy))
    assert np.all(status == 0)
    np.testing.assert_allclose(ll, ref, rtol=RTOL)
    k1 = gp.log_likelihood_batch(y, p, kernel="k1")
    np.testing.assert_allclose(ll, k1, rtol=RTOL)


def test_scan_is_the_automatic_choice_for_small_batches_and_k1_for_large():
    """Same numbers either way; the launch counter tells which kernel ran (one launch each)."""
    t, y, dy, _ = fake_lightcurve(1000, 3)
    kernel, spec = build("drw+lor")
    gp = GP(kernel, mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    p = c2_truth()[None, :] + np.random.default_rng(0).normal(0, 0.1, (32, 5))
    auto = gp.log_likelihood_batch(y, p)
    np.testing.assert_array_equal(auto, gp.log_likelihood_batch(y, p, kernel="scan"))
    np.testing.assert_allclose(auto, ogp.loglike_params(spec, p, t, y, dy), rtol=RTOL)
    big = np.tile(p, (128, 1))                                 # 4096 series: above the scan threshold (2 N)
    np.testing.assert_array_equal(gp.log_likelihood_batch(y, big), gp.log_likelihood_batch(y, big, kernel="k1"))
    # a model the scan kernel does not cover falls back silently in auto mode and refuses when forced
    gp7 = GP(build("3sho+real")[0], mean=float(np.mean(y)))
    gp7.compute(t, dy + 1e-12)
    p7 = gp7.get_parameter_vector()[None, :]
    np.testing.assert_array_equal(gp7.log_likelihood_batch(y, p7), gp7.log_likelihood_batch(y, p7, kernel="k1"))
    with pytest.raises(_native.MtgError):
        gp7.log_likelihood_batch(y, p7, kernel="scan")


def test_scan_prior_not_positive_definite_and_high_signal_to_noise():
    t, y, dy, _ = fake_lightcurve(500, 6)
    kernel, spec = build("drw+lor")
    gp = GP(kernel, mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    p = np.tile(gp.get_parameter_vector(), (4, 1))
    p[1, 0] = 31.0; p[2, 3] = -6.0
    ll, status = gp.log_likelihood_batch(y, p, apply_prior=True, return_status=True, kernel="scan")
    assert list(status) == [0, 3, 3, 0] and ll[1] == -np.inf and ll[2] == -np.inf
    gp4 = GP(terms.ComplexTerm(-5.0, -8.0, -12.0, -12.5), mean=float(np.mean(y)))
    gp4.compute(t, dy + 1e-12)
    ll4, st4 = gp4.log_likelihood_batch(y, np.array([[-5.0, 3.0, -12.0, -11.0]]), return_status=True, kernel="scan")
    assert ll4[0] == -np.inf and st4[0] == 1
    # measurement errors of zero (diag = 1e-24): the operator's (I + C J)^-1 stays well conditioned
    gp0 = GP(build("drw+lor")[0], mean=float(np.mean(y)))
    gp0.compute(t, np.zeros_like(dy) + 1e-12)
    p = c2_truth()[None, :] + np.random.default_rng(5).normal(0, 0.3, (16, 5))
    ll0 = gp0.log_likelihood_batch(y, p, kernel="scan")
    ref = ogp.loglike_params(spec, p, t, y, np.zeros_like(dy))
    np.testing.assert_allclose(ll0, ref, rtol=1e-9)


def test_scan_fitted_means_and_per_dataset_layout():
    t, y, dy, _ = fake_lightcurve(700, 8)
    gp = GP(DampedRandomWalk(np.log(0.01), np.log(1e-6)), mean=LinearModel(1e-9, float(np.mean(y))), fit_mean=True)
    gp.compute(t, dy + 1e-12)
    p = np.tile(gp.get_parameter_vector(), (9, 1))
    p[:, -1] += np.linspace(-0.05, 0.05, 9)
    np.testing.assert_allclose(gp.log_likelihood_batch(y, p, kernel="scan"), gp.log_likelihood_batch(y, p, kernel="k1"), rtol=RTOL)
    # S data sets x group parameter sets, frozen per-data-set mean (the LRT layout), through the C ABI
    S, group, N = 5, 13, 333
    t, _, _, _ = fake_lightcurve(N, 9)
    rng = np.random.default_rng(S)
    Y = 0.5 + 0.05 * rng.normal(size=(S, N)); DY = rng.uniform(0.01, 0.03, (S, N))
    kernel, spec = build("drw+lor")
    model = GP(kernel, mean=0.0).device_model()
    B = S * group
    p = c2_truth()[None, :] + rng.normal(0, 0.1, (B, 5))
    ll = np.empty(B); status = np.empty(B, dtype=np.int32)
    diag = np.ascontiguousarray((DY + 1e-12) ** 2); means = np.ascontiguousarray(Y.mean(axis=1))
    _native.check(_native.lib().mtg_loglike_host(ctypes.byref(model), N, S, group, B, _native.ptr(np.ascontiguousarray(t)),
                                                 _native.ptr(Y), _native.ptr(diag), _native.ptr(means), _native.ptr(p),
                                                 _native.ptr(ll), _native.ptr(status), _native.FLAG_DS_MEAN | _native.FLAG_SCAN))
    ref = np.concatenate([ogp.loglike_params(spec, p[s * group:(s + 1) * group], t, Y[s], DY[s]) for s in range(S)])
    assert np.all(status == 0)
    np.testing.assert_allclose(ll, ref, rtol=RTOL)


def test_scan_long_series_property():
    """N = 10^5 (config C5's length) on a narrow kernel: scan == K1 to 1e-10; both obey the scaling identity."""
    N = 100000
    rng = np.random.default_rng(6)
    t = np.cumsum(rng.uniform(0.5, 1.5, N)) * 600.0
    y = 1.0 + 0.05 * np.cumsum(rng.normal(0, 0.05, N)) + rng.normal(0, 0.01, N)
    dy = np.full(N, 0.01)
    gp = GP(terms.SHOTerm(-9.0, np.log(3.0), np.log(2 * np.pi / (30 * 86400.0))) + terms.RealTerm(-8.0, -16.0), mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    p = gp.get_parameter_vector()[None, :] + rng.normal(0, 0.1, (8, 5))
    a = gp.log_likelihood_batch(y, p, kernel="scan"); b = gp.log_likelihood_batch(y, p, kernel="k1")
    assert np.all(np.isfinite(a))
    np.testing.assert_allclose(a, b, rtol=RTOL)


@pytest.mark.parametrize("x", [0.1, 30.0, 1000.0])
def test_scan_matern32_conditioning(x):
    """celerite's Matern-3/2 term = complex term with d = eps, b / a = w0 / eps = x: the stationary covariance of the slot
    in the scan kernel has condition number ~4 x^2.  Densely sampled series (lags ~ 1 / w0) so that it matters."""
    rng = np.random.default_rng(int(x))
    w0 = 0.01 * x                                     # eps = 0.01
    N = 600
    t = np.cumsum(rng.uniform(0.2, 3.0, N)) / w0
    y = rng.normal(0.0, 1.0, N).cumsum() * 0.1 + rng.normal(0, 0.05, N)
    dy = np.full(N, 0.05)
    gp = GP(terms.Matern32Term(0.0, np.log(np.sqrt(3.0) / w0)), mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    p = gp.get_parameter_vector()[None, :] + rng.normal(0, 0.1, (16, 2))
    ll = gp.log_likelihood_batch(y, p, kernel="scan")
    ref = ogp.loglike_params([("matern32", 0.01)], p, t, y, dy)
    np.testing.assert_allclose(ll, ref, rtol=1e-9)
