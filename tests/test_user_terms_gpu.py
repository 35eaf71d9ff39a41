"""GPU: user-defined celerite Term subclasses (no device coefficient map) through the raw-coefficient entry
(MTG_TERM_RAW_*; the reference's own kernels are such subclasses, models/celerite_models.py:7-90)."""
import ctypes
import warnings

import numpy as np
import pytest

from mind_the_gaps_b200 import _native
from mind_the_gaps_b200.celerite import GP, terms
from mind_the_gaps_b200.gpmodelling import GPModelling
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from oracle import gp as ogp

from helpers import C2_BOUNDS, c2_truth, fake_lightcurve

pytestmark = pytest.mark.gpu


class MyLorentzian(terms.Term):
    """The reference's Lorentzian written as a user would (celerite_models.py:7-34), without any device hint."""
    parameter_names = ("log_S0", "log_Q", "log_omega0")

    def get_real_coefficients(self, params):
        return 0, 0

    def get_complex_coefficients(self, params):
        log_S0, log_Q, log_omega0 = params
        w0 = np.exp(log_omega0)
        return np.exp(log_S0), 0.0, 0.5 * w0 / np.exp(log_Q), w0


class Wobble(terms.Term):
    """Something the library has never seen: a complex term with b tied to a, plus its own jitter and prior."""
    parameter_names = ("log_a", "log_c", "log_d", "log_s")

    def get_complex_coefficients(self, params):
        log_a, log_c, log_d, log_s = params
        a = np.exp(log_a)
        return a, 0.3 * a * np.exp(log_c - log_d), np.exp(log_c), np.exp(log_d)

    def get_jitter(self, params):
        return float(np.exp(2.0 * params[3]))

    def log_prior(self):
        return -np.inf if self.log_c > self.log_d else super(Wobble, self).log_prior()


@pytest.mark.parametrize("N,B", [(300, 40), (1000, 5000)])
def test_user_lorentzian_equals_builtin_and_oracle(N, B):
    t, y, dy, _ = fake_lightcurve(N, 2)
    p0 = c2_truth()
    user = GP(DampedRandomWalk(p0[0], p0[1], bounds=C2_BOUNDS[:2]) + MyLorentzian(p0[2], p0[3], p0[4], bounds=C2_BOUNDS[2:]), mean=float(np.mean(y)))
    builtin = GP(DampedRandomWalk(p0[0], p0[1], bounds=C2_BOUNDS[:2]) + Lorentzian(p0[2], p0[3], p0[4], bounds=C2_BOUNDS[2:]), mean=float(np.mean(y)))
    user.compute(t, dy + 1e-12); builtin.compute(t, dy + 1e-12)
    P = p0[None, :] + np.random.default_rng(3).normal(0, 0.1, (B, 5))
    for kern in ("k1", "scan") if B <= 2000 else ("auto",):
        lu = user.log_likelihood_batch(y, P, kernel=kern)
        np.testing.assert_allclose(lu, builtin.log_likelihood_batch(y, P, kernel=kern), rtol=1e-13)
    idx = np.arange(0, B, max(B // 40, 1))
    np.testing.assert_allclose(lu[idx], ogp.loglike_params([("drw",), ("lorentzian",)], P[idx], t, y, dy), rtol=1e-10)
    # single-point celerite API and the box prior (host side)
    user.set_parameter_vector(P[0])
    assert abs(user.log_likelihood(y) - lu[0]) <= 1e-12 * abs(lu[0])
    Pout = P[:4].copy(); Pout[1, 0] = 31.0
    lp = user.log_probability_batch(y, Pout)
    assert np.isneginf(lp[1]) and np.all(np.isfinite(lp[[0, 2, 3]]))


def test_unknown_user_term_with_jitter_and_prior_against_dense_oracle():
    t, y, dy, _ = fake_lightcurve(250, 7)
    k = Wobble(-6.0, -14.0, -12.0, -4.5, bounds=dict(log_a=(-12, 0), log_c=(-18, -8), log_d=(-16, -8), log_s=(-8, -2))) + terms.RealTerm(-6.5, -13.0)
    gp = GP(k, mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    P = gp.get_parameter_vector()[None, :] + np.random.default_rng(1).normal(0, 0.2, (16, 6))
    P[3, 1] = P[3, 2] + 0.5                                  # violates the term's own prior
    got = gp.log_probability_batch(y, P)
    assert np.isneginf(got[3])
    diag = (dy + 1e-12) ** 2
    for i in (0, 5, 11):
        a, c, d, s = np.exp(P[i, 0]), np.exp(P[i, 1]), np.exp(P[i, 2]), np.exp(2 * P[i, 3])
        ref = ogp.loglike_dense(t, y - np.mean(y), diag, [np.exp(P[i, 4])], [np.exp(P[i, 5])], [a], [0.3 * a * c / d], [c], [d], jitter=s)
        assert abs(got[i] - ref) <= 1e-10 * abs(ref)


def test_gpmodelling_with_a_user_term_fits_and_samples():
    t, y, dy, ex = fake_lightcurve(200, 2)
    p0 = c2_truth()
    lc = GappyLightcurve(t, y, dy, ex)
    gpm = GPModelling(lc, DampedRandomWalk(p0[0], p0[1], bounds=C2_BOUNDS[:2]) + MyLorentzian(p0[2], p0[3], p0[4], bounds=C2_BOUNDS[2:]))
    ref = GPModelling(lc, DampedRandomWalk(p0[0], p0[1], bounds=C2_BOUNDS[:2]) + Lorentzian(p0[2], p0[3], p0[4], bounds=C2_BOUNDS[2:]))
    assert abs(gpm._log_probability(p0) - ref._log_probability(p0)) < 1e-9
    sol = gpm.fit()                                           # scipy-driven: the device optimiser needs a device map
    sol_ref = ref.fit(optimizer="scipy")
    assert sol.success == sol_ref.success and abs(sol.fun - sol_ref.fun) < 1e-6 * abs(sol_ref.fun)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(1)
        gpm.derive_posteriors(fit=False, max_steps=40, convergence_steps=20, walkers=12, progress=False)
    assert gpm.mcmc_samples.shape[1] == 5 and np.all(np.isfinite(gpm.loglikelihoods))
    i = int(np.argmax(gpm.loglikelihoods))
    np.testing.assert_allclose(gpm.loglikelihoods[i], ref._log_probability(gpm.mcmc_samples[i]), rtol=1e-10)


def test_raw_coefficient_gradient_through_the_c_abi():
    """mtg_loglike_grad_host on a raw model returns d loglike / d coefficient: against complex-step on the coefficients."""
    N = 300
    t, y, dy, _ = fake_lightcurve(N, 5)
    m = _native.MtgModel()
    m.n_terms = 3
    m.kind[0] = 10; m.pidx[0][0] = 0; m.pidx[0][1] = 1
    m.kind[1] = 11
    for i in range(4):
        m.pidx[1][i] = 2 + i
    m.kind[2] = 12; m.pidx[2][0] = 6
    m.mean_kind = _native.MEAN_CONSTANT; m.mean_pidx[0] = 7
    m.n_full = 8; m.full_default[7] = float(np.mean(y))
    m.n_free = 7
    for k in range(7):
        m.free_idx[k] = k; m.lo[k] = -np.inf; m.hi[k] = np.inf
    rng = np.random.default_rng(2)
    base = np.array([3e-3, 2e-6, 2e-3, 4e-4, 4e-7, 7e-6, 1e-4])
    P = np.ascontiguousarray(base[None, :] * np.exp(rng.normal(0, 0.1, (6, 7))))
    diag = np.ascontiguousarray((dy + 1e-12) ** 2)
    for flag in (_native.FLAG_NO_SCAN, _native.FLAG_SCAN):
        ll = np.empty(6); g = np.empty((6, 7)); st = np.empty(6, dtype=np.int32)
        _native.check(_native.lib().mtg_loglike_grad_host(ctypes.byref(m), N, 1, 6, 6, _native.ptr(np.ascontiguousarray(t)), _native.ptr(np.ascontiguousarray(y)),
                                                          _native.ptr(diag), None, _native.ptr(P), _native.ptr(ll), _native.ptr(g), _native.ptr(st), flag))
        assert np.all(st == 0)
        resid = y - np.mean(y)
        for b in range(6):
            ref = np.empty(7)
            for k in range(7):
                pc = P[b].astype(complex); pc[k] += 1e-30j
                v = ogp.loglike_recursion(t, resid, diag, pc[None, 0:1], pc[None, 1:2], pc[None, 2:3], pc[None, 3:4], pc[None, 4:5], pc[None, 5:6], jitter=pc[6])[0]
                ref[k] = v.imag / 1e-30
                if k == 0:
                    assert abs(v.real - ll[b]) <= 1e-10 * abs(ll[b])
            np.testing.assert_allclose(g[b], ref, rtol=1e-8)
