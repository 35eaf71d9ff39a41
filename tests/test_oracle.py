"""CPU tests pinning the oracle (no GPU): recursion vs dense Cholesky, the PSD identities the
reference pins in tests/models_test.py, the C restatement vs numpy."""
import numpy as np
import pytest

from oracle import cref, gp, terms

from helpers import fake_lightcurve, c2_truth


SPECS = [
    ([("drw",)], [np.log(0.01), np.log(1e-6)]),
    ([("drw",), ("lorentzian",)], list(c2_truth())),
    ([("complex3",), ("real",)], [-6.0, -16.0, -11.8, -6.0, -13.0]),
    ([("sho",), ("sho",), ("sho",), ("real",)], [-6, 1.0, -12.0, -7, 0.5, -11.0, -8, 2.0, -10.5, -6.0, -14.0]),
    ([("sho",), ("jitter",)], [-6.0, np.log(0.2), -12.0, -4.0]),
    ([("matern32", 0.01), ("real",)], [-2.0, 13.0, -6.0, -13.0]),
    ([("bendingpl",), ("cosinus",)], [-5.0, -7.0, -12.0, -8.0, -11.0]),
    ([("complex4",)], [-5.0, -8.0, -12.0, -12.5]),
]


@pytest.mark.parametrize("spec,p0", SPECS)
def test_recursion_matches_dense(spec, p0):
    t, y, dy, _ = fake_lightcurve(300, 11)
    rng = np.random.default_rng(5)
    p = np.asarray(p0)[None, :] + rng.normal(0, 0.05, (3, len(p0)))
    a = gp.loglike_params(spec, p, t, y, dy)
    b = gp.loglike_params(spec, p, t, y, dy, dense=True)
    assert np.all(np.isfinite(b))
    np.testing.assert_allclose(a, b, rtol=1e-11)


def test_c_restatement_matches_numpy():
    t, y, dy, _ = fake_lightcurve(500, 3)
    spec = [("drw",), ("lorentzian",)]
    p = c2_truth()[None, :] + np.random.default_rng(0).normal(0, 0.1, (64, 5))
    a = cref.loglike_params(spec, p, t, y, dy)
    b = gp.loglike_params(spec, p, t, y, dy)
    np.testing.assert_allclose(a, b, rtol=1e-12)


def test_complex_step_gradient_matches_central_differences():
    t, y, dy, _ = fake_lightcurve(200, 4)
    spec = [("drw",), ("lorentzian",)]
    p = c2_truth()[None, :]
    g = gp.grad_complex_step(spec, p, t, y, dy)
    eps = 1e-6
    for k in range(5):
        pp = p.copy(); pp[:, k] += eps
        pm = p.copy(); pm[:, k] -= eps
        fd = (gp.loglike_params(spec, pp, t, y, dy) - gp.loglike_params(spec, pm, t, y, dy)) / (2 * eps)
        assert abs(fd[0] - g[0, k]) <= 1e-5 * max(1.0, abs(g[0, k]))


def test_not_positive_definite_is_minus_inf():
    t, y, dy, _ = fake_lightcurve(100, 4)
    # complex term with b >> a violates a c >= b d
    ll = gp.loglike_params([("complex4",)], [[-5.0, 3.0, -12.0, -11.0]], t, y, dy)
    assert ll[0] == -np.inf
    assert gp.loglike_params([("complex4",)], [[-5.0, 3.0, -12.0, -11.0]], t, y, dy, dense=True)[0] == -np.inf


# ---- PSD identities pinned by the reference (tests/models_test.py:14-102, 6 decimals) ------------

def test_psd_drw_equals_bending_powerlaw():
    w = np.arange(1, 1000)
    np.testing.assert_array_almost_equal(terms.psd_bendingpowerlaw(w, S0=10, omega0=5, Q=0.5),
                                         terms.psd([("drw",)], [np.log(10), np.log(5)], w))


@pytest.mark.parametrize("Q", [10, 1, 1 / np.sqrt(2), 0.1])
def test_psd_sho(Q):
    w = np.arange(1, 1000)
    np.testing.assert_array_almost_equal(terms.psd_sho(w, S0=10, Q=Q, omega0=5),
                                         terms.psd([("sho",)], [np.log(10), np.log(Q), np.log(5)], w))


@pytest.mark.parametrize("rho", [1, 10, 20])
def test_psd_matern32(rho):
    w = np.arange(1, 1000)
    np.testing.assert_array_almost_equal(terms.psd_matern32(w, sigma=10, rho=rho),
                                         terms.psd([("matern32", 1e-15)], [np.log(10), np.log(rho)], w))


@pytest.mark.parametrize("Q", [1, 10, 100])
@pytest.mark.parametrize("S0", [1, 10, 100])
def test_psd_lorentzian(Q, S0):
    w = np.arange(1, 1000)
    np.testing.assert_array_almost_equal(terms.psd_lorentzian(w, S0=S0, Q=Q, omega0=5),
                                         terms.psd([("lorentzian",)], [np.log(S0), np.log(Q), np.log(5)], w))


@pytest.mark.parametrize("floating_mean", [False, True])
def test_lombscargle_restatement_matches_scipy(floating_mean):
    """oracle.periodogram (astropy's algorithm restated) against scipy.signal.lombscargle, an independent implementation:
    'psd' normalisation with unit errors = scipy's unnormalised power; 'standard' = scipy's normalize=True."""
    from scipy.signal import lombscargle as sp_ls
    from oracle import periodogram as op
    rng = np.random.default_rng(0)
    t = np.sort(rng.uniform(0, 300.0, 180))
    y = 3.0 + np.sin(2 * np.pi * 0.07 * t) + 0.5 * rng.normal(size=len(t))
    f = np.linspace(0.005, 0.5, 97)
    mine = op.lombscargle(t, y, f, fit_mean=floating_mean, center_data=True, normalization="psd")
    ref = sp_ls(t, y, 2 * np.pi * f, normalize=False, floating_mean=floating_mean) if floating_mean else \
        sp_ls(t, y - y.mean(), 2 * np.pi * f, normalize=False)
    np.testing.assert_allclose(mine, ref, rtol=1e-9)
    mine_n = op.lombscargle(t, y, f, fit_mean=floating_mean, center_data=True, normalization="standard")
    ref_n = sp_ls(t, y, 2 * np.pi * f, normalize=True, floating_mean=floating_mean) if floating_mean else \
        sp_ls(t, y - y.mean(), 2 * np.pi * f, normalize=True)
    np.testing.assert_allclose(mine_n, ref_n, rtol=1e-9)


@pytest.mark.parametrize("spec,p0", [
    ([("drw",), ("lorentzian",)], None),
    ([("sho",)], [-6.0, np.log(3.0), np.log(2 * np.pi / (30 * 86400))]),
    ([("sho",)], [-6.0, np.log(0.3), np.log(2 * np.pi / (30 * 86400))]),
    ([("complex4",), ("real",)], [-7.0, -8.5, -14.0, -12.5, -6.0, -15.0]),
    ([("matern32", 0.01)], [-3.0, 14.0]),
    ([("sho",), ("sho",)], [-6.0, np.log(3.0), -12.0, -7.0, np.log(2.0), -11.0]),
])
def test_kalman_and_scan_forms_equal_the_dense_likelihood(spec, p0):
    """The formulation behind kernels K9 / K10: state-space model of a celerite kernel, sequential Kalman filter, and the
    parallel-in-time scan, against the dense Cholesky oracle -- also with measurement errors of zero."""
    from oracle import kalman, terms as ot
    from helpers import c2_truth, fake_lightcurve
    t, y, dy, _ = fake_lightcurve(300, 2)
    p = np.asarray(c2_truth() if p0 is None else p0, dtype=float)
    co = ot.coefficients(spec, p)
    ar, cr, ac, bc, cc, dc = [np.atleast_1d(np.asarray(x, dtype=float)) for x in co[:6]]
    jit = float(co[6]) if len(co) > 6 else 0.0
    r = y - np.mean(y)
    for scale in (1.0, 0.0):
        diag = (dy * scale + 1e-12) ** 2
        if scale:
            ref = gp.loglike_dense(t, r, diag, ar, cr, ac, bc, cc, dc, jitter=jit)
        else:
            ref = float(np.ravel(np.real(gp.loglike_recursion(t, r, diag, ar, cr, ac, bc, cc, dc, jitter=jit)))[0])
        model = kalman.statespace(ar, cr, ac, bc, cc, dc)
        assert abs(kalman.loglike_sequential(t, r, diag + jit, model) - ref) <= 1e-11 * abs(ref)
        assert abs(kalman.loglike_scan(t, r, diag + jit, model, threads=64) - ref) <= 1e-11 * abs(ref)


def test_forward_mode_derivatives_survive_the_scan():
    """Complex-step differentiation THROUGH the parallel-in-time scan (operator inverse included) reproduces the gradient of
    the sequential recursion to rounding: the numerical groundwork for a scan-based gradient kernel."""
    from oracle import kalman
    t, y, dy, _ = fake_lightcurve(200, 2)
    spec = [("drw",), ("lorentzian",)]
    p = c2_truth()
    g = kalman.grad_complex_step_scan(spec, p, t, y - np.mean(y), (dy + 1e-12) ** 2)
    ref = gp.grad_complex_step(spec, p[None, :], t, y, dy)[0]
    np.testing.assert_allclose(g, ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max())


# ---- oracle.sampler: emcee 3.1.4's stretch move restated (gpmodelling.py:247-264) -----------------------------------------
def test_oracle_sampler_draw_order_and_posterior():
    """The restated stretch move (i) consumes the RandomState exactly as emcee does -- per iteration shuffle(inds), then per
    half rand(Ns), randint(Nc, Ns) and one rand() per accept test -- and (ii) samples the right distribution: mean and
    covariance of a correlated 2-d Gaussian."""
    from oracle import sampler as osamp
    W, nsteps = 10, 3
    rs = np.random.RandomState(42)
    split, uz, rint, uacc = osamp.draws_like_emcee(rs, W, nsteps)
    rs2 = np.random.RandomState(42)
    for it in range(nsteps):
        inds = np.arange(W) % 2
        rs2.shuffle(inds)
        assert np.array_equal(split[it], inds) and inds.sum() == W // 2
        for h in range(2):
            ns = int(np.sum(inds == h))
            np.testing.assert_array_equal(uz[it, h, :ns], rs2.rand(ns))
            np.testing.assert_array_equal(rint[it, h, :ns], rs2.randint(W - ns, size=(ns,)))
            np.testing.assert_array_equal(uacc[it, h, :ns], [rs2.rand() for _ in range(ns)])
    cov = np.array([[1.0, 0.6], [0.6, 2.0]]); icov = np.linalg.inv(cov); mu = np.array([0.5, -1.0])
    logp = lambda x: -0.5 * np.einsum("bi,ij,bj->b", x - mu, icov, x - mu)
    W, nsteps = 32, 4000
    rs = np.random.RandomState(7)
    p0 = mu + rs.normal(0, 0.1, (W, 2))
    chain, lp, nacc = osamp.stretch_move_run(logp, p0, nsteps, *osamp.draws_like_emcee(rs, W, nsteps))
    flat = chain[500:].reshape(-1, 2)
    np.testing.assert_allclose(lp[-1], logp(chain[-1]), rtol=1e-12)
    assert np.all(np.abs(flat.mean(axis=0) - mu) < 0.08)
    np.testing.assert_allclose(np.cov(flat.T), cov, atol=0.2)
    assert 0.5 < nacc.mean() / nsteps < 0.85          # ~0.7 for a 2-d Gaussian with a = 2
    # a proposal is c - (c - s) zz with zz = ((a - 1) u + 1)^2 / a: u = 0 contracts the walker towards its partner by 1/2
    one = osamp.stretch_move_run(lambda x: np.zeros(len(x)), np.array([[0.0], [4.0]]), 1, np.array([[0, 1]]),
                                 np.zeros((1, 2, 1)), np.zeros((1, 2, 1), dtype=int), np.zeros((1, 2, 1)))[0]
    np.testing.assert_array_equal(one[0], [[2.0], [3.0]])


# ---- oracle.kalman.grad_chunked_adjoint: the NumPy statement of kernel K11 --------------------------------------------------
@pytest.mark.parametrize("threads", [1, 3, 8, 64, 300])
def test_chunked_adjoint_equals_complex_step_for_any_chunking(threads):
    """Reverse mode through the Kalman filter chunk by chunk -- per-chunk (Phi, q, Omega), suffix composition, independent
    chunk adjoint sweeps (what K11 runs with one chunk per thread) -- gives the gradient of the celerite log-likelihood with
    respect to every coefficient, the jitter and the residuals: equal to complex-step differentiation of the oracle's
    recursion, whatever the number of chunks (more chunks than observations included)."""
    from oracle import kalman
    t, y, dy, _ = fake_lightcurve(157, 3)
    resid = y - y.mean(); diag = (dy + 1e-12) ** 2
    ar = np.array([0.003, 0.001]); cr = np.array([2e-6, 3e-7])
    ac = np.array([0.002, 0.0015]); bc = np.array([0.0, 0.0004]); cc = np.array([1e-7, 4e-7]); dc = np.array([7e-6, 2e-6]); jit = 1e-4
    ll, g = kalman.grad_chunked_adjoint(t, resid, diag, ar, cr, ac, bc, cc, dc, jitter=jit, threads=threads)
    co = dict(ar=ar, cr=cr, ac=ac, bc=bc, cc=cc, dc=dc)
    ref = gp.loglike_recursion(t, resid, diag, *[co[k][None] for k in ("ar", "cr", "ac", "bc", "cc", "dc")], jitter=jit)[0]
    assert abs(ll - ref) < 1e-12 * abs(ref)
    h = 1e-30
    for name in co:
        for j in range(2):
            c = {k: v.astype(complex)[None] for k, v in co.items()}
            c[name][0, j] += 1j * h
            d = np.imag(gp.loglike_recursion(t, resid, diag, c["ar"], c["cr"], c["ac"], c["bc"], c["cc"], c["dc"], jitter=jit)[0]) / h
            assert abs(g[name][j] - d) <= 1e-11 * max(abs(d), 1.0), (name, j, g[name][j], d)
    d = np.imag(gp.loglike_recursion(t, resid, diag, *[co[k][None] for k in ("ar", "cr", "ac", "bc", "cc", "dc")], jitter=jit + 1j * h)[0]) / h
    assert abs(g["jitter"] - d) <= 1e-11 * abs(d)
    rc = resid.astype(complex); rc[40] += 1j * h
    d = np.imag(gp.loglike_recursion(t, rc, diag, *[co[k][None] for k in ("ar", "cr", "ac", "bc", "cc", "dc")], jitter=jit)[0]) / h
    assert abs(g["resid"][40] - d) <= 1e-11 * abs(d)
