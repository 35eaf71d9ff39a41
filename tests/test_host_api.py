"""CPU tests of the host-side mirror of the reference API.  They restate the reference's own tests
(tests/gpmodelling_test.py, tests/models_test.py, tests/stats_test.py) against this package."""
import warnings

import numpy as np
import pytest

from mind_the_gaps_b200.celerite import GP, terms
from mind_the_gaps_b200.gpmodelling import GPModelling
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import BendingPowerlaw, DampedRandomWalk, Lorentzian, celerite_models, psd_models
from mind_the_gaps_b200 import stats as mstats
from mind_the_gaps_b200.sampler import EnsembleSampler, integrated_time


def _gpmodel(bounds_drw, bounds_lor, lor_params=(10, 5, -5), drw_params=(5.0, 10.0)):
    kernel = DampedRandomWalk(drw_params[0], drw_params[1], bounds=bounds_drw) + \
        Lorentzian(lor_params[0], lor_params[1], lor_params[2], bounds=bounds_lor)
    lc = GappyLightcurve(np.arange(100), np.arange(100), np.arange(100))
    return GPModelling(lc, kernel)


# ---- tests/gpmodelling_test.py ---------------------------------------------------------------------------

def test_parameters_within_bounds():
    bounds_drw = [(4.0, 6.0), (8.0, 12.0)]; bounds_lor = [(5, 15), (1, 6), (-7, -1)]
    gpmodel = _gpmodel(bounds_drw, bounds_lor)
    parameters = [5.0, 10.0, 10, 5, -5]
    bounds = bounds_drw + bounds_lor
    for percent, attempts in ((0.1, 100), (0.9, 2)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            samples = gpmodel.spread_walkers(100, parameters, bounds, percent=percent, max_attempts=attempts)
        for i, sample in enumerate(samples.T):
            assert np.all(np.logical_and(bounds[i][0] <= sample, sample <= bounds[i][1]))


def test_infinite_bounds():
    bounds_drw = [(None, None), (8.0, 12.0)]; bounds_lor = [(5, 15), (1, 6), (-7, -1)]
    gpmodel = _gpmodel(bounds_drw, bounds_lor)
    bounds = bounds_drw + bounds_lor
    samples = gpmodel.spread_walkers(100, [5.0, 10.0, 10, 5, -5], bounds, percent=0.1, max_attempts=50)
    assert np.all(np.isfinite(samples[:, 0]))
    for b, sample in zip(bounds[1:], samples.T[1:]):
        assert np.all(np.logical_and(b[0] <= sample, sample <= b[1]))


def test_percent_zero_gives_exact_parameters():
    bounds_drw = [(4.0, 6.0), (8.0, 12.0)]; bounds_lor = [(5, 15), (1, 6), (-7, -1)]
    gpmodel = _gpmodel(bounds_drw, bounds_lor)
    parameters = [5.0, 10.0, 10, 5, -5]
    samples = gpmodel.spread_walkers(50, parameters, bounds_drw + bounds_lor, percent=0.0)
    assert np.array_equal(samples, np.tile(parameters, (50, 1)))
    with pytest.raises(ValueError):
        gpmodel.spread_walkers(5, parameters, bounds_drw + bounds_lor, percent=1.5)


def test_max_attempts_with_impossibly_tight_bounds():          # gpmodelling_test.py:93-113
    drw, lor = [5.0, 10.0], [10, 5, -5]
    bounds_drw = [(v - 0.01, v + 0.01) for v in drw]; bounds_lor = [(v - 0.01, v + 0.01) for v in lor]
    gpmodel = _gpmodel(bounds_drw, bounds_lor)
    samples = gpmodel.spread_walkers(100, drw + lor, bounds_drw + bounds_lor, percent=0, max_attempts=50)
    for i, sample in enumerate(samples.T):
        assert np.all(sample == (drw + lor)[i])


def test_gp_parameter_protocol():
    bounds_drw = [(4.0, 6.0), (8.0, 12.0)]; bounds_lor = [(5, 15), (1, 6), (-7, -1)]
    gpmodel = _gpmodel(bounds_drw, bounds_lor)
    assert gpmodel.k == 5
    assert gpmodel.parameter_names == ("kernel:terms[0]:log_S0", "kernel:terms[0]:log_omega0", "kernel:terms[1]:log_S0",
                                       "kernel:terms[1]:log_Q", "kernel:terms[1]:log_omega0")
    assert gpmodel.gp.get_parameter_bounds() == bounds_drw + bounds_lor
    gpmodel.gp.set_parameter_vector([5.5, 9.0, 11.0, 2.0, -3.0])
    assert gpmodel.gp.log_prior() == 0.0
    gpmodel.gp.set_parameter_vector([6.5, 9.0, 11.0, 2.0, -3.0])
    assert gpmodel.gp.log_prior() == -np.inf
    with pytest.raises(AttributeError):
        gpmodel.max_loglikelihood
    with pytest.raises(AttributeError):
        gpmodel.mcmc_samples
    with pytest.raises(RuntimeError):
        gpmodel.generate_from_posteriors(2)
    # dict-style bounds and keyword parameters (tutorial_ppp.ipynb cell 7)
    k = terms.RealTerm(log_a=1.0, log_c=-2.0, bounds=dict(log_a=(-10, 50), log_c=(-10, 10)))
    assert k.get_parameter_bounds() == [(-10, 50), (-10, 10)]
    with pytest.raises(ValueError):
        terms.RealTerm(log_a=100.0, log_c=-2.0, bounds=dict(log_a=(-10, 50), log_c=(-10, 10)))
    # mean models of GPModelling._build_mean_model
    lc = GappyLightcurve(np.arange(50.0), np.random.default_rng(0).normal(10, 1, 50), np.ones(50))
    assert GPModelling(lc, DampedRandomWalk(0.0, 0.0), "constant").k == 3
    assert GPModelling(lc, DampedRandomWalk(0.0, 0.0), "linear").k == 4
    assert GPModelling(lc, DampedRandomWalk(0.0, 0.0), "gaussian").k == 6
    with pytest.raises(ValueError):
        GPModelling(lc, DampedRandomWalk(0.0, 0.0), "quadratic")
    bpl = BendingPowerlaw(2.0, 1.0, 0.0)
    assert bpl.log_prior() == 0.0
    bpl.set_parameter_vector([1.0, 2.0, 0.0])
    assert bpl.log_prior() == -np.inf            # log_S0 < log_Q (celerite_models.py:85-90)
    with pytest.raises(ValueError):              # like celerite, a non-finite prior at construction raises
        BendingPowerlaw(1.0, 2.0, 0.0)


# ---- tests/models_test.py (6 decimals) ----------------------------------------------------------------------

def test_psd_identities():
    w = np.arange(1, 1000)
    np.testing.assert_array_almost_equal(psd_models.BendingPowerlaw(S0=10, omega0=5, Q=0.5)(w),
                                         celerite_models.DampedRandomWalk(log_S0=np.log(10), log_omega0=np.log(5)).get_psd(w))
    for Q in [10, 1, 1 / np.sqrt(2), 0.1]:
        np.testing.assert_array_almost_equal(psd_models.SHO(S0=10, omega0=5, Q=Q)(w),
                                             terms.SHOTerm(log_S0=np.log(10), log_Q=np.log(Q), log_omega0=np.log(5)).get_psd(w))
    for rho in [1, 10, 20]:
        np.testing.assert_array_almost_equal(psd_models.Matern32(sigma=10, rho=rho)(w),
                                             terms.Matern32Term(log_sigma=np.log(10), log_rho=np.log(rho), eps=1e-15).get_psd(w))
        np.testing.assert_array_almost_equal(psd_models.Matern32(sigma=10, rho=rho)(w), psd_models.Matern(w, 10, rho, 1, 3 / 2))
        np.testing.assert_array_almost_equal(psd_models.Matern52(sigma=10, rho=rho)(w), psd_models.Matern(w, 10, rho, 1, 5 / 2))
    for Q in [1, 10, 100]:
        for S0 in [1, 10, 100]:
            np.testing.assert_array_almost_equal(psd_models.Lorentzian(S0=S0, omega0=5, Q=Q)(w),
                                                 celerite_models.Lorentzian(log_S0=np.log(S0), log_Q=np.log(Q), log_omega0=np.log(5)).get_psd(w))
    s = psd_models.Lorentzian(1, 80, 0.6) + psd_models.BendingPowerlaw(1, 0.3)
    np.testing.assert_allclose(s(w), psd_models.Lorentzian(1, 80, 0.6)(w) + psd_models.BendingPowerlaw(1, 0.3)(w))


# ---- tests/stats_test.py ------------------------------------------------------------------------------------

def test_distribution_factories_and_chi_statistics():
    for mean, std in [(10.0, 2.0), (0.5, 0.3)]:
        d = mstats.create_log_normal(mean, std)
        assert abs(d.mean() - mean) < 1e-9 and abs(d.var() - std ** 2) < 1e-9
        d = mstats.create_uniform_distribution(mean, std)
        assert abs(d.mean() - mean) < 1e-12 and abs(d.var() - std ** 2) < 1e-12
    p = np.array([1.0, 2.0, 3.0]); m = np.array([2.0, 2.0, 2.0])
    assert mstats.chi_log_likehood(p, m) == pytest.approx(2 * np.sum(np.log(m) + p / m))
    sig = np.array([0.5, 1.0, 2.0])
    assert mstats.chi_cov(p, m, np.diag(1 / sig ** 2)) == pytest.approx(mstats.chi_square(p, m, sig))
    assert mstats.aicc(-10.0, 100, 3) == pytest.approx(2 * 3 + 20 + 2 * 3 * 4 / 96)
    assert mstats.bic(-10.0, 100, 3) == pytest.approx(20 + 3 * np.log(100))
    # the Kraft posterior integrates to one (np.math bit-rot of the reference fixed)
    from scipy import integrate
    assert integrate.quad(mstats.kraft_pdf(a=0, b=35).pdf, 0, 35, args=(3, 1.5))[0] == pytest.approx(1.0, abs=1e-6)


def test_lrt_pvalue_matches_percentileofscore():
    from scipy.stats import percentileofscore
    rng = np.random.default_rng(0)
    T = rng.chisquare(3, 1000)
    for t_obs in [0.0, 1.0, 7.3, T[17], 50.0]:
        assert mstats.lrt_pvalue(T, t_obs) == pytest.approx(1 - percentileofscore(T, t_obs) / 100)
    s = mstats.lrt_summary(T, 9.0)
    assert s["n"] == 1000 and s["hist"].sum() == 1000 and s["percentiles"][95.0] == pytest.approx(np.percentile(T, 95))
    assert np.allclose(mstats.lrt_statistic([-10.0, -8.0], [-7.0, -8.0]), [6.0, 0.0])


# ---- sampler (host logic; CPU log-probability) ------------------------------------------------------------------

def test_ensemble_sampler_recovers_gaussian():
    cov = np.array([[1.0, 0.6], [0.6, 2.0]]); icov = np.linalg.inv(cov)
    logp = lambda p: -0.5 * np.einsum("bi,ij,bj->b", p, icov, p)
    s = EnsembleSampler(20, 2, logp, vectorize=True, seed=1)
    p0 = np.random.default_rng(0).normal(size=(20, 2))
    for _ in s.sample(p0, iterations=3000):
        pass
    chain = s.get_chain(discard=500, flat=True)
    assert np.allclose(np.cov(chain.T), cov, atol=0.25) and np.allclose(chain.mean(axis=0), 0, atol=0.15)
    tau = s.get_autocorr_time(tol=0)
    assert tau.shape == (2,) and np.all(tau > 1) and np.all(tau < 200)
    assert s.get_chain(discard=10, thin=5).shape == ((3000 - 10 - 4 + 4) // 5, 20, 2)
    assert s.get_log_prob(discard=10, thin=5, flat=True).shape[0] == s.get_chain(discard=10, thin=5, flat=True).shape[0]
    with pytest.raises(ValueError):
        EnsembleSampler(3, 2, logp)
    with pytest.raises(ValueError):          # linearly dependent walkers (percent = 0 spreads cannot be sampled)
        next(EnsembleSampler(8, 2, logp, vectorize=True).sample(np.zeros((8, 2)), iterations=1))


def test_integrated_time_of_ar1():
    rng = np.random.default_rng(2)
    rho = 0.9
    x = np.zeros((20000, 4, 1))
    for i in range(1, len(x)):
        x[i] = rho * x[i - 1] + rng.normal(size=(4, 1))
    tau = integrated_time(x, tol=0)
    assert abs(tau[0] - (1 + rho) / (1 - rho)) < 3.0


def test_finite_difference_steps_follow_scipy_at_the_bounds():
    """fitting._forward_steps: the absolute 2-point steps scipy's L-BFGS-B takes without ``jac`` (1e-8 forward, backward
    at an upper bound, towards the farther bound in a box narrower than the step) -- checked against
    scipy.optimize._numdiff.approx_derivative on a function whose gradient is known."""
    from scipy.optimize._numdiff import approx_derivative
    from mind_the_gaps_b200.fitting import _forward_steps
    lo = np.array([-1.0, 0.0, 2.0, -np.inf, 0.0]); hi = np.array([1.0, 5.0, 2.0 + 1.2e-8, np.inf, 5e-9])
    x = np.array([0.3, 5.0, 2.0 + 1.0e-8, 7.0, 1e-9])
    h = _forward_steps(x, lo, hi)
    assert h[0] == 1e-8 and h[1] == -1e-8 and h[3] == 1e-8
    assert h[2] < 0 and abs(h[2]) <= 1.0e-8 + 1e-20          # neither side fits: towards the farther (lower) bound
    assert h[4] > 0 and h[4] <= 4e-9 + 1e-20
    assert np.all(x + h >= lo) and np.all(x + h <= hi)
    f = lambda z: np.sum(z ** 2) + z[0] * z[1]
    mine = np.array([(f(x + np.eye(5)[k] * ((x + np.eye(5)[k] * h[k])[k] - x[k])) - f(x)) / ((x + np.eye(5)[k] * h[k])[k] - x[k]) for k in range(5)])
    ref = approx_derivative(f, x, method="2-point", abs_step=1e-8, bounds=(lo, hi))
    np.testing.assert_allclose(mine[[0, 1, 3]], ref[[0, 1, 3]], rtol=1e-12)      # where the plain step fits, identical to scipy


def test_generate_lc_from_params_plumbing():
    """GPModelling._generate_lc_from_params (gpmodelling.py:516-539) with a stand-in simulator: parameters set on the GP,
    the kernel's PSD handed to the simulator, rates -> noise -> GappyLightcurve on the input timestamps."""
    bounds_drw = [(4.0, 6.0), (8.0, 12.0)]; bounds_lor = [(5, 15), (1, 6), (-7, -1)]
    gpmodel = _gpmodel(bounds_drw, bounds_lor)

    class FakeSimulator(object):
        psd_model = None

        def generate_lightcurve(self):
            self.psd_at_call = self.psd_model(np.array([0.1, 1.0]))
            return np.full(100, 2.0)

        def add_noise(self, rates):
            return rates + 0.5, np.full(len(rates), 0.25)

    sim = FakeSimulator()
    params = np.array([4.5, 9.0, 9.0, 2.0, -3.0])
    lc = gpmodel._generate_lc_from_params(params, sim)
    np.testing.assert_array_equal(gpmodel.gp.get_parameter_vector(), params)
    np.testing.assert_allclose(sim.psd_at_call, gpmodel.gp.kernel.get_psd(np.array([0.1, 1.0])))
    assert np.all(lc.y == 2.5) and np.all(lc.dy == 0.25) and np.array_equal(lc.times, np.arange(100))


def test_small_reference_helpers():
    """stats.chi_log_likehood(_nonyq), fit_N / chi_square_N and simulator.add_poisson_noise (stats.py:31-113,
    simulator.py:423-465)."""
    from mind_the_gaps_b200 import stats as st
    from mind_the_gaps_b200.simulator import add_poisson_noise
    rng = np.random.default_rng(0)
    model = 3.0 * np.linspace(0.01, 0.5, 50) ** -1.0
    powers = model * rng.exponential(1.0, 50)
    s = st.chi_log_likehood_nonyq(powers, model)
    assert s == 2.0 * np.sum(np.log(model) + powers / model) == st.chi_log_likehood(powers, model)
    with_nyq = st.chi_log_likehood(powers, model, nyquist=True)
    assert abs(with_nyq - (st.chi_log_likehood_nonyq(powers[:-1], model[:-1]) + np.log(np.pi * powers[-1] * model[-1]) + 2 * powers[-1] / model[-1])) < 1e-12
    # best-fitting normalisation of a model that is off by a factor 2.5
    data = 2.5 * model
    assert abs(st.fit_N(lambda N, p, m, sg: st.chi_square(p, np.ravel(N)[0] * m, sg), (data, model, np.ones(50)))[0] - 2.5) < 1e-4
    assert st.chi_square_N(data, model, np.ones(50)) < 1e-6
    np.random.seed(4)
    rates, dy = add_poisson_noise(np.full(2000, 0.5), 1000.0)
    np.random.seed(4)
    counts = np.random.poisson(np.full(2000, 500.0))
    np.testing.assert_array_equal(rates, counts / 1000.0)
    np.testing.assert_array_equal(dy, np.sqrt(counts) / 1000.0)
    rates_b, dy_b = add_poisson_noise(np.full(10, 0.5), 1000.0, background_counts=np.full(10, 20), bkg_rate_err=np.full(10, 0.001))
    assert np.all(dy_b > 0) and abs(np.mean(rates_b) - 0.5) < 0.1


def test_lrt_result_round_trip(tmp_path):
    from mind_the_gaps_b200.lrt import LRTResult
    rng = np.random.default_rng(0)
    res = LRTResult(rng.chisquare(3, 50), rng.normal(size=50), rng.normal(size=50), rng.normal(size=(50, 2)), rng.normal(size=(50, 5)),
                    np.zeros(50, dtype=np.int32), np.ones(50, dtype=np.int32))
    path = str(tmp_path / "lrt.npz")
    res.save(path)
    back = LRTResult.load(path)
    for k in LRTResult._FIELDS:
        np.testing.assert_array_equal(getattr(back, k), getattr(res, k))
    assert back.pvalue(5.0) == res.pvalue(5.0)


# ---- LRT nesting: the null model's slots inside the alternative's parameter vector (lrt.infer_null_slots) ----------
def test_infer_null_slots_matches_term_structure_not_names():
    from mind_the_gaps_b200 import lrt
    from mind_the_gaps_b200.celerite import terms
    from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
    # the reference tutorial's own models (tutorial_ppp.ipynb cells 7 / 9): the stripped names of the ComplexTerm's
    # first two parameters equal the RealTerm's -- the match has to be on the term, not on the names
    null = terms.RealTerm(0.0, 0.0)
    alt = terms.ComplexTerm(0.0, 0.0, 0.0) + terms.RealTerm(0.0, 0.0)
    assert lrt.infer_null_slots(null, alt) == [3, 4]
    assert lrt.infer_null_slots(DampedRandomWalk(0.0, 0.0), Lorentzian(0.0, 1.0, 0.0) + DampedRandomWalk(0.0, 0.0)) == [3, 4]
    assert lrt.infer_null_slots(DampedRandomWalk(0.0, 0.0), DampedRandomWalk(0.0, 0.0) + Lorentzian(0.0, 1.0, 0.0)) == [0, 1]
    assert lrt.infer_null_slots(DampedRandomWalk(0.0, 0.0), DampedRandomWalk(0.0, 0.0) + DampedRandomWalk(0.0, 0.0)) == [2, 3]
    # a frozen parameter in the leading term shifts the free-vector offsets
    alt2 = terms.ComplexTerm(0.0, 0.0, 0.0) + terms.RealTerm(0.0, 0.0)
    alt2.freeze_parameter("terms[0]:log_d")
    assert lrt.infer_null_slots(null, alt2) == [2, 3]
    with pytest.raises(ValueError):          # two interior matches: ambiguous
        lrt.infer_null_slots(DampedRandomWalk(0.0, 0.0), DampedRandomWalk(0.0, 0.0) + DampedRandomWalk(0.0, 0.0) + Lorentzian(0.0, 1.0, 0.0))
    with pytest.raises(ValueError):          # no structural match
        lrt.infer_null_slots(terms.RealTerm(0.0, 0.0), Lorentzian(0.0, 1.0, 0.0) + DampedRandomWalk(0.0, 0.0))


def test_alternative_starts_nest_the_null_optimum_for_tutorial_kernels():
    from mind_the_gaps_b200 import lrt
    from mind_the_gaps_b200.celerite import terms
    null = terms.RealTerm(-1.0, -2.0)
    alt = terms.ComplexTerm(-3.0, -4.0, -5.0, bounds=dict(log_a=(-10, 0), log_c=(-10, 0), log_d=(-10, 0))) + terms.RealTerm(-1.0, -2.0)
    slots = lrt.infer_null_slots(null, alt)
    x_null = np.array([[-1.5, -2.5], [-0.5, -1.0]])
    lo = np.array([-10, -10, -10, -np.inf, -np.inf]); hi = np.array([0, 0, 0, np.inf, np.inf])
    st = lrt.alternative_starts(alt.get_parameter_vector(), x_null, slots, lo, hi, 8, np.random.default_rng(0))
    np.testing.assert_array_equal(st[:, 0, 3:], x_null)                  # start 0: null optimum in the RealTerm's slots
    np.testing.assert_array_equal(st[:, 0, :3], np.tile([-3.0, -4.0, -5.0], (2, 1)))
    np.testing.assert_array_equal(st[:, 5:, 3:], np.repeat(x_null[:, None, :], 3, axis=1))      # uniform starts keep the null part


def test_device_prior_supported_flags_user_overrides():
    from mind_the_gaps_b200.celerite import GP, terms
    from mind_the_gaps_b200.models import BendingPowerlaw, DampedRandomWalk

    class Strange(DampedRandomWalk):
        def log_prior(self):
            return -np.inf if self.log_S0 > 0 else super(Strange, self).log_prior()

    assert GP(DampedRandomWalk(0.0, 0.0) + terms.ComplexTerm(0.0, -1.0, 0.0, 0.0)).device_prior_supported()
    assert GP(BendingPowerlaw(0.0, -1.0, 0.0)).device_prior_supported()
    assert not GP(Strange(0.0, 0.0) + DampedRandomWalk(0.0, 0.0)).device_prior_supported()


def test_lrt_result_counts_used_simulations():
    from mind_the_gaps_b200.lrt import LRTResult
    T = np.array([1.0, 2.0, np.nan, 4.0])
    r = LRTResult(T, T, T, np.zeros((4, 1)), np.zeros((4, 1)), np.zeros(4, int), np.array([0, 0, 3, 0]))
    assert r.n_used == 3 and r.n_invalid == 1
    p, n = r.pvalue(2.5, return_n=True)
    assert n == 3 and abs(p - (1 - 2 / 3)) < 1e-12


def test_user_defined_term_compiles_to_raw_coefficient_model():
    """A Term subclass without a device map (the way the reference writes its kernels, celerite_models.py:7-34): the host
    evaluates the coefficients per parameter set; the degenerate (0, 0) real term is dropped."""
    from mind_the_gaps_b200.celerite import GP, terms

    class MyLorentzian(terms.Term):
        parameter_names = ("log_S0", "log_Q", "log_omega0")

        def get_real_coefficients(self, params):
            return 0, 0

        def get_complex_coefficients(self, params):
            log_S0, log_Q, log_omega0 = params
            w0 = np.exp(log_omega0)
            return np.exp(log_S0), 0.0, 0.5 * w0 / np.exp(log_Q), w0

    gp = GP(MyLorentzian(-6.0, 2.0, -12.0) + terms.RealTerm(-5.0, -13.0), mean=0.3)
    assert gp.needs_host_coefficients() and not gp.device_prior_supported()
    with pytest.raises(NotImplementedError):
        gp.device_model()
    P = np.array([[-6.0, 2.0, -12.0, -5.0, -13.0], [-5.5, 1.0, -11.0, -4.0, -12.0]])
    m, rows = gp._raw_model_and_rows(P)
    assert m.n_terms == 3 and list(m.kind[:3]) == [10, 11, 12] and m.n_free == 7
    w0 = np.exp(P[:, 2])
    expect = np.column_stack([np.exp(P[:, 3]), np.exp(P[:, 4]), np.exp(P[:, 0]), np.zeros(2), 0.5 * w0 / np.exp(P[:, 1]), w0, np.zeros(2)])
    np.testing.assert_allclose(rows, expect, rtol=1e-15)
    np.testing.assert_array_equal(gp.get_parameter_vector(), P[0])          # the GP's own state is restored
