"""CPU: pin the simulator oracle on the reference's golden vectors (tests/simulator_test.py:192-253,
tests/stats_test.py:9-34) and on independent numerics for the Kraft posterior."""
import numpy as np
import pytest
from scipy import integrate, optimize, stats

from oracle import simulator as osim


def _downsampling_case(exposures):
    timestamps = np.append(np.arange(1, 3.1, 2), np.arange(5, 7.1, 2))
    dt = 0.1
    times = np.arange(0.5, 10.1, dt)
    counts = np.linspace(5, 20, len(times))
    countrates = counts / exposures
    cfg = osim.setup(timestamps, exposures, aliasing_factor=1, extension_factor=1.0)
    return times, countrates, cfg["strategy"]


@pytest.mark.parametrize("exposures,expected", [
    (0.5, [[3, 4, 5, 6, 7], [23, 24, 25, 26, 27], [43, 44, 45, 46, 47], [63, 64, 65, 66, 67]]),
    (0.6, [[2, 3, 4, 5, 6, 7, 8], [22, 23, 24, 25, 26, 27, 28], [42, 43, 44, 45, 46, 47, 48], [62, 63, 64, 65, 66, 67, 68]]),
    (0.1, [[5], [25], [45], [65]]),
])
def test_downsampling_golden_index_sets(exposures, expected):
    """reference tests/simulator_test.py test_downsampling_{1,2,3}: exact index sets and exact means"""
    times, countrates, strategy = _downsampling_case(exposures)
    idx = osim.downsample_indices(times, strategy)
    assert [list(i) for i in idx] == expected
    truerates = [np.mean(countrates[e[0]:e[-1] + 1]) for e in expected]
    assert osim.downsample(times, countrates, strategy) == truerates        # exact equality, as assertListEqual


def test_lognormal_and_uniform_factories():
    """reference tests/stats_test.py:9-34: the factories hit the requested mean / variance"""
    for mean, std in [(10.0, 2.0), (0.5, 0.3)]:
        s, scale = osim.create_log_normal_params(mean, std)
        d = stats.lognorm(s, scale=scale)
        assert abs(d.mean() - mean) < 1e-10 and abs(d.var() - std ** 2) < 1e-10
        loc, sc = osim.create_uniform_params(mean, std)
        d = stats.uniform(loc=loc, scale=sc)
        assert abs(d.mean() - mean) < 1e-12 and abs(d.var() - std ** 2) < 1e-12


@pytest.mark.parametrize("N,B", [(0, 0.5), (3, 0.0), (3, 1.7), (7, 9.2), (14, 2.0)])
def test_kraft_median_closed_form_vs_numerical_cdf(N, B):
    """the reference takes scipy's numerical median of kraft_pdf on [0, 35] (noise_models.py:139-141)"""
    # scipy's rv_continuous does not renormalise a user _pdf on [a, b]: cdf(x) = int_a^x pdf
    cdf = lambda s: integrate.quad(osim.kraft_pdf, 0, s, args=(N, B), epsabs=1e-13, epsrel=1e-13)[0]
    med = optimize.brentq(lambda s: cdf(s) - 0.5, 0, 35, xtol=1e-12)
    assert abs(med - osim.kraft_median(N, B)) < 1e-8


@pytest.mark.parametrize("N,B", [(0, 0.5), (3, 0.0), (3, 1.7), (7, 9.2), (14, 2.0)])
def test_kbn_interval_properties(N, B):
    lo, hi = osim.kbn_interval(N, B, 0.68)
    prob = integrate.quad(osim.kraft_pdf, lo, hi, args=(N, B))[0]
    assert abs(prob - 0.68) < 1e-9
    if lo > 0:
        assert abs(osim.kraft_pdf(lo, N, B) - osim.kraft_pdf(hi, N, B)) < 1e-10      # equal-density ends = minimal length
    else:
        assert osim.kraft_pdf(0.0, N, B) >= osim.kraft_pdf(hi, N, B) - 1e-12


def test_tk95_variance_normalisation():
    """reference tests/simulator_test.py:306-339: Var(rate) -> S0 for a unit-variance DRW (tolerance 0.02)"""
    from oracle import terms
    rng = np.random.default_rng(100)
    exposures = 0.2
    times = np.arange(0, 2000, exposures)
    cfg = osim.setup(times, exposures, aliasing_factor=1, extension_factor=1.05)
    psd = lambda w: terms.psd_bendingpowerlaw(w, S0=1.0, omega0=2 * np.pi / 20)
    n = cfg["n_fft"]
    v, m = [], []
    for _ in range(60):
        re, im = rng.normal(size=(2, n // 2 + 1))
        r = osim.simulate_regular(n, cfg["sim_dt"], psd, 100.0, re, im)
        v.append(np.var(r)); m.append(np.mean(r))
    assert abs(np.mean(v) - 1.0) < 0.05 and abs(np.mean(m) - 100.0) < 1e-9


def test_e13_preserves_multiset_and_converges():
    rng = np.random.default_rng(3)
    n = 2000
    from oracle import terms
    re, im = rng.normal(size=(2, n // 2 + 1))
    lc = osim.simulate_regular(n, 1.0, lambda w: terms.psd_bendingpowerlaw(w, S0=4.0, omega0=0.05), 10.0, re, im)
    s, scale = osim.create_log_normal_params(10.0, np.std(lc))
    draws = scale * np.exp(s * rng.normal(size=n))
    out, its = osim.e13_adjust(lc, draws, max_iter=400)
    assert its < 400
    np.testing.assert_array_equal(np.sort(out), np.sort(draws))
