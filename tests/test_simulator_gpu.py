"""GPU parity for the simulator kernels (K3/K5/K6/K7) against the NumPy oracle on IDENTICAL noise
arrays: index work bit-exact, floating point to FFT rounding (north_star)."""
import numpy as np
import pytest

from mind_the_gaps_b200.lightcurves import GappyLightcurve, Lightcurve
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from mind_the_gaps_b200.models.psd_models import BendingPowerlaw
from mind_the_gaps_b200.noise_models import GaussianNoise, KraftNoise, PoissonNoise
from mind_the_gaps_b200.simulator import E13Simulator, Simulator
from oracle import simulator as osim
from oracle import terms as oterms

from helpers import window

pytestmark = pytest.mark.gpu


# ---- exposure binning: the reference's golden vectors (tests/simulator_test.py:192-253) ---------------

@pytest.mark.parametrize("exposures,expected", [
    (0.5, [[3, 4, 5, 6, 7], [23, 24, 25, 26, 27], [43, 44, 45, 46, 47], [63, 64, 65, 66, 67]]),
    (0.6, [[2, 3, 4, 5, 6, 7, 8], [22, 23, 24, 25, 26, 27, 28], [42, 43, 44, 45, 46, 47, 48], [62, 63, 64, 65, 66, 67, 68]]),
    (0.1, [[5], [25], [45], [65]]),
])
def test_downsampling_golden(exposures, expected):
    timestamps = np.append(np.arange(1, 3.1, 2), np.arange(5, 7.1, 2))
    dt = 0.1
    times = np.arange(0.5, 10.1, dt)
    countrates = np.linspace(5, 20, len(times)) / exposures
    lc = Lightcurve(times, countrates, input_counts=False)
    simu = Simulator(lambda w: 10 * w ** -2.0, timestamps, exposures, 0, extension_factor=1.0, aliasing_factor=1)
    truerates = [np.mean(countrates[e[0]:e[-1] + 1]) for e in expected]
    assert simu.downsample(lc) == truerates                      # exact, as the reference's assertListEqual
    first, count = simu.downsample_indices(lc)
    assert [list(range(f, f + c)) for f, c in zip(first, count)] == expected


def test_downsampling_matches_numpy_mean_bit_exactly_for_all_bin_lengths():
    """np.mean's pairwise summation order reproduced for 1..400 samples per bin (and an empty bin)."""
    rng = np.random.default_rng(0)
    lengths = list(range(1, 140)) + [255, 256, 257, 300, 400]
    dt = 1.0
    edges = np.concatenate([[0], np.cumsum(np.array(lengths) + 3)])
    n = int(edges[-1]) + 10
    time = np.arange(n) * dt
    rate = rng.normal(100.0, 30.0, n) * 10.0 ** rng.integers(-3, 4, n)
    centers = np.array([edges[i] + 1 + lengths[i] / 2.0 - 0.5 for i in range(len(lengths))])
    exposures = np.array(lengths, dtype=float) * dt
    simu = Simulator(lambda w: w ** -2.0, centers, exposures, 0, extension_factor=1.0, aliasing_factor=1, epsilon=1.0)
    lc = Lightcurve(time, rate, input_counts=False, dt=dt)
    got = simu.downsample(lc)
    ref = osim.downsample(time, rate, simu.strategy)
    assert got == ref
    counts = simu.downsample_indices(lc)[1]
    assert list(counts) == [len(i) for i in osim.downsample_indices(time, simu.strategy)]


# ---- TK95 + segment + binning on injected noise -----------------------------------------------------------

def _gapped_setup(N=60, seed=3):
    t, exposures = window(N, seed)
    return t, exposures


@pytest.mark.parametrize("n_lc", [1, 5])
def test_generate_matches_oracle_on_identical_noise(n_lc):
    t, exposures = _gapped_setup()
    day = 86400.0
    psd = BendingPowerlaw(S0=0.01, omega0=2 * np.pi / (20 * day))
    simu = Simulator(psd, t, exposures, 0.5, pdf="Gaussian", extension_factor=2, aliasing_factor=2)
    cfg = osim.setup(t, exposures, aliasing_factor=2, extension_factor=2)
    assert cfg["n_fft"] == simu.fftndatapoints and np.array_equal(cfg["grid"], simu.sim_timestamps)
    M = cfg["n_fft"] // 2 + 1
    rng = np.random.default_rng(7)
    normals = rng.normal(size=(n_lc, 2, M)); u = rng.random(n_lc)
    rates, info = simu.generate_lightcurves(n_lc, normals=normals, u_segment=u, return_info=True)
    for l in range(n_lc):
        full = osim.simulate_regular(cfg["n_fft"], cfg["sim_dt"], psd, 0.5, normals[l, 0], normals[l, 1])
        i0, i1, shift = osim.cut_segment(cfg["grid"], cfg["sim_dt"], simu.sim_duration, simu.strategy[0][0], u[l])
        assert (info["segment"][l, 0], info["segment"][l, 1]) == (i0, i1 - i0)            # index work: exact
        ref = osim.downsample(cfg["grid"][i0:i1] + shift, full[i0:i1], cfg["strategy"])
        np.testing.assert_allclose(rates[l], ref, rtol=1e-9, atol=1e-12)


def test_celerite_psd_on_device_equals_host_table():
    t, exposures = _gapped_setup(40, 5)
    day = 86400.0
    kernel = DampedRandomWalk(np.log(0.01), np.log(2 * np.pi / (20 * day))) + Lorentzian(np.log(0.005), np.log(30.0), np.log(2 * np.pi / (5 * day)))
    s1 = Simulator(kernel.get_psd, t, exposures, 0.5, extension_factor=2)
    s2 = Simulator(lambda w: kernel.get_psd(w), t, exposures, 0.5, extension_factor=2)       # opaque callable -> host table
    a = s1.generate_lightcurves(3, seed=11)
    b = s2.generate_lightcurves(3, seed=11)
    np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-13)
    # per-lightcurve parameters
    P = np.tile(kernel.get_parameter_vector(), (3, 1)); P[1, 0] += 1.0
    c = s1.generate_lightcurves(3, parameters=P, seed=11)
    np.testing.assert_allclose(c[0], a[0], rtol=1e-12)
    assert not np.allclose(c[1], a[1])


def test_regular_lightcurve_mean_and_variance():
    """reference tests/simulator_test.py:306-339 (tolerances 0.01 / 0.02 over 100 sims of 1e5 points)"""
    exposures = 0.2
    times = np.arange(0, 20000, exposures)
    simu = Simulator(BendingPowerlaw(S0=1.0, omega0=2 * np.pi / 200), times, exposures, 100, "Gaussian",
                     extension_factor=1.05, aliasing_factor=1)
    means, variances = [], []
    np.random.seed(100)
    for _ in range(100):
        lc = simu.simulate_regularly_sampled()
        means.append(np.mean(lc.countrate)); variances.append(np.var(lc.countrate))
    assert abs(np.mean(means) - 100) < 0.01
    assert abs(np.mean(variances) - 1.0) < 0.02


# ---- noise -----------------------------------------------------------------------------------------------------

def test_poisson_noise_bit_exact_on_injected_counts():
    rng = np.random.default_rng(1)
    N = 500
    exposures = rng.uniform(500, 2000, N); rates = rng.uniform(0.2, 0.8, N)
    bkg = rng.uniform(0, 3, N); err = rng.uniform(0, 1e-3, N)
    counts = rng.poisson(rates * exposures + bkg)
    got_r, got_dy = PoissonNoise(exposures, bkg, err).add_noise(rates, counts=counts)
    ref_r, ref_dy = osim.poisson_noise(rates, exposures, bkg, err, counts)
    assert np.array_equal(got_r, ref_r) and np.array_equal(got_dy, ref_dy)


def test_gaussian_noise_bit_exact_on_injected_normals():
    rng = np.random.default_rng(2)
    rates = rng.uniform(0.2, 0.8, 300); z = rng.normal(size=300)
    got_r, got_dy = GaussianNoise(np.ones(300), 0.05).add_noise(rates, normals=z)
    ref_r, ref_dy = osim.gaussian_noise(rates, 0.05, z)
    assert np.array_equal(got_r, ref_r) and np.array_equal(got_dy, ref_dy)


def test_kraft_noise_on_injected_counts():
    rng = np.random.default_rng(3)
    N = 400
    exposures = rng.uniform(500, 2000, N); rates = rng.uniform(0.002, 0.02, N)
    bkg = 0.001 * exposures; err = np.full(N, 1e-4)
    counts = rng.poisson(rates * exposures + bkg)
    got_r, got_dy = KraftNoise(exposures, bkg, err).add_noise(rates, counts=counts)
    ref_r, ref_dy = osim.kraft_noise(rates, exposures, bkg, err, counts)
    low = (ref_r * 0 + (((counts - bkg) / exposures) * exposures + bkg)) < 15
    assert low.sum() > 50 and (~low).sum() > 50
    assert np.array_equal(got_r[~low], ref_r[~low]) and np.array_equal(got_dy[~low], ref_dy[~low])     # Poisson branch: exact
    np.testing.assert_allclose(got_r[low], ref_r[low], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(got_dy[low], ref_dy[low], rtol=1e-9)


def test_philox_poisson_and_normal_statistics():
    N = 200000
    for lam in (0.3, 4.0, 9.9, 10.1, 37.0, 800.0):
        r, _ = PoissonNoise(np.ones(N)).add_noise(np.full(N, lam), seed=5)
        assert abs(r.mean() - lam) < 5 * np.sqrt(lam / N) and abs(r.var() - lam) < 0.03 * lam + 0.01
        assert np.all(r == np.round(r)) and r.min() >= 0
    z, _ = GaussianNoise(np.ones(N), 1.0).add_noise(np.zeros(N), seed=6)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.02 and abs(np.mean(z ** 4) - 3) < 0.1
    with pytest.raises(ValueError):
        PoissonNoise(np.ones(4)).add_noise(np.array([1.0, -1.0, 2.0, 3.0]))


# ---- E13 --------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("n", [2000, 2001, 40000])
def test_e13_adjust_matches_oracle(n):
    rng = np.random.default_rng(n)
    re, im = rng.normal(size=(2, n // 2 + 1))
    lcr = osim.simulate_regular(n, 1.0, lambda w: oterms.psd_bendingpowerlaw(w, S0=4.0, omega0=0.05), 10.0, re, im)
    s, scale = osim.create_log_normal_params(10.0, np.std(lcr))
    draws = scale * np.exp(s * rng.normal(size=n))
    ref, its = osim.e13_adjust(lcr, draws, max_iter=400)
    e13 = E13Simulator(10.0, "lognormal", max_iter=400)
    lc = Lightcurve(np.arange(n, dtype=float), lcr, input_counts=False, dt=1.0)
    out = e13.adjust_lightcurve_pdf(lc, None, max_iter=400, xsim=draws).countrate
    assert np.array_equal(np.sort(out), np.sort(draws))              # the multiset never changes
    assert abs(e13.last_iterations - its) <= 1
    assert np.mean(out == ref) > 0.995                               # identical ranks except FFT-rounding near-ties


@pytest.mark.parametrize("n", [3000, 40000])
def test_e13_stops_at_max_iter_with_the_last_accepted_iterate(n):
    """simulator.py:111-129: the loop ends after max_iter iterations even if not converged, and the result is the last
    ACCEPTED iterate (xsim, not xsim_adjust); both the batched (short series) and the per-slot (long series) loop."""
    rng = np.random.default_rng(n + 1)
    re, im = rng.normal(size=(2, n // 2 + 1))
    lcr = osim.simulate_regular(n, 1.0, lambda w: oterms.psd_bendingpowerlaw(w, S0=4.0, omega0=0.05), 10.0, re, im)
    s, scale = osim.create_log_normal_params(10.0, np.std(lcr))
    draws = scale * np.exp(s * rng.normal(size=n))
    import warnings
    for max_iter in (0, 3):
        ref, its = osim.e13_adjust(lcr, draws, max_iter=max_iter)
        assert its == max_iter
        e13 = E13Simulator(10.0, "lognormal", max_iter=max_iter)
        lc = Lightcurve(np.arange(n, dtype=float), lcr.copy(), input_counts=False, dt=1.0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = e13.adjust_lightcurve_pdf(lc, None, max_iter=max_iter, xsim=draws).countrate
        assert e13.last_iterations == max_iter
        assert np.array_equal(np.sort(out), np.sort(draws))
        assert np.mean(out == ref) > (0.999 if max_iter == 0 else 0.995)


def test_e13_rank_order_with_colliding_float_keys():
    """The rank order of an iteration comes from a radix sort on float(adj) plus a fix-up of runs of equal float keys by the
    double values.  A series of unit variance on a level of 2e6 has a float resolution of 0.25: almost every key collides
    (runs of thousands), and the result must still be the oracle's."""
    n = 40000
    rng = np.random.default_rng(5)
    re, im = rng.normal(size=(2, n // 2 + 1))
    lcr = osim.simulate_regular(n, 1.0, lambda w: oterms.psd_bendingpowerlaw(w, S0=4.0, omega0=0.05), 2.0e6, re, im)
    draws = 2.0e6 + np.std(lcr) * rng.normal(size=n)
    ref, its = osim.e13_adjust(lcr, draws, max_iter=50)
    e13 = E13Simulator(2.0e6, "gaussian", max_iter=50)
    lc = Lightcurve(np.arange(n, dtype=float), lcr, input_counts=False, dt=1.0)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = e13.adjust_lightcurve_pdf(lc, None, max_iter=50, xsim=draws).countrate
    assert np.array_equal(np.sort(out), np.sort(draws))
    assert abs(e13.last_iterations - its) <= 1
    assert np.mean(out == ref) > 0.99


def test_e13_on_the_lengthened_grid_is_the_same_process():
    """fft_length="fast" lets the PDF adjustment run on the segment lengthened to a 7-smooth length (MTG_SIM_E13_PADDED: one
    batch per chunk, no Bluestein FFT).  The binned log-normal lightcurves must be the same process as on the reference's
    grid: flux distribution, per-lightcurve mean / variance (two-sample KS), all positive, converged."""
    import scipy.stats as sps
    t, exposures = _gapped_setup(120, 9)
    day = 86400.0
    psd = BendingPowerlaw(S0=0.02, omega0=2 * np.pi / (20 * day))
    out = {}
    for mode, seed in (("exact", 3), ("fast", 4)):
        simu = Simulator(psd, t, exposures, 0.5, pdf="Lognormal", extension_factor=2, aliasing_factor=2, max_iter=400, fft_length=mode)
        rates, info = simu.generate_lightcurves(160, seed=seed, return_info=True)
        assert np.all(rates > 0) and np.all(np.isfinite(rates))
        assert np.all(info["e13_iterations"] < 400)
        out[mode] = rates
        simu.close()
    assert sps.ks_2samp(out["exact"].ravel()[::7], out["fast"].ravel()[::7]).pvalue > 1e-3
    for f in (lambda r: r.mean(axis=1), lambda r: r.var(axis=1)):
        assert sps.ks_2samp(f(out["exact"]), f(out["fast"])).pvalue > 1e-3


def test_e13_generate_lognormal_properties():
    t, exposures = _gapped_setup(50, 9)
    day = 86400.0
    psd = BendingPowerlaw(S0=0.02, omega0=2 * np.pi / (20 * day))
    simu = Simulator(psd, t, exposures, 0.5, pdf="Lognormal", extension_factor=2, aliasing_factor=2, max_iter=400)
    rates, info = simu.generate_lightcurves(6, seed=3, return_info=True)
    assert np.all(rates > 0) and np.all(np.isfinite(rates))
    assert np.all(info["e13_iterations"] < 400)
    assert abs(rates.mean() - 0.5) < 0.15
