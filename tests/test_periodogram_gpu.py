"""GPU: batched periodograms (csrc/mtg_periodogram.cu) against numpy / the oracle's Lomb-Scargle (astropy's algorithm
restated and pinned on scipy.signal.lombscargle in tests/test_oracle.py)."""
import numpy as np
import pytest

from mind_the_gaps_b200 import periodogram
from mind_the_gaps_b200.fitting import fit_lomb_scargle, fit_psd_powerlaw
from oracle import periodogram as op

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1000, 1001, 4096])
def test_power_spectrum_matches_numpy(n):
    rng = np.random.default_rng(n)
    dt = 0.5
    t = np.arange(n) * dt + dt / 2
    rates = np.cumsum(rng.normal(size=(7, n)), axis=1) + 10.0
    f, p = periodogram.power_spectrum(t, rates)
    freqs = np.fft.rfftfreq(n, dt)
    sl = slice(1, -1) if len(freqs) % 2 == 0 else slice(1, None)
    ref = np.abs(np.fft.rfft(rates, axis=1)[:, sl]) ** 2
    np.testing.assert_array_equal(f, freqs[sl])
    np.testing.assert_allclose(p, ref, rtol=1e-10, atol=1e-10 * ref.max())
    f1, p1 = periodogram.power_spectrum(t, rates[3])
    np.testing.assert_allclose(p1, ref[3], rtol=1e-10, atol=1e-10 * ref.max())


@pytest.mark.parametrize("fit_mean,center_data", [(True, True), (False, True), (False, False), (True, False)])
@pytest.mark.parametrize("normalization", ["standard", "psd", "model", "log"])
def test_lomb_scargle_matches_oracle(fit_mean, center_data, normalization):
    rng = np.random.default_rng(5)
    N, L, M = 437, 37, 150                                   # no dimension is a multiple of the 64 x 64 x 16 tiles
    t = np.sort(rng.uniform(0, 900.0, N))
    Y = 5.0 + np.sin(2 * np.pi * 0.031 * t)[None, :] * rng.uniform(0.2, 2.0, (L, 1)) + rng.normal(size=(L, N))
    f = np.linspace(0.002, 0.2, M)
    got = periodogram.lomb_scargle(t, Y, f, fit_mean=fit_mean, center_data=center_data, normalization=normalization)
    assert got.shape == (L, M)
    ref = np.array([op.lombscargle(t, Y[l], f, fit_mean, center_data, normalization) for l in range(L)])
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-12 * np.abs(ref).max())
    one = periodogram.lomb_scargle(t, Y[4], f, fit_mean=fit_mean, center_data=center_data, normalization=normalization)
    np.testing.assert_allclose(one, ref[4], rtol=1e-8, atol=1e-12 * np.abs(ref).max())


def test_lomb_scargle_bias_study_shape():
    """The notebook's experiment in one call: an ensemble of red-noise lightcurves with points removed; the LS slopes of
    the ensemble (fit_lomb_scargle) scatter around the slope of the evenly sampled periodograms."""
    from mind_the_gaps_b200.simulator import Simulator
    np.random.seed(27)
    dt, n = 1.0, 1000
    t = np.arange(n) * dt + dt / 2
    sim = Simulator(lambda w: np.asarray(w, dtype=float) ** -1.0, t, dt, 0, aliasing_factor=1, extension_factor=1.05)
    rates = sim.generate_lightcurves(64)
    f, p = periodogram.power_spectrum(t, rates)
    even = np.array([float(fit_psd_powerlaw(f, pl)[0]) for pl in p])
    keep = np.sort(np.random.default_rng(1).choice(np.arange(1, n - 1), n - 2 - 300, replace=False))
    keep = np.r_[0, keep, n - 1]
    ls = periodogram.lomb_scargle(t[keep], rates[:, keep], f, fit_mean=True, center_data=True, normalization="psd")
    uneven = np.array([float(fit_lomb_scargle(f, pl, sigma=0)[0]) for pl in ls])
    assert abs(np.mean(even) + 1.0) < 0.1
    assert np.all(np.isfinite(uneven)) and -1.5 < np.mean(uneven) < -0.3        # flatter: power leaks to high frequencies


@pytest.mark.parametrize("n_t", [1, 500, 4096, 20001])
def test_device_autocorrelation_matches_emcee_function_1d(n_t):
    """mtg_autocorr_host (the convergence diagnostic of the device sampler) against the host restatement of emcee's
    function_1d, and the integrated autocorrelation times that follow from both."""
    from mind_the_gaps_b200 import sampler
    rng = np.random.default_rng(n_t)
    x = np.empty((n_t, 6, 3))
    x[0] = rng.normal(size=(6, 3))
    rho = np.array([0.5, 0.9, 0.99])
    for i in range(1, n_t):                                # AR(1) chains with different correlation lengths per dimension
        x[i] = rho * x[i - 1] + np.sqrt(1 - rho ** 2) * rng.normal(size=(6, 3))
    if n_t > 1:
        dev = sampler._mean_acf_device(x); host = sampler._mean_acf(x)
        np.testing.assert_allclose(dev, host, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(sampler.integrated_time(x, tol=0, acf=sampler._mean_acf_device), sampler.integrated_time(x, tol=0), rtol=1e-9)
    else:
        assert np.all(np.isnan(sampler._mean_acf_device(x)) | (sampler._mean_acf_device(x) == 1.0))
