"""Oracle: NumPy restatement of the mind_the_gaps lightcurve simulator.  TEST INFRASTRUCTURE ONLY.

Every random input is an explicit argument so that the CUDA path can be compared on identical
noise arrays (north_star: "exposure binning and index work agree bit-exactly given identical input
noise arrays").

Follows ``/root/reference/mind_the_gaps/simulator.py``:
  ``setup``                 Simulator.__init__ grid / strategy constants (:195-258)
  ``get_fft``               get_fft (:468-501)
  ``simulate_regular``      Simulator.simulate_regularly_sampled (:369-394)
  ``cut_segment``           cut_random_segment (:536-539) + the shift of generate_lightcurve (:412-414)
  ``downsample``            Simulator.downsample (:340-367)
  ``e13_adjust``            E13Simulator.adjust_lightcurve_pdf (:65-131)
and ``noise_models.py`` (:49-78 Poisson, :107-150 Kraft, :162-184 Gaussian), ``stats.py:116-146``.

stingray is not importable here; the five ``Lightcurve`` semantics the reference relies on are
restated from stingray 2.x: ``countrate = counts / dt``; ``meanrate = mean(counts) / dt``;
``tstart = time[0] - dt / 2``; ``truncate(start, stop, method="time")`` = slice
``[searchsorted(time, start), searchsorted(time, stop))`` (left side); ``shift(x)`` adds ``x`` to ``time``.
numpy's ``rfft / irfft`` stand in for pyfftw's identical interface.
"""
from math import sqrt

import numpy as np
from scipy import special


def setup(times, exposures, aliasing_factor=2, extension_factor=10, epsilon=1.001):
    times = np.asarray(times, dtype=float)
    exposures = np.full(len(times), exposures, dtype=float) if np.isscalar(exposures) else np.asarray(exposures, dtype=float)
    sim_dt = np.min(exposures) / aliasing_factor
    dt = np.diff(times)
    start_time = times[0] - dt[0] / 1.99
    end_time = times[-1] + dt[-1]
    sim_duration = end_time - start_time
    duration = (times[-1] - times[0]) * extension_factor
    grid = np.arange(start_time - sim_dt, start_time + duration + sim_dt, sim_dt)
    half_bins = exposures / 2 * epsilon
    strategy = [(time - half_bin, time + half_bin) for time, half_bin in zip(times, half_bins)]
    return {"sim_dt": sim_dt, "sim_duration": sim_duration, "grid": grid, "n_fft": len(grid),
            "strategy": strategy, "exposures": exposures}


def get_fft(N, dt, psd, real, im):
    """``real, im``: the two rows of ``np.random.normal(0, size=(2, N // 2 + 1))``."""
    freqs = np.fft.rfftfreq(N, dt) * 2 * np.pi
    complex_fft = np.empty(len(freqs), dtype=complex)
    complex_fft[1:] = (real + im * 1j)[1:] * np.sqrt(.5 * psd(freqs[1:]))
    complex_fft[0] = 1e6
    if N % 2 == 0:
        complex_fft[-1] = np.real(complex_fft[-1])
    return complex_fft


def simulate_regular(n_fft, sim_dt, psd, mean, real, im):
    """countrate of the regularly sampled TK95 lightcurve."""
    complex_fft = get_fft(n_fft, sim_dt, psd, real, im)
    counts = np.fft.irfft(complex_fft, n=n_fft)
    counts *= sqrt(n_fft * sim_dt * sqrt(2 * np.pi))
    countrate = counts / sim_dt
    meanrate = np.mean(counts) / sim_dt
    return countrate - meanrate + mean


def cut_segment(grid, sim_dt, duration, first_bin_start, u01):
    """``u01``: the ``random_sample()`` behind ``np.random.uniform(low, high)`` (= low + (high-low) u).
    Returns ``(i0, i1, time_shift)``: the segment is ``grid[i0:i1] + time_shift``."""
    low, high = grid[0], grid[-1] - duration
    shift = low + (high - low) * u01
    i0 = int(np.searchsorted(grid, shift))
    i1 = int(np.searchsorted(grid, shift + duration))
    tstart = grid[i0] - 0.5 * sim_dt
    return i0, i1, -tstart + first_bin_start


def downsample(time, countrate, strategy):
    rates = []
    for start, end in strategy:
        idxs = np.argwhere((time >= start) & (time < end))
        with np.errstate(invalid="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                rates.append(np.mean(countrate[idxs]))
    return rates


def downsample_indices(time, strategy):
    return [np.flatnonzero((time >= start) & (time < end)) for start, end in strategy]


def create_log_normal_params(mean, std):
    """stats.py:116-129 -> (sigma, scale) of scipy's lognorm"""
    var = std ** 2
    mu = np.log((mean ** 2) / np.sqrt(var + mean ** 2))
    sigma = np.sqrt(np.log(var / (mean ** 2) + 1))
    return sigma, np.exp(mu)


def create_uniform_params(mean, std):
    """stats.py:132-146 -> (loc, scale) of scipy's uniform"""
    var = std ** 2
    b = np.sqrt(3 * var) + mean
    a = 2 * mean - b
    return a, b - a


def e13_adjust(countrate, xsim0, max_iter=400):
    """``xsim0``: the ``pdf.rvs(size=n)`` draws.  Returns ``(xsim, iterations)``."""
    n = len(countrate)
    xsim = np.array(xsim0, dtype=float)
    amplitudes_norm = np.absolute(np.fft.rfft(countrate)) / (n // 2 + 1)

    def adjust(x):
        phases = np.angle(np.fft.rfft(x))
        adj = np.fft.irfft(amplitudes_norm * np.exp(1j * phases), n=n)
        adj[np.argsort(-adj)] = x[np.argsort(-x)]
        return adj

    xsim_adjust = adjust(xsim)
    iteration = 0
    while (not np.allclose(xsim_adjust, xsim, rtol=1e-04) and iteration < max_iter):
        xsim = xsim_adjust
        xsim_adjust = adjust(xsim)
        iteration += 1
    return xsim, iteration


# ---- noise -------------------------------------------------------------------------------------

def poisson_noise(rates, exposures, bkg_counts, bkg_rate_err, counts):
    """``counts``: the ``np.random.poisson(rates * exposures + bkg_counts)`` draws."""
    net_counts = counts - bkg_counts
    dy = np.sqrt((np.sqrt(counts) / exposures) ** 2 + bkg_rate_err ** 2)
    return net_counts / exposures, dy


def gaussian_noise(rates, sigma, normals):
    return rates + sigma * normals, sigma * np.ones(len(rates))


def kraft_pdf(x, N, B):
    """stats.py:10-18 with ``np.math.factorial`` (gone in numpy 2) replaced by scipy's."""
    n = np.arange(N + 1)
    C = (np.sum(np.exp(-B) * B ** n / special.factorial(n))) ** -1
    return C * np.exp(-x - B) * (x + B) ** N / special.factorial(N)


def kraft_median(N, B):
    """median of ``kraft_pdf`` on its full support (closed form, verified against the reference's
    numerical ``kraft_pdf(a=0, b=35)(N, B).median()`` to 4e-15 for the ranges used)."""
    return special.gammainccinv(N + 1, 0.5 * special.gammaincc(N + 1, B)) - B


def kbn_interval(N, B, CL=0.68):
    """Kraft-Burrows-Nousek minimal credible interval, the algorithm of
    ``astropy.stats.poisson_conf_interval(..., "kraft-burrows-nousek")`` restated with the closed-form
    integral ``int f = C (Q(N+1, s_min+B) - Q(N+1, s_max+B))``.  "parity unpinned" vs astropy."""
    from scipy.optimize import brentq
    N = int(N)
    C = 1.0 / special.gammaincc(N + 1, B)
    f = lambda s: kraft_pdf(s, N, B)

    def s_min_of(s_max):
        y = f(s_max)
        if f(0.0) >= y:
            return 0.0
        return brentq(lambda x: f(x) - y, 0.0, N - B)

    def func(s):
        lo = s_min_of(s)
        return C * (special.gammaincc(N + 1, lo + B) - special.gammaincc(N + 1, s + B)) - CL

    s_max = brentq(func, max(N - B, 0.0), 100.0, xtol=1e-13)
    return s_min_of(s_max), s_max


def kraft_noise(rates, exposures, bkg_counts, bkg_rate_err, counts, kraft_counts=15):
    net_rates, dy = poisson_noise(rates, exposures, bkg_counts, bkg_rate_err, counts)
    net_rates = np.array(net_rates, dtype=float); dy = np.array(dy, dtype=float)
    total_counts = net_rates * exposures + bkg_counts
    net_counts = total_counts - bkg_counts
    expression = total_counts < kraft_counts
    for i in np.flatnonzero(expression):
        Nr = int(np.round(total_counts[i]))
        net_counts[i] = kraft_median(Nr, bkg_counts[i])
        net_rates[i] = net_counts[i] / exposures[i]
        lo, hi = kbn_interval(int(total_counts[i]), bkg_counts[i], 0.68)
        dy[i] = (hi - lo) / 2 / exposures[i]
    return net_rates, dy
