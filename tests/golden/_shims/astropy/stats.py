"""astropy.stats.poisson_conf_interval stand-in: 'kraft-burrows-nousek' only (golden generation only).

Restates the published algorithm (Kraft, Burrows & Nousek 1991, ApJ 374, 344): the minimal-length
interval [s_min, s_max] of the posterior f(s) = C e^{-(s+B)} (s+B)^N / N! holding the confidence level,
i.e. f(s_min) = f(s_max) unless s_min = 0.  NOT astropy's code; fixtures depending on it are flagged."""
import numpy as np
from scipy import integrate, optimize, special


def _pdf(s, N, B):
    n = np.arange(N + 1)
    C = 1.0 / np.sum(np.exp(-B + n * np.log(B) - special.gammaln(n + 1))) if B > 0 else 1.0
    return C * np.exp(-(s + B) + N * np.log(s + B) - special.gammaln(N + 1)) if (s + B) > 0 else (C if N == 0 else 0.0)


def _kbn(N, B, CL):
    N = int(N); B = float(B)
    mode = max(N - B, 0.0)
    cdf = lambda a, b: integrate.quad(_pdf, a, b, args=(N, B), epsabs=1e-13, epsrel=1e-13)[0]
    # try s_min = 0 first: valid when f(0) >= f(s_max)
    smax0 = optimize.brentq(lambda s: cdf(0, s) - CL, 0, 200 + 10 * N)
    if _pdf(0.0, N, B) >= _pdf(smax0, N, B):
        return 0.0, smax0
    def smin_of(smax):
        target = _pdf(smax, N, B)
        if _pdf(0.0, N, B) >= target:
            return 0.0
        if _pdf(mode, N, B) <= target:
            return mode
        return optimize.brentq(lambda s: _pdf(s, N, B) - target, 0, mode, xtol=1e-14)
    f = lambda smax: cdf(smin_of(smax), smax) - CL
    smax = optimize.brentq(f, mode + 1e-9, 200 + 10 * N, xtol=1e-12)
    return smin_of(smax), smax


def poisson_conf_interval(n, interval="root-n", sigma=1, background=0, confidence_level=None):
    if interval != "kraft-burrows-nousek":
        raise NotImplementedError(interval)
    n = np.atleast_1d(n); background = np.broadcast_to(np.asarray(background, dtype=float), n.shape)
    out = np.array([_kbn(a, b, confidence_level) for a, b in zip(n, background)]).T.reshape(2, -1)
    return out
