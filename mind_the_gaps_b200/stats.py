"""Statistics helpers (``/root/reference/mind_the_gaps/stats.py``) plus the likelihood-ratio-test
summary the reference only has as notebook cells (``docs/notebooks/tutorial_ppp.ipynb`` cell 15).

On the hot path: ``create_log_normal`` / ``create_uniform_distribution`` (E13 target PDFs,
stats.py:116-146), ``kraft_pdf`` (Kraft noise, stats.py:10-18), ``aic / aicc / bic`` (model selection,
stats.py:155-195) and the new ``lrt_statistic`` / ``lrt_pvalue``.  The periodogram chi^2 helpers of the
reference (``chi_cov``, ``chi_log_likehood``, ``chi_square``, ``fit_N``) are not on the GP path and are kept
only as thin NumPy one-liners for import compatibility.
"""
import numpy as np
from scipy import special, stats
from scipy.stats import lognorm, uniform

__all__ = ["kraft_pdf", "lognormal", "chi_cov", "chi_log_likehood", "chi_log_likehood_nonyq", "chi_square", "chi_square_N", "fit_N",
           "neg_log_like", "create_log_normal",
           "create_uniform_distribution", "aic", "aicc", "bic", "lrt_statistic", "lrt_pvalue", "lrt_summary"]


class kraft_pdf(stats.rv_continuous):
    """Posterior of the source counts given N total counts and background B (Kraft et al. 1991)."""

    def _argcheck(self, N, B):
        return (N >= 0) and (B >= 0)

    def _pdf(self, x, N, B):
        n = np.arange(N + 1)
        C = (np.sum(np.exp(-B) * B ** n / special.factorial(n))) ** -1
        return C * np.exp(-x - B) * (x + B) ** N / special.factorial(N)      # np.math is gone in numpy 2


class lognormal(stats.rv_continuous):
    def _argcheck(self, center, sigma):
        return sigma >= 0

    def _pdf(self, x, center, sigma):
        return 1 / (sigma * x * np.sqrt(2 * np.pi)) * np.exp(-(np.log(x) - center) ** 2 / (2 * sigma ** 2))


def create_log_normal(mean, std):
    """Log-normal with the requested mean and standard deviation (stats.py:116-129)."""
    var = std ** 2
    mu = np.log((mean ** 2) / np.sqrt(var + mean ** 2))
    sigma = np.sqrt(np.log(var / (mean ** 2) + 1))
    return lognorm(sigma, scale=np.exp(mu))


def create_uniform_distribution(mean, std):
    """Uniform with the requested mean and standard deviation (stats.py:132-146)."""
    var = std ** 2
    b = np.sqrt(3 * var) + mean
    a = 2 * mean - b
    return uniform(loc=a, scale=b - a)


def chi_cov(powers_data, model_powers=None, inv_cov=None):
    d = powers_data - model_powers
    return np.matmul(np.matmul(d, inv_cov), d.T)


def chi_log_likehood_nonyq(powers_data, model_pows=None):
    """Whittle statistic of Vaughan+2005 eq. A.3 / Emmanoulopoulos+2013 A.11 for chi^2_2-distributed powers (stats.py:81-91)."""
    return 2. * np.sum(np.log(model_pows) + powers_data / model_pows)


def chi_log_likehood(powers_data, model_pows=None, nyquist=False):
    """The same with the last (Nyquist, chi^2_1-distributed) power treated apart when ``nyquist`` (stats.py:59-78)."""
    if nyquist:
        ll = chi_log_likehood_nonyq(powers_data[:-1], model_pows[:-1])
        return ll + np.log(np.pi * powers_data[-1] * model_pows[-1]) + 2 * powers_data[-1] / model_pows[-1]
    return chi_log_likehood_nonyq(powers_data, model_pows)


def chi_square(powers_data, model_powers=None, sigmas=None):
    return np.sum(((model_powers - powers_data) / sigmas) ** 2)


def fit_N(loglikehood, log_like_args=()):
    """Normalisation minimising ``loglikehood(N, *log_like_args)`` (BFGS from 1; stats.py:31-41)."""
    from scipy.optimize import minimize
    return minimize(loglikehood, 1, args=log_like_args, method="BFGS").x


def chi_square_N(powers_data, model_power=None, std_power=None):
    """``chi_square`` at the best-fitting normalisation of the model (the reference's version, stats.py:108-113, passes the
    normalisation as an extra leading argument that ``chi_square`` does not take and cannot run; here the normalisation
    scales the model powers)."""
    scaled = lambda N, p, m, s: chi_square(p, np.ravel(N)[0] * m, s)
    N = fit_N(scaled, (powers_data, model_power, std_power))
    return scaled(N, powers_data, model_power, std_power)


def neg_log_like(params, y, gp):
    """``-gp.log_likelihood(y)`` at ``params`` (stats.py:149-152)."""
    gp.set_parameter_vector(params)
    return -gp.log_likelihood(y)


def bic(loglikehood, n, k):
    return -2. * loglikehood + k * np.log(n)


def aic(loglikehood, k):
    return 2 * k - 2 * loglikehood


def aicc(loglikehood, n, k):
    return aic(loglikehood, k) + 2 * k * (k + 1) / (n - k - 1)


# ---- likelihood-ratio test (tutorial_ppp.ipynb cell 15) --------------------------------------------

def lrt_statistic(loglike_null, loglike_alt):
    """``T = -2 (L_null - L_alt)``"""
    return -2.0 * (np.asarray(loglike_null, dtype=np.float64) - np.asarray(loglike_alt, dtype=np.float64))


def lrt_pvalue(T_dist, T_obs):
    """``1 - percentileofscore(T_dist, T_obs) / 100`` with scipy's default ``kind="rank"``."""
    T = np.asarray(T_dist, dtype=np.float64)
    T = T[np.isfinite(T)]
    n = len(T)
    if n == 0:
        return np.nan
    left = np.count_nonzero(T < T_obs)
    right = np.count_nonzero(T <= T_obs)
    perc = (left + right + (1 if right > left else 0)) * 50.0 / n
    return 1.0 - perc / 100.0


def lrt_summary(T_dist, T_obs, bins=10, levels=(95.0, 99.7)):
    """Histogram, p-value and the percentile lines drawn in the notebook."""
    T = np.asarray(T_dist, dtype=np.float64)
    T = T[np.isfinite(T)]
    hist, edges = np.histogram(T, bins=bins)
    return {"T_obs": float(T_obs), "p_value": lrt_pvalue(T, T_obs), "hist": hist, "edges": edges,
            "percentiles": {lv: float(np.percentile(T, lv)) for lv in levels}, "n": int(len(T))}
