"""Batched maximum-likelihood fitting of GP models on the GPU (north_star: "fitting.py (batched
optimiser)").

In the reference, the GP optimiser is ``GPModelling.fit``: one
``scipy.optimize.minimize(self._neg_log_like, x0, method="L-BFGS-B", bounds=gp.get_parameter_bounds())``
call per lightcurve with finite-difference gradients
(``/root/reference/mind_the_gaps/gpmodelling.py:172-194``).  This module fits MANY lightcurves x MANY starting
points in lock step with the device L-BFGS of ``csrc/mtg_lbfgs.cu`` driven by the analytic gradient kernel.  The
reference's own ``fitting.py`` (periodogram power-law fits, done there with lmfit) is restated at
the end of the file with the same function names and returns.
"""
import ctypes

import numpy as np
from scipy.optimize import OptimizeResult

from . import _native

__all__ = ["lbfgs_batch", "fit_lightcurves", "multistart_points", "scipy_lbfgsb", "chose_ls_model", "chose_frequency_range", "brokenpl",
           "linear_fit_break", "linear_fit", "s_statistic", "s_stat_powerlaw", "minimize_powerlaw", "fit_lomb_scargle", "fit_psd_powerlaw"]

_MESSAGES = {0: "CONVERGENCE: projected gradient / relative reduction below tolerance",
             1: "STOP: iteration limit reached",
             2: "ABNORMAL: line search failed or non-finite objective"}


def lbfgs_batch(model, t, Y, DY, x0, group, ds_mean=None, max_iter=200, history=10, ftol=2.220446049250313e-09,
                gtol=1e-5, max_ls=20, apply_prior=False):
    """Minimise ``-loglike`` for ``B = len(x0)`` problems; problem ``b`` uses lightcurve ``b // group``.

    ``model``: ``mtg_model_t`` (``GP.device_model()``); ``Y, DY``: ``(S, N)`` rates and 1-sigma errors
    (the diagonal is ``(DY + 1e-12)**2`` as in ``gpmodelling.py:54``); ``ds_mean``: per-lightcurve frozen
    mean (``None``: the mean comes from the model parameters).  Returns a dict of arrays
    ``x (B, ndim), fun (B,), nit (B,), status (B,)``.
    """
    _native.require_device()
    t = np.ascontiguousarray(t, dtype=np.float64)
    Y = np.ascontiguousarray(np.atleast_2d(Y), dtype=np.float64)
    diag = np.ascontiguousarray((np.atleast_2d(np.asarray(DY, dtype=np.float64)) + 1e-12) ** 2)
    x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
    S, N = Y.shape
    B, nd = x0.shape
    if diag.shape != Y.shape or len(t) != N:
        raise ValueError("dimension mismatch")
    if nd != model.n_free:
        raise ValueError("x0 has %d columns, the model has %d free parameters" % (nd, model.n_free))
    flags = _native.FLAG_APPLY_PRIOR if apply_prior else 0
    means = None
    if ds_mean is not None:
        means = np.ascontiguousarray(np.broadcast_to(ds_mean, (S,)), dtype=np.float64)
        flags |= _native.FLAG_DS_MEAN
    opts = _native.MtgLbfgsOpts(int(max_iter), int(history), float(ftol), float(gtol), int(max_ls), 0)
    x = np.empty((B, nd)); fun = np.empty(B); nit = np.empty(B, dtype=np.int32); status = np.empty(B, dtype=np.int32)
    rc = _native.lib().mtg_lbfgs_host(ctypes.byref(model), N, S, int(group), B, _native.ptr(t), _native.ptr(Y),
                                      _native.ptr(diag), _native.ptr(means), _native.ptr(x0), ctypes.byref(opts),
                                      _native.ptr(x), _native.ptr(fun), _native.ptr(nit), _native.ptr(status), flags)
    _native.check(rc)
    return {"x": x, "fun": fun, "nit": nit, "status": status}


def multistart_points(x_center, bounds, n_starts, rng, percent=0.1, spread=None):
    """``n_starts`` start points per problem: the centre itself plus Gaussian balls around it
    (same recipe as ``GPModelling.spread_walkers``, gpmodelling.py:289-350), clipped to the box.
    ``spread``: optional per-parameter absolute sigma overriding ``percent * |x|``."""
    x_center = np.atleast_2d(np.asarray(x_center, dtype=np.float64))
    P, nd = x_center.shape
    lo = np.array([-np.inf if b[0] is None else b[0] for b in bounds], dtype=np.float64)
    hi = np.array([np.inf if b[1] is None else b[1] for b in bounds], dtype=np.float64)
    sig = np.abs(x_center) * percent if spread is None else np.broadcast_to(np.asarray(spread, dtype=np.float64), (P, nd))
    out = np.repeat(x_center[:, None, :], n_starts, axis=1)
    if n_starts > 1:
        out[:, 1:, :] += rng.normal(size=(P, n_starts - 1, nd)) * np.asarray(sig)[:, None, :]
    return np.clip(out, lo, hi).reshape(P * n_starts, nd)


def _forward_steps(x, lo, hi, h0=1e-8):
    """Absolute finite-difference steps as scipy's ``approx_derivative(method="2-point", abs_step=1e-8, bounds=...)``
    takes them for L-BFGS-B without ``jac`` (``optimize._numdiff._adjust_scheme_to_bounds``): forward by ``h0``, backward
    where the forward point leaves the box and the backward one fits, otherwise towards the farther bound."""
    h = np.full(len(x), h0)
    up, dn = hi - x, x - lo
    fwd_ok, bwd_ok = up >= h0, dn >= h0
    flip = ~fwd_ok & bwd_ok
    h[flip] = -h0
    neither = ~fwd_ok & ~bwd_ok
    h[neither & (up >= dn)] = up[neither & (up >= dn)]
    h[neither & (up < dn)] = -dn[neither & (up < dn)]
    return h


def scipy_lbfgsb(gp, y, x0, bounds=None, **options):
    """``scipy.optimize.minimize(neg_log_like, x0, method="L-BFGS-B", bounds=...)`` exactly as ``GPModelling.fit`` calls
    it in the reference (gpmodelling.py:192-193: no ``jac``, so scipy differentiates by 2-point finite differences with
    an absolute step of 1e-8) -- with the ``ndim + 1`` likelihood evaluations of every function-and-gradient request
    done in ONE batched device launch (the parallel-in-time kernel K9 for the narrow kernels: ~20 us) instead of
    ``ndim + 1`` celerite calls.  Same optimiser, same finite-difference scheme, same ``OptimizeResult``.

    ``gp``: a computed ``mind_the_gaps_b200.celerite.GP``; bounds default to ``gp.get_parameter_bounds()``."""
    from scipy.optimize import minimize
    if bounds is None:
        bounds = gp.get_parameter_bounds()
    lo = np.array([-np.inf if b[0] is None else b[0] for b in bounds], dtype=np.float64)
    hi = np.array([np.inf if b[1] is None else b[1] for b in bounds], dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    nd = len(lo)

    def fun_and_grad(x):
        x = np.clip(np.asarray(x, dtype=np.float64), lo, hi)
        h = _forward_steps(x, lo, hi)
        X = np.repeat(x[None, :], nd + 1, axis=0)
        X[1:][np.diag_indices(nd)] += h
        hs = X[1:][np.diag_indices(nd)] - x                      # the step actually representable
        ll = gp.log_likelihood_batch(y, X)
        f = -ll[0]
        if not np.isfinite(f):
            return 1e25, np.zeros(nd)                            # celerite would raise LinAlgError here
        with np.errstate(invalid="ignore", divide="ignore"):
            g = -(ll[1:] - ll[0]) / hs
        g[~np.isfinite(g)] = 0.0
        return f, g

    return minimize(fun_and_grad, np.asarray(x0, dtype=np.float64), jac=True, method="L-BFGS-B", bounds=list(zip(lo, hi)),
                    options=options or None)


def fit_lightcurves(gp, t, Y, DY, x0=None, n_starts=1, rng=None, frozen_mean="lightcurve", spread=None, **opts):
    """MLE-fit the kernel of ``gp`` to each row of ``Y`` (``S`` lightcurves on the common times ``t``).

    ``frozen_mean="lightcurve"`` reproduces ``GPModelling(lc, kernel)`` with ``mean_model=None``: the mean
    is frozen at each lightcurve's own ``mean(y)`` (gpmodelling.py:83-87).  With ``n_starts > 1`` every
    lightcurve is started from ``x0`` and ``n_starts - 1`` perturbed copies and the best optimum kept
    (the reference guards against local maxima with a short MCMC instead, tutorial_ppp.ipynb cell 13).
    Returns a list of ``OptimizeResult`` (``x, fun, nit, status, success, message``) plus, under
    ``.all_fun``, the optimum of every start.
    """
    Y = np.atleast_2d(np.asarray(Y, dtype=np.float64))
    S = Y.shape[0]
    model = gp.device_model()
    nd = model.n_free
    if x0 is None:
        x0 = gp.get_parameter_vector()
    x0 = np.asarray(x0, dtype=np.float64)
    if x0.ndim == 1:
        x0 = np.tile(x0, (S, 1))
    if x0.shape != (S, nd):
        raise ValueError("x0 must have shape (%d, %d)" % (S, nd))
    rng = np.random.default_rng() if rng is None else rng
    starts = multistart_points(x0, gp.get_parameter_bounds(), n_starts, rng, spread=spread)
    ds_mean = Y.mean(axis=1) if (frozen_mean == "lightcurve" and not np.any(gp.mean.unfrozen_mask)) else None
    res = lbfgs_batch(model, t, Y, DY, starts, n_starts, ds_mean=ds_mean, **opts)
    fun = res["fun"].reshape(S, n_starts)
    fun = np.where(np.isfinite(fun), fun, np.inf)
    best = np.argmin(fun, axis=1)
    idx = np.arange(S) * n_starts + best
    out = []
    for s in range(S):
        i = idx[s]
        st = int(res["status"][i])
        out.append(OptimizeResult(x=res["x"][i].copy(), fun=float(res["fun"][i]), nit=int(res["nit"][i]), status=st,
                                  success=(st == 0), message=_MESSAGES.get(st, "?"), all_fun=fun[s].copy()))
    return out


# ---- periodogram fitting (the reference's own ``mind_the_gaps/fitting.py``: same names, arguments and returns) --------
# The reference does these small fits with lmfit (not a dependency here): ``linear_fit`` / ``linear_fit_break`` are
# restated as bounded least squares with scipy, returning an object with the lmfit attributes the reference's callers
# read (``params[name].value / .stderr``, ``best_fit``, ``bic``, ``nvarys``, ``errorbars``, ``chisqr``).

class _Value(float):
    """A float that also answers ``.value`` (the reference hands lmfit ``Parameter`` objects to its callers)."""
    stderr = None

    @property
    def value(self):
        return float(self)


class _Params(dict):
    def get(self, key, default=None):
        return dict.get(self, key, default)


class PeriodogramFit(object):
    def __init__(self, names, values, stderr, residual, best_fit):
        self.params = _Params()
        for n, v, e in zip(names, values, stderr):
            val = _Value(v); val.stderr = None if e is None or not np.isfinite(e) else float(e)
            self.params[n] = val
        self.nvarys = len(names)
        self.ndata = len(residual)
        self.chisqr = float(np.sum(residual ** 2))
        self.redchi = self.chisqr / max(self.ndata - self.nvarys, 1)
        self.best_fit = best_fit
        self.errorbars = all(e is not None and np.isfinite(e) for e in stderr)
        _neg2ll = self.ndata * np.log(max(self.chisqr, 1e-250) / self.ndata)       # lmfit's definitions
        self.aic = _neg2ll + 2 * self.nvarys
        self.bic = _neg2ll + np.log(self.ndata) * self.nvarys


def _plain(x):
    return np.asarray(x.value if hasattr(x, "unit") else x, dtype=np.float64)


def _bounded_least_squares(fun, x0, lo, hi, names, x, y):
    from scipy.optimize import least_squares
    res = least_squares(lambda p: fun(x, *p) - y, x0, bounds=(lo, hi))
    r = res.fun
    dof = max(len(y) - len(x0), 1)
    try:
        cov = np.linalg.inv(res.jac.T @ res.jac) * (np.sum(r ** 2) / dof)
        err = np.sqrt(np.diag(cov))
    except np.linalg.LinAlgError:
        err = [None] * len(x0)
    return PeriodogramFit(names, res.x, err, r, fun(x, *res.x))


def brokenpl(x, intercept, slope, x_break):
    """Broken power law (in log-log) that flattens above the break (fitting.py:39-47)."""
    x = np.asarray(x, dtype=np.float64)
    return np.where(x < x_break, intercept + slope * (x - x_break), intercept)


def linear_fit_break(frequencies, powers):
    """Straight line with a break in ln-ln space (fitting.py:50-68): slope in [-4, 0], break inside the sampled range."""
    log_f, log_p = np.log(_plain(frequencies)), np.log(_plain(powers))
    return _bounded_least_squares(brokenpl, [1.0, -1.0, np.mean(log_f)], [-np.inf, -4.0, np.min(log_f)], [np.inf, 0.0, np.max(log_f)],
                                  ["intercept", "slope", "x_break"], log_f, log_p)


def linear_fit(frequencies, powers):
    """Straight line in log10-log10 space (fitting.py:71-89): ``pow_slope`` in [-4, 0.5], ``pow_intercept`` free."""
    log_f, log_p = np.log10(_plain(frequencies)), np.log10(_plain(powers))
    slope0, icpt0 = np.polyfit(log_f, log_p, 1) if len(log_f) > 1 else (-1.0, np.mean(log_p))
    return _bounded_least_squares(lambda x, m, c: m * x + c, [min(max(slope0, -4.0), 0.5), icpt0], [-4.0, -np.inf], [0.5, np.inf],
                                  ["pow_slope", "pow_intercept"], log_f, log_p)


def chose_ls_model(frequencies, powers):
    """Line with or without a break, by BIC (fitting.py:12-20)."""
    outfit_break = linear_fit_break(frequencies, powers)
    outfit = linear_fit(frequencies, powers)
    if outfit_break.bic < outfit.bic:
        print("Best fit model: Linear Fit with Break")
        return outfit_break
    print("Best fit model: Linear Fit")
    return outfit


def chose_frequency_range(frequencies, power):
    """High-frequency end of the usable range = the break of the broken power law, when one is preferred (fitting.py:23-36)."""
    best_model = chose_ls_model(frequencies, power)
    if best_model.nvarys == 3:
        maximum_frequency = float(np.exp(best_model.params.get("x_break")))
        print("Break found in the LS: restricting high frequency range to %.3g (frequency units)" % maximum_frequency)
    else:
        print("No break found in the LS: High frequency range won't be modified.")
        maximum_frequency = float(np.max(_plain(frequencies)))
    return best_model, maximum_frequency


def s_statistic(observed_powers, model_powers):
    """Whittle statistic ``S = sum(ln M + P / M)`` -- Vaughan+2003 eq. A.3 (fitting.py:92-95)."""
    observed_powers = np.asarray(observed_powers, dtype=np.float64); model_powers = np.asarray(model_powers, dtype=np.float64)
    return np.sum(np.log(model_powers) + observed_powers / model_powers)


def s_stat_powerlaw(x, frequencies, observed_powers):
    """``S`` of the power law ``x[0] f^x[1]`` (fitting.py:98-101)."""
    return s_statistic(observed_powers, x[0] * np.asarray(frequencies, dtype=np.float64) ** x[1])


def minimize_powerlaw(frequencies, observed_powers):
    """Power-law fit by minimising ``S`` (fitting.py:104-108): norm >= 0, slope in [-4, 0]."""
    from scipy.optimize import minimize
    res = minimize(s_stat_powerlaw, [np.mean(observed_powers), -1], args=(frequencies, observed_powers), bounds=((0, np.inf), (-4, 0)),
                   method="L-BFGS-B")
    return res.x


def fit_lomb_scargle(frequencies, powers, sigma=1):
    """Power-law fit of a Lomb-Scargle periodogram = straight line in log-log (fitting.py:111-137).  The reference takes
    the uncertainties from lmfit's profile-likelihood ``conf_interval``; here they are ``sigma`` standard errors of the
    least-squares fit (the same for a linear model with Gaussian residuals)."""
    out_fit = linear_fit(frequencies, powers)
    psd_slope = out_fit.params.get("pow_slope")
    icpt = out_fit.params.get("pow_intercept")
    psd_norm = 10 ** icpt
    if out_fit.errorbars and sigma > 0:
        psd_slope_err = abs(2 * sigma * psd_slope.stderr)
        psd_norm_err = abs(10 ** (icpt + sigma * icpt.stderr) - 10 ** (icpt - sigma * icpt.stderr))
    else:
        psd_slope_err = 0
        psd_norm_err = 0
    return psd_slope, psd_slope_err, psd_norm, psd_norm_err


def fit_psd_powerlaw(frequencies, powers):
    """Power-law fit of a periodogram after Vaughan (2005): least squares in log10-log10 space, the ``+0.25068`` bias
    correction of the normalisation and the analytic uncertainties of that paper (fitting.py:140-178).  The zero
    frequency (and the Nyquist frequency of even data sets) must have been removed.  Returns ``(psd_slope, psd_slope_err,
    psd_norm, psd_norm_err)``; ``psd_slope`` answers ``.value`` like the lmfit parameter the reference returns."""
    out_fit = linear_fit(frequencies, powers)
    n_prime = len(powers)
    sigma_2 = np.pi ** 2 / (6 * np.log(10) ** 2)
    log_f = np.log10(_plain(frequencies))
    log_f_2_sum = np.sum(np.power(log_f, 2))
    delta = n_prime * log_f_2_sum - np.sum(log_f) ** 2
    psd_slope_err = np.sqrt(n_prime * sigma_2 / delta)
    psd_log10norm_err = np.sqrt(sigma_2 * log_f_2_sum / delta)
    psd_slope = out_fit.params.get("pow_slope")
    lognorm = out_fit.params.get("pow_intercept") + 0.25068
    psd_norm = 10 ** lognorm
    psd_norm_err = psd_norm - 10 ** (lognorm - psd_log10norm_err)
    return psd_slope, psd_slope_err, psd_norm, psd_norm_err
