"""Batched maximum-likelihood fitting of GP models on the GPU (north_star: "fitting.py (batched
optimiser)").

In the reference, the GP optimiser is ``GPModelling.fit``: one
``scipy.optimize.minimize(self._neg_log_like, x0, method="L-BFGS-B", bounds=gp.get_parameter_bounds())``
call per lightcurve with finite-difference gradients
(``/root/reference/mind_the_gaps/gpmodelling.py:172-194``); the reference's own ``fitting.py`` holds
periodogram power-law fits with lmfit and is not on the GP path.  This module
fits MANY lightcurves x MANY starting points in lock step with the device L-BFGS of
``csrc/mtg_lbfgs.cu`` driven by the analytic gradient kernel.
"""
import ctypes

import numpy as np
from scipy.optimize import OptimizeResult

from . import _native

__all__ = ["lbfgs_batch", "fit_lightcurves", "multistart_points", "scipy_lbfgsb", "s_statistic", "fit_psd_powerlaw"]

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


# ---- periodogram helpers the simulator's validation needs (reference fitting.py:92-95, 140-178) -----

def s_statistic(powers, model_powers):
    """Whittle statistic ``S = 2 sum(P/M + ln M)`` (reference fitting.py:92-95; Vaughan 2005)."""
    powers = np.asarray(powers, dtype=np.float64); model_powers = np.asarray(model_powers, dtype=np.float64)
    return 2.0 * np.sum(powers / model_powers + np.log(model_powers))


def fit_psd_powerlaw(frequencies, powers):
    """Least-squares straight line in log-log space with the ``0.25068`` bias correction of
    Vaughan (2005) (reference fitting.py:140-178, there done with lmfit).  Returns
    ``(slope, slope_err, log10_norm, log10_norm_err)``."""
    x = np.log10(np.asarray(frequencies, dtype=np.float64))
    yv = np.log10(np.asarray(powers, dtype=np.float64))
    A = np.vstack([x, np.ones_like(x)]).T
    coef, res, _, _ = np.linalg.lstsq(A, yv, rcond=None)
    n = len(x)
    dof = max(n - 2, 1)
    sigma2 = float(np.sum((yv - A @ coef) ** 2)) / dof
    cov = sigma2 * np.linalg.inv(A.T @ A)
    return float(coef[0]), float(np.sqrt(cov[0, 0])), float(coef[1] + 0.25068), float(np.sqrt(cov[1, 1]))
