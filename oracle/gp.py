"""Oracle: celerite GP log-likelihood.  TEST INFRASTRUCTURE ONLY.

Two independent evaluations of what ``celerite.GP.log_likelihood`` returns at the reference's
call sites (``/root/reference/mind_the_gaps/gpmodelling.py:51-54, 147-152, 168-169``):

``loglike_dense``      primary truth: explicit covariance matrix + dense Cholesky, O(N^3).
``loglike_recursion``  the O(N J^2) semiseparable LDL^T recursion celerite's ``CholeskySolver``
                       implements (Foreman-Mackey et al. 2017, section 5, eqs. 43-46; restated in
                       SURVEY.md section 3.1), vectorised over a batch of coefficient sets and valid
                       in ``complex128`` so that ``grad_complex_step`` differentiates it.

Conventions copied from the call sites: ``diag = (dy + 1e-12)**2`` (``gpmodelling.py:54`` passes
``yerr = dy + 1e-12`` and celerite squares it), residual ``= y - mean`` with the mean frozen at
``mean(y)`` when ``mean_model=None`` (``gpmodelling.py:83-87``), non positive-definite -> ``-inf``.

celerite is not importable here, so this module is "parity unpinned" at the celerite boundary;
``loglike_recursion`` is pinned against ``loglike_dense`` (tests/test_oracle.py).
"""
import numpy as np
import scipy.linalg

from . import terms as _terms

LOG2PI = float(np.log(2.0 * np.pi))


def loglike_dense(t, resid, diag, ar, cr, ac, bc, cc, dc, jitter=0.0):
    """-0.5 (r^T K^-1 r + log det K + N log 2 pi) with K built entry by entry."""
    t = np.asarray(t, dtype=float)
    tau = t[:, None] - t[None, :]
    K = _terms.kernel_value(tau, ar, cr, ac, bc, cc, dc)
    K[np.diag_indices_from(K)] += np.asarray(diag, dtype=float) + float(jitter)
    try:
        cf = scipy.linalg.cho_factor(K, lower=True, check_finite=True)
    except (np.linalg.LinAlgError, ValueError):
        return -np.inf
    alpha = scipy.linalg.cho_solve(cf, resid)
    logdet = 2.0 * np.sum(np.log(np.diag(cf[0])))
    return -0.5 * (float(np.dot(resid, alpha)) + logdet + len(t) * LOG2PI)


def loglike_recursion(t, resid, diag, ar, cr, ac, bc, cc, dc, jitter=0.0, return_status=False):
    """Batched semiseparable recursion.

    ``ar, cr``: ``(B, Jr)``; ``ac, bc, cc, dc``: ``(B, Jc)``; ``jitter``: scalar or ``(B,)``;
    ``resid`` and ``diag``: ``(N,)`` shared or ``(B, N)`` per series.  Returns ``(B,)`` log-likelihoods
    (``-inf`` where the factorisation meets ``D_n <= 0`` or a non-finite value).
    """
    ar = np.atleast_2d(ar); cr = np.atleast_2d(cr)
    ac = np.atleast_2d(ac); bc = np.atleast_2d(bc); cc = np.atleast_2d(cc); dc = np.atleast_2d(dc)
    B = max(ar.shape[0], ac.shape[0])
    Jr = ar.shape[1]; Jc = ac.shape[1]; J = Jr + 2 * Jc
    cplx = any(np.iscomplexobj(x) for x in (ar, cr, ac, bc, cc, dc, jitter, resid, diag))
    dt_ = np.complex128 if cplx else np.float64
    t = np.asarray(t, dtype=float)
    N = len(t)
    resid = np.asarray(resid); diag = np.asarray(diag)
    r_of = (lambda n: resid[:, n]) if resid.ndim == 2 else (lambda n: resid[n])
    d_of = (lambda n: diag[:, n]) if diag.ndim == 2 else (lambda n: diag[n])
    a_sum = np.asarray(jitter, dtype=dt_) + ar.sum(axis=1) + ac.sum(axis=1)
    a_sum = np.broadcast_to(a_sum, (B,)).astype(dt_)

    def uv(tn):
        U = np.empty((B, J), dtype=dt_); V = np.empty((B, J), dtype=dt_)
        U[:, :Jr] = ar; V[:, :Jr] = 1.0
        cd = np.cos(dc * tn); sd = np.sin(dc * tn)
        U[:, Jr::2] = ac * cd + bc * sd
        U[:, Jr + 1::2] = ac * sd - bc * cd
        V[:, Jr::2] = cd
        V[:, Jr + 1::2] = sd
        return U, V

    def phi_of(dtn):
        P = np.empty((B, J), dtype=dt_)
        P[:, :Jr] = np.exp(-cr * dtn)
        e = np.exp(-cc * dtn)
        P[:, Jr::2] = e; P[:, Jr + 1::2] = e
        return P

    bad = np.zeros(B, dtype=bool)
    U, V = uv(t[0])
    D = d_of(0) + a_sum
    bad |= ~(np.real(D) > 0)
    W = V / D[:, None]
    S = np.zeros((B, J, J), dtype=dt_)
    f = np.zeros((B, J), dtype=dt_)
    z = np.broadcast_to(np.asarray(r_of(0), dtype=dt_), (B,)).copy()
    chi2 = z * z / D
    logdet = np.log(D)
    for n in range(1, N):
        P = phi_of(t[n] - t[n - 1])
        S = (S + D[:, None, None] * W[:, :, None] * W[:, None, :]) * (P[:, :, None] * P[:, None, :])
        f = P * (f + W * z[:, None])
        U, V = uv(t[n])
        SU = np.einsum("bjk,bk->bj", S, U)
        D = d_of(n) + a_sum - np.einsum("bj,bj->b", U, SU)
        bad |= ~(np.real(D) > 0)
        W = (V - SU) / D[:, None]
        z = r_of(n) - np.einsum("bj,bj->b", U, f)
        chi2 = chi2 + z * z / D
        with np.errstate(invalid="ignore", divide="ignore"):
            logdet = logdet + np.log(D)
    ll = -0.5 * (chi2 + logdet + N * LOG2PI)
    bad |= ~np.isfinite(np.real(ll))
    ll = np.where(bad, -np.inf, ll)
    if return_status:
        return ll, bad.astype(np.int32)
    return ll


def log_prior(params, bounds):
    """celerite ``Model.log_prior``: 0 inside the closed box, -inf outside (``None`` = unbounded)."""
    p = np.atleast_2d(np.asarray(params, dtype=float))
    lp = np.zeros(p.shape[0])
    for k, (lo, hi) in enumerate(bounds):
        if lo is not None:
            lp = np.where(p[:, k] < lo, -np.inf, lp)
        if hi is not None:
            lp = np.where(p[:, k] > hi, -np.inf, lp)
    return lp


def loglike_params(spec, params, t, y, dy, mean=None, fit_mean=False, dense=False):
    """Reference-call-site convenience: ``GPModelling._neg_log_like`` without the sign.

    ``params``: ``(B, ndim)`` kernel log-parameters (+ one trailing constant-mean value per row
    when ``fit_mean``).  ``mean=None`` -> frozen at ``mean(y)`` (``gpmodelling.py:83-87``).
    """
    p = np.atleast_2d(np.asarray(params))
    y = np.asarray(y, dtype=float)
    diag = (np.asarray(dy, dtype=float) + 1e-12) ** 2
    nk = _terms.n_params(spec)
    if fit_mean in ("linear", "gaussian"):
        # the fitted mean models of GPModelling._build_mean_model (gpmodelling.py:98-122) with the expressions of
        # models/mean_models.py:9-10 (GaussianModel: amplitude / (2 pi sigma) exp(-(t - mean)^2 / (2 sigma^2)) + constant)
        # and :26-27 (LinearModel: slope t + intercept); parameters follow the kernel's in the vector
        tt = np.asarray(t, dtype=float)[None, :]
        if fit_mean == "linear":
            mu = p[:, nk, None] * tt + p[:, nk + 1, None]
        else:
            m0, sg, am, c0 = (p[:, nk + i, None] for i in range(4))
            mu = am / (2 * np.pi * sg) * np.exp(-(tt - m0) ** 2 / (2 * sg ** 2)) + c0
        resid = y[None, :] - mu
    elif fit_mean:
        mu = p[:, nk]
        resid = y[None, :] - mu[:, None]
    else:
        mu0 = np.mean(y) if mean is None else mean
        resid = y - mu0
    co = _terms.coefficients(spec, p[:, :nk])
    if dense:
        out = np.empty(p.shape[0])
        for b in range(p.shape[0]):
            rb = resid[b] if resid.ndim == 2 else resid
            out[b] = loglike_dense(t, np.real(rb), diag, *[np.real(c[b]) for c in co])
        return out
    return loglike_recursion(t, resid, diag, *co)


def grad_complex_step(spec, params, t, y, dy, mean=None, fit_mean=False, h=1e-30):
    """d loglike / d params by complex-step differentiation of the recursion."""
    p = np.atleast_2d(np.asarray(params, dtype=float))
    B, ndim = p.shape
    g = np.empty((B, ndim))
    for k in range(ndim):
        pc = p.astype(np.complex128)
        # (the imaginary part travels separately from the real one, so h = 1e-30 works for parameters of any magnitude:
        #  mean-model parameters span 1e-10 .. 1e9 here)
        pc[:, k] += 1j * h
        ll = loglike_params(spec, pc, t, y, dy, mean=mean, fit_mean=fit_mean)
        g[:, k] = np.imag(ll) / h
    return g


def predict_dense(t, resid, diag, ar, cr, ac, bc, cc, dc, jitter=0.0):
    """celerite ``GP.predict(y, return_var=True)`` at the training times: conditional mean of the
    *noise-free* process (added to the mean model by the caller) and its variance.
    ``mu = K_* K^-1 r``, ``var = k(0) - diag(K_* K^-1 K_*^T)`` with ``K_* = K - diag - jitter``."""
    t = np.asarray(t, dtype=float)
    tau = t[:, None] - t[None, :]
    Ks = _terms.kernel_value(tau, ar, cr, ac, bc, cc, dc)
    K = Ks.copy()
    K[np.diag_indices_from(K)] += np.asarray(diag, dtype=float) + float(jitter)
    cf = scipy.linalg.cho_factor(K, lower=True)
    mu = Ks @ scipy.linalg.cho_solve(cf, resid)
    KinvKs = scipy.linalg.cho_solve(cf, Ks.T)
    var = np.diag(Ks) - np.einsum("ij,ji->i", Ks, KinvKs)
    return mu, var
