"""Oracle: the celerite log-likelihood as a Kalman filter and as its parallel-in-time (associative scan) form.
TEST INFRASTRUCTURE ONLY -- the NumPy statement of what ``csrc/mtg_scan.cu`` (kernels K9 / K10) computes.

* ``statespace``: a celerite kernel ``sum a e^{-c tau} + sum e^{-c tau}(a cos d tau + b sin d tau)`` (Foreman-Mackey et al.
  2017, eq. 8/9) as a linear-Gaussian state-space model, block diagonal over the terms: real term -> scalar decay with
  stationary variance ``a``; complex term -> damped rotation ``e^{-c dt} R(d dt)`` with stationary covariance
  ``[[a, b], [b, a + 2 c b / d]]`` observed through its first component.
* ``loglike_sequential``: prediction-error decomposition (the classical filter).
* ``loglike_scan``: the same filter as a prefix sum over the associative operator of Sarkka & Garcia-Fernandez (IEEE TAC
  2021) with the chunk / scan / re-filter structure of the kernel.
Both equal ``oracle.gp.loglike_dense`` to rounding (tests/test_oracle.py).  Every function is analytic in the coefficients and
works in ``complex128``, so ``grad_complex_step_scan`` differentiates THROUGH the scan: forward-mode derivatives survive the
operator (inverse included) to rounding -- the groundwork for a scan-based gradient kernel."""
import numpy as np


def statespace(ar, cr, ac, bc, cc, dc):
    blocks = [("r", a, c) for a, c in zip(ar, cr) if np.real(a) != 0.0] + [("c", a, b, c, d) for a, b, c, d in zip(ac, bc, cc, dc)]
    J = sum(1 if bl[0] == "r" else 2 for bl in blocks)
    dt_ = np.result_type(*[np.asarray(v).dtype for bl in blocks for v in bl[1:]], np.float64)      # complex for complex-step derivatives
    Pinf = np.zeros((J, J), dtype=dt_); h = np.zeros(J)
    i = 0
    for bl in blocks:
        if bl[0] == "r":
            Pinf[i, i] = bl[1]; h[i] = 1.0; i += 1
        else:
            _, a, b, c, d = bl
            p22 = a if np.real(b) == 0.0 and np.imag(b) == 0.0 else a + 2.0 * c * b / d
            Pinf[i:i + 2, i:i + 2] = [[a, b], [b, p22]]; h[i] = 1.0; i += 2

    def transition(dt):
        F = np.zeros((J, J), dtype=dt_); i = 0
        for bl in blocks:
            if bl[0] == "r":
                F[i, i] = np.exp(-bl[2] * dt); i += 1
            else:
                _, a, b, c, d = bl
                e, cs, sn = np.exp(-c * dt), np.cos(d * dt), np.sin(d * dt)
                F[i:i + 2, i:i + 2] = e * np.array([[cs, sn], [-sn, cs]]); i += 2
        return F
    return J, Pinf, h, transition


def _step(m, P, F, Pinf, h, y, R):
    m = F @ m; P = F @ (P - Pinf) @ F.T + Pinf
    v = y - h @ m; s = h @ P @ h + R
    K = P @ h / s
    return m + K * v, P - np.outer(K, K) * s, -0.5 * (v * v / s + np.log(s) + np.log(2 * np.pi))


def loglike_sequential(t, resid, diag, model):
    J, Pinf, h, transition = model
    m, P, ll = np.zeros(J, dtype=Pinf.dtype), Pinf.copy(), 0.0
    for n in range(len(t)):
        F = transition(t[n] - t[n - 1]) if n else np.eye(J)
        m, P, l = _step(m, P, F, Pinf, h, resid[n], diag[n])
        ll += l
    return ll


def _append(e, F, Pinf, h, y, R):
    """element (A, b, C, eta, Jm) followed by one observation: rank-one form of the operator"""
    A, b, C, eta, Jm = e
    A = F @ A; b = F @ b; P = F @ (C - Pinf) @ F.T + Pinf
    Ph = P @ h; s = h @ Ph + R; w = y - h @ b
    ga = A.T @ h
    eta = eta + ga * w / s; Jm = Jm + np.outer(ga, ga) / s
    K = Ph / s
    return A - np.outer(K, ga), b + K * w, P - np.outer(K, Ph), eta, Jm


def combine(ei, ej):
    Ai, bi, Ci, etai, Ji = ei; Aj, bj, Cj, etaj, Jj = ej
    M = np.linalg.inv(np.eye(len(bi)) + Ci @ Jj)
    AjM = Aj @ M
    return (AjM @ Ai, AjM @ (bi + Ci @ etaj) + bj, AjM @ Ci @ Aj.T + Cj,
            Ai.T @ M.T @ (etaj - Jj @ bi) + etai, Ai.T @ M.T @ Jj @ Ai + Ji)


def loglike_scan(t, resid, diag, model, threads=256):
    J, Pinf, h, transition = model
    N = len(t)
    ch = -(-N // threads)
    bounds = [(min(k * ch, N), min(k * ch + ch, N)) for k in range(threads)]
    cd = Pinf.dtype
    ident = lambda: (np.eye(J, dtype=cd), np.zeros(J, dtype=cd), np.zeros((J, J), dtype=cd), np.zeros(J, dtype=cd), np.zeros((J, J), dtype=cd))
    comps = []
    for k, (n0, n1) in enumerate(bounds):                       # phase 1
        e = (np.zeros((J, J), dtype=cd), np.zeros(J, dtype=cd), Pinf.copy(), np.zeros(J, dtype=cd), np.zeros((J, J), dtype=cd)) if k == 0 else ident()
        for n in range(n0, n1):
            F = transition(t[n] - t[n - 1]) if n else np.eye(J)
            e = _append(e, F, Pinf, h, resid[n], diag[n])
        comps.append(e)
    pref = list(comps); d = 1                                   # phase 2: Kogge-Stone inclusive scan
    while d < threads:
        pref = [combine(pref[i - d], pref[i]) if i >= d else pref[i] for i in range(threads)]
        d *= 2
    ll = 0.0
    for k, (n0, n1) in enumerate(bounds):                       # phase 3
        m, P = (np.zeros(J, dtype=cd), Pinf.copy()) if k == 0 else (pref[k - 1][1], pref[k - 1][2])
        for n in range(n0, n1):
            F = transition(t[n] - t[n - 1]) if n else np.eye(J)
            m, P, l = _step(m, P, F, Pinf, h, resid[n], diag[n])
            ll += l
    return ll


def grad_complex_step_scan(spec, params, t, resid, diag, h=1e-30, threads=64):
    """d loglike / d params by complex-step differentiation through ``loglike_scan`` (one evaluation per parameter)."""
    from . import terms as _terms
    params = np.asarray(params, dtype=float)
    g = np.empty(len(params))
    for k in range(len(params)):
        p = params.astype(complex); p[k] += 1j * h
        co = _terms.coefficients(spec, p)
        ar, cr, ac, bc, cc, dc = [np.atleast_1d(np.asarray(x)) for x in co[:6]]
        jit = co[6] if len(co) > 6 else 0.0
        g[k] = np.imag(loglike_scan(t, resid, diag + jit, statespace(ar, cr, ac, bc, cc, dc), threads=threads)) / h
    return g
