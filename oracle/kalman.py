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


# ---- reverse mode through the filter, chunk by chunk: the NumPy statement of kernel K11 (csrc/mtg_scan_grad.cu) ------------
def _blocks(ar, cr, ac, bc, cc, dc):
    """(kind, first row, coefficient index) per state-space block, in statespace()'s order"""
    out = []; i = 0
    for j in range(len(ar)):
        if np.real(ar[j]) != 0.0:
            out.append(("r", i, j)); i += 1
    for j in range(len(ac)):
        out.append(("c", i, j)); i += 2
    return out, i


def _step_adjoint(lam, Lam, m_prev, P_prev, F, Pinf, h, y, R, first):
    """Adjoint of one filter step.  In: adjoints (lam, Lam) of the FILTERED mean / covariance after the step
    (d loglike = lam . dm + sum_ij Lam_ij dP_ij, Lam symmetric) and the step's inputs.  Out: adjoints of the previous
    filtered state, and of F, P_inf, R and the residual y."""
    mm = F @ m_prev; X = P_prev - Pinf; Pm = F @ X @ F.T + Pinf
    g = Pm @ h; s = h @ g + R; v = y - h @ mm
    lg = lam @ g
    vb = (lg - v) / s
    sb = (-lg * v + g @ Lam @ g) / (s * s) + 0.5 * v * v / (s * s) - 0.5 / s
    gb = lam * (v / s) - 2.0 * (Lam @ g) / s + sb * h
    mb = lam - vb * h
    Pb = Lam + 0.5 * (np.outer(gb, h) + np.outer(h, gb))
    lam_prev = F.T @ mb
    Lam_prev = F.T @ Pb @ F
    Pinf_b = Pb - Lam_prev
    F_b = np.zeros_like(F) if first else 2.0 * Pb @ F @ X + np.outer(mb, m_prev)
    return lam_prev, Lam_prev, F_b, Pinf_b, sb, vb


def grad_chunked_adjoint(t, resid, diag, ar, cr, ac, bc, cc, dc, jitter=0.0, threads=8):
    """log-likelihood and its gradient with respect to the celerite coefficients, the jitter and the residuals, with the
    structure of kernel K11: (1) forward filter (the kernel gets every chunk's start state from the scan; here it is simply
    run in order), recording per chunk the closed-loop transition ``Phi = prod (I - K h^T) F`` and the backward information
    ``q = sum w v / s``, ``Omega = sum w w^T / s`` (``w = Psi^T F^T h``); (2) a SUFFIX scan over the chunks with the
    associative operator ``(a then b) -> (Phi_b Phi_a, Phi_a^T q_b + q_a, Phi_a^T Omega_b Phi_a + Omega_a)`` gives the
    adjoint of every chunk's end state in closed form, ``lam = q_suf``, ``Lam = (q_suf q_suf^T - Omega_suf) / 2``; (3) every
    chunk runs the step adjoints over its own observations independently.  Returns ``(ll, dict of gradients)``."""
    ar, cr, ac, bc, cc, dc = [np.atleast_1d(np.asarray(x, dtype=float)) for x in (ar, cr, ac, bc, cc, dc)]
    J, Pinf, h, transition = statespace(ar, cr, ac, bc, cc, dc)
    blocks, _ = _blocks(ar, cr, ac, bc, cc, dc)
    N = len(t)
    ch = -(-N // threads)
    bounds = [(min(k * ch, N), min(k * ch + ch, N)) for k in range(threads)]
    # (1) forward
    m, P, ll = np.zeros(J), Pinf.copy(), 0.0
    tape = [None] * N
    elems = []
    for (n0, n1) in bounds:
        Psi = np.eye(J); q = np.zeros(J); Om = np.zeros((J, J))
        for n in range(n0, n1):
            F = transition(t[n] - t[n - 1]) if n else np.eye(J)
            R = diag[n] + jitter
            tape[n] = (m, P, F, R)
            mm = F @ m; Pm = F @ (P - Pinf) @ F.T + Pinf
            g = Pm @ h; s = h @ g + R; v = resid[n] - h @ mm
            w = (F @ Psi).T @ h
            q = q + w * (v / s); Om = Om + np.outer(w, w) / s
            Psi = (np.eye(J) - np.outer(g, h) / s) @ F @ Psi
            m = mm + g * (v / s); P = Pm - np.outer(g, g) / s
            ll += -0.5 * (v * v / s + np.log(s) + np.log(2 * np.pi))
        elems.append((Psi, q, Om))
    # (2) exclusive suffix composites
    suf = [None] * threads
    acc = (np.eye(J), np.zeros(J), np.zeros((J, J)))
    for k in range(threads - 1, -1, -1):
        suf[k] = acc
        Pa, qa, Oa = elems[k]; Pb, qb, Ob = acc
        acc = (Pb @ Pa, Pa.T @ qb + qa, Pa.T @ Ob @ Pa + Oa)
    # (3) per-chunk adjoint sweeps
    g_ar = np.zeros(len(ar)); g_cr = np.zeros(len(cr)); g_ac = np.zeros(len(ac)); g_bc = np.zeros(len(ac))
    g_cc = np.zeros(len(ac)); g_dc = np.zeros(len(ac)); g_jit = 0.0; g_res = np.zeros(N)
    for k, (n0, n1) in enumerate(bounds):
        _, qs, Os = suf[k]
        lam = qs.copy(); Lam = 0.5 * (np.outer(qs, qs) - Os)
        for n in range(n1 - 1, n0 - 1, -1):
            m_prev, P_prev, F, R = tape[n]
            lam, Lam, F_b, Pinf_b, sb, vb = _step_adjoint(lam, Lam, m_prev, P_prev, F, Pinf, h, resid[n], R, n == 0)
            g_jit += sb; g_res[n] = vb
            dt = t[n] - t[n - 1] if n else 0.0
            for kind, i, j in blocks:
                if kind == "r":
                    g_ar[j] += Pinf_b[i, i]; g_cr[j] += -dt * F[i, i] * F_b[i, i]
                else:
                    a, b, c, d = ac[j], bc[j], cc[j], dc[j]
                    Fb = F_b[i:i + 2, i:i + 2]; Fk = F[i:i + 2, i:i + 2]; Pb = Pinf_b[i:i + 2, i:i + 2]
                    g_cc[j] += -dt * np.sum(Fb * Fk)
                    # dF/dd = dt e [[-sn, cs], [-cs, -sn]] = dt [[F01, F00], [-F00, F01]] with F00 = e cs, F01 = e sn
                    g_dc[j] += dt * (-Fk[0, 1] * Fb[0, 0] + Fk[0, 0] * Fb[0, 1] - Fk[0, 0] * Fb[1, 0] - Fk[0, 1] * Fb[1, 1])
                    g_ac[j] += Pb[0, 0] + Pb[1, 1]
                    g_bc[j] += Pb[0, 1] + Pb[1, 0]
                    if b != 0.0:
                        g_bc[j] += Pb[1, 1] * 2.0 * c / d; g_cc[j] += Pb[1, 1] * 2.0 * b / d; g_dc[j] += -Pb[1, 1] * 2.0 * c * b / (d * d)
        if k == 0:                                   # the prior covariance IS P_inf
            for kind, i, j in blocks:
                if kind == "r":
                    g_ar[j] += Lam[i, i]
                else:
                    a, b, c, d = ac[j], bc[j], cc[j], dc[j]
                    Pb = Lam[i:i + 2, i:i + 2]
                    g_ac[j] += Pb[0, 0] + Pb[1, 1]; g_bc[j] += Pb[0, 1] + Pb[1, 0]
                    if b != 0.0:
                        g_bc[j] += Pb[1, 1] * 2.0 * c / d; g_cc[j] += Pb[1, 1] * 2.0 * b / d; g_dc[j] += -Pb[1, 1] * 2.0 * c * b / (d * d)
    return ll, {"ar": g_ar, "cr": g_cr, "ac": g_ac, "bc": g_bc, "cc": g_cc, "dc": g_dc, "jitter": g_jit, "resid": g_res}
