"""Oracle: kernel-term parameter -> celerite coefficient maps and PSDs.  TEST INFRASTRUCTURE ONLY.

A kernel is described by a plain list of ``(kind, n_params[, extra])`` entries consumed
left to right from a flat log-parameter vector, exactly the order celerite's ``TermSum``
concatenates parameters and coefficients in (``k1 + k2`` -> terms[0], terms[1], ...).

Follows
-------
* ``/root/reference/mind_the_gaps/models/celerite_models.py:7-34``  Lorentzian (one complex term AND
  one degenerate real term ``(0, 0)``)
* ``celerite_models.py:36-52``  Cosinus, ``:55-68`` DampedRandomWalk, ``:71-90`` BendingPowerlaw
* celerite 0.4.2 ``terms.py`` (un-vendored; published in Foreman-Mackey et al. 2017, eqs. 8-9, 23-30):
  RealTerm, ComplexTerm (3- and 4-parameter forms), SHOTerm (both Q regimes), Matern32Term, JitterTerm
* PSD: celerite ``get_psd`` (F-M+17 eq. 9):
  ``S(w) = sqrt(2/pi) [ sum a c/(c^2+w^2) + sum ((ac+bd)(c^2+d^2) + (ac-bd) w^2) / (w^4 + 2(c^2-d^2)w^2 + (c^2+d^2)^2) ]``
"""
import numpy as np

# kind -> number of parameters
NPARAMS = {
    "real": 2,          # log_a, log_c
    "complex3": 3,      # log_a, log_c, log_d      (b = 0)
    "complex4": 4,      # log_a, log_b, log_c, log_d
    "sho": 3,           # log_S0, log_Q, log_omega0
    "matern32": 2,      # log_sigma, log_rho       (extra: eps, default 0.01)
    "jitter": 1,        # log_sigma
    "drw": 2,           # log_S0, log_omega0
    "lorentzian": 3,    # log_S0, log_Q, log_omega0
    "cosinus": 2,       # log_S0, log_omega0
    "bendingpl": 3,     # log_S0, log_Q, log_omega0
}


def n_params(spec):
    return sum(NPARAMS[e[0]] for e in spec)


def coefficients(spec, params):
    """Map a flat parameter vector (or a batch ``(B, ndim)``) to celerite coefficients.

    Returns ``(ar, cr, ac, bc, cc, dc, jitter)`` with shapes ``(B, Jr)``, ``(B, Jc)`` and ``(B,)``
    (leading axis dropped for a 1-D input).  dtype follows ``params`` so the map can be
    evaluated in ``complex128`` for complex-step differentiation.
    """
    p = np.asarray(params)
    squeeze = p.ndim == 1
    p = np.atleast_2d(p)
    B = p.shape[0]
    dt = p.dtype if np.iscomplexobj(p) else np.float64
    ar, cr, ac, bc, cc, dc = [], [], [], [], [], []
    jitter = np.zeros(B, dtype=dt)
    zero = np.zeros(B, dtype=dt)
    k = 0
    for entry in spec:
        kind = entry[0]
        q = [p[:, k + i] for i in range(NPARAMS[kind])]
        k += NPARAMS[kind]
        if kind == "real":
            ar.append(np.exp(q[0])); cr.append(np.exp(q[1]))
        elif kind == "complex3":
            ac.append(np.exp(q[0])); bc.append(zero); cc.append(np.exp(q[1])); dc.append(np.exp(q[2]))
        elif kind == "complex4":
            ac.append(np.exp(q[0])); bc.append(np.exp(q[1])); cc.append(np.exp(q[2])); dc.append(np.exp(q[3]))
        elif kind == "drw":
            # celerite_models.py:57-65 : Q = 1/2 -> c = 0.5*w0/Q = w0
            S0 = np.exp(q[0]); w0 = np.exp(q[1])
            ar.append(S0); cr.append(0.5 * w0 / 0.5)
        elif kind == "lorentzian":
            # celerite_models.py:9-31 : degenerate real (0, 0) + complex (S0, 0, w0/2Q, w0)
            S0 = np.exp(q[0]); Q = np.exp(q[1]); w0 = np.exp(q[2])
            ar.append(zero); cr.append(zero)
            ac.append(S0); bc.append(zero); cc.append(0.5 * w0 / Q); dc.append(w0)
        elif kind == "cosinus":
            S0 = np.exp(q[0]); w0 = np.exp(q[1])
            ac.append(S0); bc.append(zero); cc.append(zero); dc.append(w0)
        elif kind == "bendingpl":
            w0 = np.exp(q[2])
            ac.append(np.exp(q[0])); bc.append(np.exp(q[1])); cc.append(w0); dc.append(w0)
        elif kind == "sho":
            S0 = np.exp(q[0]); Q = np.exp(q[1]); w0 = np.exp(q[2])
            Qr = np.real(Q)
            if np.all(Qr >= 0.5):
                f = np.sqrt(4.0 * Q * Q - 1.0)
                ac.append(S0 * w0 * Q); bc.append(S0 * w0 * Q / f)
                cc.append(0.5 * w0 / Q); dc.append(0.5 * w0 / Q * f)
            elif np.all(Qr < 0.5):
                f = np.sqrt(1.0 - 4.0 * Q * Q)
                ar.append(0.5 * S0 * w0 * Q * (1.0 + 1.0 / f)); cr.append(0.5 * w0 / Q * (1.0 - f))
                ar.append(0.5 * S0 * w0 * Q * (1.0 - 1.0 / f)); cr.append(0.5 * w0 / Q * (1.0 + f))
            else:
                raise ValueError("oracle: a batch must not straddle Q = 1/2 for an SHO term")
        elif kind == "matern32":
            eps = entry[1] if len(entry) > 1 else 0.01
            w0 = np.sqrt(3.0) * np.exp(-q[1])
            S0 = np.exp(2.0 * q[0]) / w0
            ac.append(w0 * S0); bc.append(w0 * w0 * S0 / eps); cc.append(w0); dc.append(zero + eps)
        elif kind == "jitter":
            jitter = jitter + np.exp(2.0 * q[0])
        else:
            raise ValueError("unknown term kind %r" % (kind,))
    if k != p.shape[1]:
        raise ValueError("parameter vector has %d entries, kernel spec consumes %d" % (p.shape[1], k))

    def stack(lst):
        return np.stack(lst, axis=1) if lst else np.zeros((B, 0), dtype=dt)

    out = [stack(x) for x in (ar, cr, ac, bc, cc, dc)] + [jitter]
    if squeeze:
        out = [o[0] for o in out]
    return tuple(out)


def psd_from_coefficients(omega, ar, cr, ac, bc, cc, dc):
    """celerite ``get_psd`` (F-M+17 eq. 9) for one coefficient set; ``omega`` angular frequency."""
    w = np.asarray(omega, dtype=float)
    w2 = w * w
    p = np.zeros_like(w)
    with np.errstate(invalid="ignore", divide="ignore"):
        for a, c in zip(np.atleast_1d(ar), np.atleast_1d(cr)):
            p = p + a * c / (c * c + w2)
        for a, b, c, d in zip(np.atleast_1d(ac), np.atleast_1d(bc), np.atleast_1d(cc), np.atleast_1d(dc)):
            w02 = c * c + d * d
            p = p + ((a * c + b * d) * w02 + (a * c - b * d) * w2) / (w2 * w2 + 2.0 * (c * c - d * d) * w2 + w02 * w02)
    return np.sqrt(2.0 / np.pi) * p


def psd(spec, params, omega):
    ar, cr, ac, bc, cc, dc, _ = coefficients(spec, np.asarray(params, dtype=float))
    return psd_from_coefficients(omega, ar, cr, ac, bc, cc, dc)


def kernel_value(tau, ar, cr, ac, bc, cc, dc):
    """k(tau) (F-M+17 eq. 8) for one coefficient set (jitter NOT included)."""
    tau = np.abs(np.asarray(tau, dtype=float))
    k = np.zeros_like(tau)
    for a, c in zip(np.atleast_1d(ar), np.atleast_1d(cr)):
        k = k + a * np.exp(-c * tau)
    for a, b, c, d in zip(np.atleast_1d(ac), np.atleast_1d(bc), np.atleast_1d(cc), np.atleast_1d(dc)):
        k = k + np.exp(-c * tau) * (a * np.cos(d * tau) + b * np.sin(d * tau))
    return k


# ---- analytic PSD models of mind_the_gaps/models/psd_models.py (formulas only) -------------------

def psd_sho(w, S0=1.0, Q=10.0, omega0=1.0):
    """psd_models.py:7-11"""
    return np.sqrt(2 / np.pi) * S0 * omega0 ** 4 / ((w ** 2 - omega0 ** 2) ** 2 + (w ** 2) * (omega0 ** 2) / Q ** 2)


def psd_lorentzian(w, S0=1.0, Q=10.0, omega0=1.0):
    """psd_models.py:14-32"""
    a = S0
    c = omega0 / 2 / Q
    return np.sqrt(1 / 2 / np.pi) * a / c * (1 / (1 + ((w - omega0) / c) ** 2) + 1 / (1 + ((w + omega0) / c) ** 2))


def psd_bendingpowerlaw(w, S0=1.0, omega0=1.0, Q=0.5):
    """psd_models.py:35-46"""
    a = S0
    c = 0.5 * omega0 / Q
    return np.sqrt(2 / np.pi) * a / c * (1 / (1 + (w / c) ** 2))


def psd_matern32(w, sigma=1.0, rho=1.0):
    """psd_models.py:63-67"""
    return 1 / np.sqrt(2 * np.pi) * sigma ** 2 * 4 / np.sqrt(3) * rho * (1 / (1 + (w * rho / np.sqrt(3)) ** 2)) ** 2


def psd_matern52(w, sigma=1.0, rho=1.0):
    """psd_models.py:69-73"""
    return 1 / np.sqrt(2 * np.pi) * sigma ** 2 * 16 / 3 / np.sqrt(5) * rho * (1 / (1 + (w * rho / np.sqrt(5)) ** 2)) ** 3
