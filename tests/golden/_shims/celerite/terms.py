"""celerite.terms protocol (restated): coefficient plumbing, sums, PSD."""
from itertools import chain
import numpy as np
from .modeling import Model, ModelSet


class Term(Model):
    @property
    def terms(self):
        return [self]

    def get_real_coefficients(self, params):
        return np.empty(0), np.empty(0)

    def get_complex_coefficients(self, params):
        return np.empty(0), np.empty(0), np.empty(0), np.empty(0)

    def get_all_coefficients(self, vector=None):
        if vector is None:
            vector = self.get_parameter_vector(include_frozen=True)
        pars = self.get_real_coefficients(vector)
        if len(pars) != 2:
            raise ValueError("there must be 2 real coefficient blocks")
        ar, cr = (np.atleast_1d(np.asarray(p, dtype=float)) for p in pars)
        pars = self.get_complex_coefficients(vector)
        if len(pars) == 3:
            a, c, d = (np.atleast_1d(np.asarray(p, dtype=float)) for p in pars)
            b = np.zeros_like(a)
        elif len(pars) == 4:
            a, b, c, d = (np.atleast_1d(np.asarray(p, dtype=float)) for p in pars)
        else:
            raise ValueError("there must be 3 or 4 complex coefficient blocks")
        return ar, cr, a, b, c, d

    coefficients = property(lambda self: self.get_all_coefficients())

    @property
    def jitter(self):
        return self.get_jitter(self.get_parameter_vector(include_frozen=True))

    def get_jitter(self, params):
        return 0.0

    def get_value(self, tau):
        tau = np.abs(np.asarray(tau, dtype=float))
        ar, cr, a, b, c, d = self.coefficients
        k = np.zeros_like(tau)
        for aj, cj in zip(ar, cr):
            k = k + aj * np.exp(-cj * tau)
        for aj, bj, cj, dj in zip(a, b, c, d):
            k = k + np.exp(-cj * tau) * (aj * np.cos(dj * tau) + bj * np.sin(dj * tau))
        return k

    def get_psd(self, omega):
        w = np.asarray(omega, dtype=float)
        ar, cr, a, b, c, d = self.coefficients
        w2 = w * w
        p = np.zeros_like(w)
        for aj, cj in zip(ar, cr):
            p = p + aj * cj / (cj * cj + w2)
        for aj, bj, cj, dj in zip(a, b, c, d):
            w02 = cj * cj + dj * dj
            p = p + ((aj * cj + bj * dj) * w02 + (aj * cj - bj * dj) * w2) / (w2 * w2 + 2.0 * (cj * cj - dj * dj) * w2 + w02 * w02)
        return np.sqrt(2.0 / np.pi) * p

    def __add__(self, b):
        return TermSum(self, b)

    def __radd__(self, b):
        return TermSum(b, self)


class TermSum(Term, ModelSet):
    def __init__(self, *terms):
        flat = []
        for t in terms:
            flat += list(t.terms)
        self._terms = flat
        ModelSet.__init__(self, [("terms[{0}]".format(i), t) for i, t in enumerate(flat)])

    @property
    def terms(self):
        return self._terms

    def get_all_coefficients(self, vector=None):
        cs = [t.get_all_coefficients() for t in self._terms]
        return tuple(np.concatenate([c[i] for c in cs]) for i in range(6))

    coefficients = property(lambda self: self.get_all_coefficients())

    @property
    def jitter(self):
        return sum(t.jitter for t in self._terms)


class RealTerm(Term):
    parameter_names = ("log_a", "log_c")

    def get_real_coefficients(self, params):
        log_a, log_c = params
        return np.exp(log_a), np.exp(log_c)


class ComplexTerm(Term):
    def __init__(self, *args, **kwargs):
        if len(args) == 3 and "log_b" not in kwargs:
            self.fit_b = False
            self.parameter_names = ("log_a", "log_c", "log_d")
        else:
            self.fit_b = "log_b" in kwargs or len(args) == 4
            self.parameter_names = ("log_a", "log_b", "log_c", "log_d") if self.fit_b else ("log_a", "log_c", "log_d")
        super(ComplexTerm, self).__init__(*args, **kwargs)

    def get_complex_coefficients(self, params):
        if self.fit_b:
            log_a, log_b, log_c, log_d = params
            return np.exp(log_a), np.exp(log_b), np.exp(log_c), np.exp(log_d)
        log_a, log_c, log_d = params
        return np.exp(log_a), 0.0, np.exp(log_c), np.exp(log_d)


class SHOTerm(Term):
    parameter_names = ("log_S0", "log_Q", "log_omega0")

    def get_real_coefficients(self, params):
        log_S0, log_Q, log_omega0 = params
        Q = np.exp(log_Q)
        if Q >= 0.5:
            return np.empty(0), np.empty(0)
        S0 = np.exp(log_S0); w0 = np.exp(log_omega0)
        f = np.sqrt(1.0 - 4.0 * Q ** 2)
        return (0.5 * S0 * w0 * Q * np.array([1.0 + 1.0 / f, 1.0 - 1.0 / f]),
                0.5 * w0 / Q * np.array([1.0 - f, 1.0 + f]))

    def get_complex_coefficients(self, params):
        log_S0, log_Q, log_omega0 = params
        Q = np.exp(log_Q)
        if Q < 0.5:
            return np.empty(0), np.empty(0), np.empty(0), np.empty(0)
        S0 = np.exp(log_S0); w0 = np.exp(log_omega0)
        f = np.sqrt(4.0 * Q ** 2 - 1)
        return S0 * w0 * Q, S0 * w0 * Q / f, 0.5 * w0 / Q, 0.5 * w0 / Q * f


class Matern32Term(Term):
    parameter_names = ("log_sigma", "log_rho")

    def __init__(self, *args, **kwargs):
        self.eps = kwargs.pop("eps", 0.01)
        super(Matern32Term, self).__init__(*args, **kwargs)

    def get_complex_coefficients(self, params):
        log_sigma, log_rho = params
        w0 = np.sqrt(3.0) * np.exp(-log_rho)
        S0 = np.exp(2.0 * log_sigma) / w0
        return w0 * S0, w0 * w0 * S0 / self.eps, w0, self.eps


class JitterTerm(Term):
    parameter_names = ("log_sigma", )

    def get_jitter(self, params):
        return np.exp(2.0 * params[0])
