"""The ``celerite.terms`` protocol used by mind_the_gaps, re-implemented without celerite.

Reference usage: ``celerite.terms.Term`` subclasses in
``/root/reference/mind_the_gaps/models/celerite_models.py:7-90`` (``parameter_names``,
``get_real_coefficients``, ``get_complex_coefficients``, ``log_prior``, ``__add__``, ``get_psd``),
``celerite.terms.RealTerm`` / ``ComplexTerm`` in ``docs/notebooks/tutorial_ppp.ipynb`` (cells 7, 9),
``SHOTerm`` / ``Matern32Term`` in ``tests/models_test.py:31-62``, ``JitterTerm`` in
``docs/notebooks/poisson_level.ipynb``.

Every concrete term also carries ``_mtg_kind``: the ``mtg_term_kind`` of ``include/mtg_b200.h`` that
tells the device kernels how to map its log-parameters to celerite coefficients (the map itself is
restated on the device in ``csrc/mtg_common.cuh: plan_coefficients``).  The NumPy maps below are the
host-side protocol (``get_all_coefficients``, ``get_psd``, ``get_value``), not the hot path.
"""
import numpy as np

from .modeling import Model, ModelSet

__all__ = ["Term", "TermSum", "RealTerm", "ComplexTerm", "SHOTerm", "Matern32Term", "JitterTerm"]

# mtg_term_kind (include/mtg_b200.h)
KIND_REAL, KIND_COMPLEX3, KIND_COMPLEX4, KIND_SHO, KIND_MATERN32, KIND_JITTER = 0, 1, 2, 3, 4, 5
KIND_DRW, KIND_LORENTZIAN, KIND_COSINUS, KIND_BENDINGPL = 6, 7, 8, 9


class Term(Model):
    """A celerite kernel term: a sum of real ``a e^{-c tau}`` and complex
    ``e^{-c tau}(a cos d tau + b sin d tau)`` components."""

    _mtg_kind = None      # set by concrete terms the device kernels know
    _mtg_extra = 0.0

    @property
    def terms(self):
        return [self]

    # -- coefficient protocol -----------------------------------------------------------------
    def get_real_coefficients(self, params):
        return np.empty(0), np.empty(0)

    def get_complex_coefficients(self, params):
        return np.empty(0), np.empty(0), np.empty(0), np.empty(0)

    def get_jitter(self, params):
        return 0.0

    @property
    def jitter(self):
        return self.get_jitter(self.get_parameter_vector(include_frozen=True))

    def get_all_coefficients(self, vector=None):
        if vector is None:
            vector = self.get_parameter_vector(include_frozen=True)
        r = self.get_real_coefficients(vector)
        c = self.get_complex_coefficients(vector)
        if len(c) == 3:
            c = (c[0], np.zeros_like(np.atleast_1d(c[0]), dtype=float), c[1], c[2])
        return tuple(np.atleast_1d(np.asarray(x, dtype=np.float64)) for x in tuple(r) + tuple(c))

    @property
    def coefficients(self):
        return self.get_all_coefficients()

    # -- evaluation ----------------------------------------------------------------------------
    def get_value(self, tau):
        ar, cr, ac, bc, cc, dc = self.coefficients
        tau = np.abs(np.asarray(tau, dtype=np.float64))
        k = np.zeros_like(tau)
        for a, c in zip(ar, cr):
            k = k + a * np.exp(-c * tau)
        for a, b, c, d in zip(ac, bc, cc, dc):
            k = k + np.exp(-c * tau) * (a * np.cos(d * tau) + b * np.sin(d * tau))
        return k

    def get_psd(self, omega):
        """Power spectral density in celerite's normalisation (Foreman-Mackey et al. 2017, eq. 9)."""
        ar, cr, ac, bc, cc, dc = self.coefficients
        w = np.asarray(omega, dtype=np.float64)
        w2 = w * w
        p = np.zeros_like(w)
        with np.errstate(invalid="ignore", divide="ignore"):
            for a, c in zip(ar, cr):
                p = p + a * c / (c * c + w2)
            for a, b, c, d in zip(ac, bc, cc, dc):
                w02 = c * c + d * d
                p = p + ((a * c + b * d) * w02 + (a * c - b * d) * w2) / (w2 * w2 + 2.0 * (c * c - d * d) * w2 + w02 * w02)
        return np.sqrt(2.0 / np.pi) * p

    # -- algebra ---------------------------------------------------------------------------------
    def __add__(self, other):
        if not isinstance(other, Term):
            return NotImplemented
        return TermSum(*(self.terms + other.terms))

    def __radd__(self, other):
        if not isinstance(other, Term):
            return NotImplemented
        return TermSum(*(other.terms + self.terms))

    def __mul__(self, other):
        raise NotImplementedError("kernel products are outside the mind_the_gaps hot path")


class TermSum(Term, ModelSet):
    """``k1 + k2``: coefficients concatenated in term order, parameters named ``terms[i]:name``."""

    def __init__(self, *terms):
        if any(isinstance(t, TermSum) for t in terms):
            raise ValueError("TermSum members must be plain terms")
        ModelSet.__init__(self, [("terms[{0}]".format(i), t) for i, t in enumerate(terms)])

    def __getattr__(self, name):
        return ModelSet.__getattr__(self, name)

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)

    @property
    def terms(self):
        return list(self.models.values())

    def __repr__(self):
        return " + ".join(repr(t) for t in self.terms)

    @property
    def jitter(self):
        return float(sum(t.jitter for t in self.terms))

    def get_all_coefficients(self, vector=None):
        if vector is not None:
            self.set_parameter_vector(vector, include_frozen=True)
        parts = [t.get_all_coefficients() for t in self.terms]
        return tuple(np.concatenate([p[i] for p in parts]) if parts else np.empty(0) for i in range(6))

    def log_prior(self):
        return ModelSet.log_prior(self)


class RealTerm(Term):
    r"""``k(tau) = a e^{-c tau}`` with ``log_a, log_c``."""
    parameter_names = ("log_a", "log_c")
    _mtg_kind = KIND_REAL

    def __repr__(self):
        return "RealTerm({0.log_a}, {0.log_c})".format(self)

    def get_real_coefficients(self, params):
        log_a, log_c = params
        return np.exp(log_a), np.exp(log_c)


class ComplexTerm(Term):
    r"""``k(tau) = e^{-c tau}(a cos d tau + b sin d tau)``; ``log_b`` is optional (``b = 0``)."""

    def __init__(self, *args, **kwargs):
        if len(args) == 4 or "log_b" in kwargs:
            self.fit_b = True
            self.parameter_names = ("log_a", "log_b", "log_c", "log_d")
        else:
            self.fit_b = False
            self.parameter_names = ("log_a", "log_c", "log_d")
        super(ComplexTerm, self).__init__(*args, **kwargs)

    # parameter_names is per instance here; route name lookups through it
    def __getattr__(self, name):
        d = self.__dict__
        if name in d.get("parameter_names", ()) and "parameter_vector" in d:
            return d["parameter_vector"][d["parameter_names"].index(name)]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        d = self.__dict__
        if name in d.get("parameter_names", ()) and "parameter_vector" in d and name != "parameter_names":
            d["parameter_vector"][d["parameter_names"].index(name)] = value
            self.dirty = True
        else:
            object.__setattr__(self, name, value)

    @property
    def _mtg_kind(self):
        return KIND_COMPLEX4 if self.fit_b else KIND_COMPLEX3

    def __repr__(self):
        return "ComplexTerm({0})".format(", ".join(repr(float(v)) for v in self.parameter_vector))

    def get_complex_coefficients(self, params):
        if not self.fit_b:
            log_a, log_c, log_d = params
            return np.exp(log_a), 0.0, np.exp(log_c), np.exp(log_d)
        log_a, log_b, log_c, log_d = params
        return np.exp(log_a), np.exp(log_b), np.exp(log_c), np.exp(log_d)

    def log_prior(self):
        # celerite: a c >= b d is required for a positive definite term
        if self.fit_b:
            a, b, c, d = self.get_complex_coefficients(self.parameter_vector)
            if a * c < b * d:
                return -np.inf
        return super(ComplexTerm, self).log_prior()


class SHOTerm(Term):
    r"""Stochastically driven damped harmonic oscillator, ``S(w) = sqrt(2/pi) S0 w0^4 /
    ((w^2-w0^2)^2 + w0^2 w^2/Q^2)`` (F-M+17 eq. 20); one complex term for ``Q >= 1/2``, two real
    terms below."""
    parameter_names = ("log_S0", "log_Q", "log_omega0")
    _mtg_kind = KIND_SHO

    def __repr__(self):
        return "SHOTerm({0.log_S0}, {0.log_Q}, {0.log_omega0})".format(self)

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
        f = np.sqrt(4.0 * Q ** 2 - 1.0)
        return S0 * w0 * Q, S0 * w0 * Q / f, 0.5 * w0 / Q, 0.5 * w0 / Q * f


class Matern32Term(Term):
    r"""celerite's approximate Matern-3/2: ``sigma^2 (1 + sqrt3 tau/rho) e^{-sqrt3 tau/rho}`` as
    ``eps -> 0``."""
    parameter_names = ("log_sigma", "log_rho")
    _mtg_kind = KIND_MATERN32

    def __init__(self, *args, **kwargs):
        eps = kwargs.pop("eps", 0.01)
        super(Matern32Term, self).__init__(*args, **kwargs)
        self.eps = float(eps)

    @property
    def _mtg_extra(self):
        return self.eps

    def __repr__(self):
        return "Matern32Term({0.log_sigma}, {0.log_rho}, eps={0.eps})".format(self)

    def get_complex_coefficients(self, params):
        log_sigma, log_rho = params
        w0 = np.sqrt(3.0) * np.exp(-log_rho)
        S0 = np.exp(2.0 * log_sigma) / w0
        return w0 * S0, w0 * w0 * S0 / self.eps, w0, self.eps


class JitterTerm(Term):
    r"""White noise ``sigma^2 delta_ij`` added to the diagonal."""
    parameter_names = ("log_sigma",)
    _mtg_kind = KIND_JITTER

    def __repr__(self):
        return "JitterTerm({0.log_sigma})".format(self)

    def get_jitter(self, params):
        return float(np.exp(2.0 * params[0]))
