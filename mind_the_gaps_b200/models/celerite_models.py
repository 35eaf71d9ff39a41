"""mind_the_gaps kernel terms (``/root/reference/mind_the_gaps/models/celerite_models.py``).

Each class keeps the reference's ``parameter_names`` and coefficient map; ``_mtg_kind`` selects the
same map inside the CUDA kernels (``csrc/mtg_common.cuh: plan_coefficients``).
"""
import numpy as np

from ..celerite import terms as _terms
from ..celerite.terms import Term

__all__ = ["Lorentzian", "Cosinus", "DampedRandomWalk", "BendingPowerlaw"]


class Lorentzian(Term):
    """QPO: complex term ``(S0, 0, w0/2Q, w0)`` plus the reference's degenerate real term ``(0, 0)``
    (celerite_models.py:7-34), which makes this a width-3 term in celerite's count."""
    parameter_names = ("log_S0", "log_Q", "log_omega0")
    _mtg_kind = _terms.KIND_LORENTZIAN

    def get_real_coefficients(self, params):
        return 0, 0

    def get_complex_coefficients(self, params):
        log_S0, log_Q, log_omega0 = params
        w0 = np.exp(log_omega0)
        return np.exp(log_S0), 0, 0.5 * w0 / np.exp(log_Q), w0

    def __repr__(self):
        return "Lorentzian({0.log_S0}, {0.log_Q}, {0.log_omega0})".format(self)


class Cosinus(Term):
    """Undamped oscillation ``S0 cos(w0 tau)`` (celerite_models.py:36-52)."""
    parameter_names = ("log_S0", "log_omega0")
    _mtg_kind = _terms.KIND_COSINUS

    def get_complex_coefficients(self, params):
        log_S0, log_omega0 = params
        return np.exp(log_S0), 0, 0, np.exp(log_omega0)

    def __repr__(self):
        return "Cosinus({0.log_S0}, {0.log_omega0})".format(self)


class DampedRandomWalk(Term):
    """Eq. 13 of Foreman-Mackey+2017 with Q = 1/2: real term ``(S0, w0)`` (celerite_models.py:55-68)."""
    parameter_names = ("log_S0", "log_omega0")
    _mtg_kind = _terms.KIND_DRW

    def get_real_coefficients(self, params):
        log_S0, log_omega0 = params
        return np.exp(log_S0), np.exp(log_omega0)

    def __repr__(self):
        return "DampedRandomWalk({0.log_S0}, {0.log_omega0})".format(self)


class BendingPowerlaw(Term):
    """Complex term ``(S0, Q, w0, w0)``; w^-2 for log_Q << log_S0, w^-4 for log_Q ~ log_S0
    (celerite_models.py:71-90)."""
    parameter_names = ("log_S0", "log_Q", "log_omega0")
    _mtg_kind = _terms.KIND_BENDINGPL

    def get_complex_coefficients(self, params):
        log_S0, log_Q, log_omega0 = params
        w0 = np.exp(log_omega0)
        return np.exp(log_S0), np.exp(log_Q), w0, w0

    def log_prior(self):
        # positive-definiteness constraint of the reference (celerite_models.py:85-90)
        if self.log_S0 < self.log_Q:
            return -np.inf
        return super(BendingPowerlaw, self).log_prior()

    def __repr__(self):
        return "BendingPowerlaw({0.log_S0}, {0.log_Q}, {0.log_omega0})".format(self)
