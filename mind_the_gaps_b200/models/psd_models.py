"""Analytic PSDs in celerite's normalisation (``/root/reference/mind_the_gaps/models/psd_models.py``).

The reference wraps these formulas in ``astropy.modeling.custom_model``; astropy is only a carrier
there (parameters as attributes, call with angular frequencies, ``+`` to add models), so a minimal
carrier class is used instead.  Each model also exposes ``device_spec()`` so the simulator's
spectrum-synthesis kernel can evaluate it on the GPU without calling back into Python.
"""
from math import pi, sqrt

import numpy as np
from scipy.special import gamma

__all__ = ["SHO", "Lorentzian", "BendingPowerlaw", "Matern", "Matern32", "Matern52", "Jitter", "PSDModel", "PSDSum"]

# device PSD component kinds (csrc/mtg_sim.cu)
PSD_SHO, PSD_LORENTZIAN, PSD_BENDINGPL, PSD_MATERN32, PSD_MATERN52, PSD_POWERLAW = 0, 1, 2, 3, 4, 5


class PSDModel(object):
    """Callable PSD with named, settable parameters (the part of astropy's Model the reference uses:
    construction by position or keyword, attribute get/set -- ``Simulator.set_psd_params``,
    simulator.py:282-298 -- and evaluation on an array of angular frequencies)."""
    param_names = tuple()
    defaults = tuple()
    _device_kind = None

    def __init__(self, *args, **kwargs):
        values = list(self.defaults)
        if len(args) > len(values):
            raise TypeError("too many parameters")
        values[:len(args)] = args
        for k, v in kwargs.items():
            if k not in self.param_names:
                raise TypeError("unknown parameter %r" % k)
            values[self.param_names.index(k)] = v
        for k, v in zip(self.param_names, values):
            setattr(self, k, float(v))

    @property
    def parameters(self):
        return np.array([getattr(self, k) for k in self.param_names])

    def evaluate(self, x, *params):
        raise NotImplementedError

    def __call__(self, x):
        return self.evaluate(np.asarray(x, dtype=float), *self.parameters)

    def __add__(self, other):
        return PSDSum(self, other)

    def device_spec(self):
        """``[(kind, p0, p1, p2)]`` for the GPU synthesis kernel, or None if host-only."""
        if self._device_kind is None:
            return None
        p = list(self.parameters) + [0.0, 0.0, 0.0]
        return [(self._device_kind, p[0], p[1], p[2])]

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__, ", ".join("%s=%g" % (k, getattr(self, k)) for k in self.param_names))


class PSDSum(PSDModel):
    def __init__(self, *parts):
        self.parts = []
        for p in parts:
            self.parts.extend(p.parts if isinstance(p, PSDSum) else [p])

    def __call__(self, x):
        x = np.asarray(x, dtype=float)
        return sum(p(x) for p in self.parts)

    def device_spec(self):
        out = []
        for p in self.parts:
            s = p.device_spec() if hasattr(p, "device_spec") else None
            if s is None:
                return None
            out.extend(s)
        return out

    def __repr__(self):
        return " + ".join(repr(p) for p in self.parts)


class SHO(PSDModel):
    """Eq. 20 of Foreman-Mackey+2017 (psd_models.py:7-11)."""
    param_names = ("S0", "Q", "omega0")
    defaults = (1, 10, 1)
    _device_kind = PSD_SHO

    def evaluate(self, x, S0, Q, omega0):
        return sqrt(2 / pi) * S0 * omega0 ** 4 / ((x ** 2 - omega0 ** 2) ** 2 + (x ** 2) * (omega0 ** 2) / Q ** 2)


class Lorentzian(PSDModel):
    """Eq. 11 of Foreman-Mackey+2017 (psd_models.py:14-32): S0 variance, Q coherence, omega0 peak."""
    param_names = ("S0", "Q", "omega0")
    defaults = (1, 10, 1)
    _device_kind = PSD_LORENTZIAN

    def evaluate(self, x, S0, Q, omega0):
        c = omega0 / 2 / Q
        return sqrt(1 / 2 / pi) * S0 / c * (1 / (1 + ((x - omega0) / c) ** 2) + 1 / (1 + ((x + omega0) / c) ** 2))


class BendingPowerlaw(PSDModel):
    """PSD of a damped random walk (psd_models.py:35-46)."""
    param_names = ("S0", "omega0", "Q")
    defaults = (1, 1, 0.5)
    _device_kind = PSD_BENDINGPL

    def evaluate(self, x, S0, omega0, Q):
        c = 0.5 * omega0 / Q
        return sqrt(2 / pi) * S0 / c * (1 / (1 + (x / c) ** 2))


def Matern(x, sigma=1.0, rho=1.0, n=1, nu=3 / 2):
    """General Matern PSD (psd_models.py:48-60); Matern-3/2 by default."""
    x = np.asarray(x, dtype=float)
    return (1 / sqrt(2 * pi) * sigma ** 2 * 2 ** n * pi ** (n / 2) * gamma(nu + n / 2) * (2 * nu) ** nu
            / (gamma(nu) * rho ** (2 * nu)) * (2 * nu / rho ** 2 + x ** 2) ** -(nu + n / 2))


class Matern32(PSDModel):
    """psd_models.py:63-67"""
    param_names = ("sigma", "rho", "n")
    defaults = (1, 1, 1)
    _device_kind = PSD_MATERN32

    def evaluate(self, x, sigma, rho, n):
        return 1 / sqrt(2 * pi) * sigma ** 2 * 4 / sqrt(3) * rho * (1 / (1 + (x * rho / sqrt(3)) ** 2)) ** 2


class Matern52(PSDModel):
    """psd_models.py:69-73"""
    param_names = ("sigma", "rho")
    defaults = (1, 1)
    _device_kind = PSD_MATERN52

    def evaluate(self, x, sigma, rho):
        return 1 / sqrt(2 * pi) * sigma ** 2 * 16 / 3 / sqrt(5) * rho * (1 / (1 + (x * rho / sqrt(5)) ** 2)) ** 3


class Jitter(PSDModel):
    """White-noise level (psd_models.py:75-85); depends on the frequency grid, host-only."""
    param_names = ("sigma",)
    defaults = (1,)

    def evaluate(self, x, sigma):
        N = len(x)
        df = np.diff(x)[0]
        return np.ones(N) * sigma ** 2 / (2 / sqrt(2 * pi)) / df / N
