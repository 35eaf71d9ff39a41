"""Mean models of the GP (reference ``mind_the_gaps/models/mean_models.py:6-31``).

``LinearModel`` and ``GaussianModel`` are what ``GPModelling._build_mean_model`` instantiates (gpmodelling.py:98-124); the
CUDA kernels evaluate them per observation from the parameter vector (``csrc/mtg_common.cuh: MeanEval``, same
expressions).  The sine models are host-side helpers kept for import compatibility.  Every model also provides the
analytic ``compute_gradient`` of the celerite ``Model`` protocol (rows ordered as ``parameter_names``)."""
import numpy as np

from ..celerite.modeling import Model

__all__ = ["GaussianModel", "SineModel", "TwoSineModel", "LinearModel"]

_TWO_PI = 2.0 * np.pi


def _times(x):
    return np.asarray(x, dtype=np.float64)


class GaussianModel(Model):
    """``constant + amplitude / (2 pi sigma) exp(-(t - mean)^2 / (2 sigma^2))`` -- the reference's normalisation
    (``2 pi sigma``, not ``sqrt(2 pi) sigma``) is kept: a drop-in must return the same numbers."""
    parameter_names = ("mean", "sigma", "amplitude", "constant")

    def _bump(self, t):
        z = (t - self.mean) / self.sigma
        return np.exp(-0.5 * z * z) / (_TWO_PI * self.sigma), z

    def get_value(self, x):
        t = _times(x)
        # written as the reference evaluates it, so that host and reference agree to the last bit
        return self.amplitude / (2 * np.pi * self.sigma) * np.exp(-(t - self.mean) ** 2 / (2 * self.sigma ** 2)) + self.constant

    def compute_gradient(self, x):
        t = _times(x)
        g, z = self._bump(t)
        peak = self.amplitude * g
        return np.array([peak * z / self.sigma, peak * (z * z - 1.0) / self.sigma, g, np.ones_like(t)])


class SineModel(Model):
    """``constant + amplitude sin(frequency t + phase)`` (``frequency`` is angular)."""
    parameter_names = ("constant", "amplitude", "frequency", "phase")

    def get_value(self, x):
        t = _times(x)
        return self.constant + self.amplitude * np.sin(self.frequency * t + self.phase)

    def compute_gradient(self, x):
        t = _times(x)
        arg = self.frequency * t + self.phase
        c = self.amplitude * np.cos(arg)
        return np.array([np.ones_like(t), np.sin(arg), c * t, c])


class TwoSineModel(Model):
    """Fundamental plus first harmonic of one angular ``frequency``."""
    parameter_names = ("constant", "amplitude0", "phase0", "amplitude1", "phase1", "frequency")

    def get_value(self, x):
        t = _times(x)
        first = self.amplitude0 * np.sin(self.frequency * t + self.phase0)
        second = self.amplitude1 * np.sin(2 * self.frequency * t + self.phase1)
        return self.constant + first + second

    def compute_gradient(self, x):
        t = _times(x)
        a0, a1 = self.frequency * t + self.phase0, 2 * self.frequency * t + self.phase1
        c0, c1 = self.amplitude0 * np.cos(a0), self.amplitude1 * np.cos(a1)
        return np.array([np.ones_like(t), np.sin(a0), c0, np.sin(a1), c1, t * (c0 + 2 * c1)])


class LinearModel(Model):
    """``slope t + intercept``."""
    parameter_names = ("slope", "intercept")

    def get_value(self, x):
        return self.slope * _times(x) + self.intercept

    def compute_gradient(self, x):
        t = _times(x)
        return np.array([t, np.ones_like(t)])
