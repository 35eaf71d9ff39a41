"""Mean models (``/root/reference/mind_the_gaps/models/mean_models.py``).  ``LinearModel`` and
``GaussianModel`` are the ones ``GPModelling._build_mean_model`` uses (gpmodelling.py:98-124) and
the CUDA kernels evaluate; the sine models are host-only helpers kept for import compatibility."""
import numpy as np

from ..celerite.modeling import Model

__all__ = ["GaussianModel", "SineModel", "TwoSineModel", "LinearModel"]


class GaussianModel(Model):
    parameter_names = ("mean", "sigma", "amplitude", "constant")

    def get_value(self, x):
        x = np.asarray(x, dtype=float)
        return self.amplitude / (2 * np.pi * self.sigma) * np.exp(-(x - self.mean) ** 2 / (2 * self.sigma ** 2)) + self.constant


class SineModel(Model):
    parameter_names = ("constant", "amplitude", "frequency", "phase")

    def get_value(self, x):
        return self.constant + self.amplitude * np.sin(self.frequency * np.asarray(x, dtype=float) + self.phase)


class TwoSineModel(Model):
    parameter_names = ("constant", "amplitude0", "phase0", "amplitude1", "phase1", "frequency")

    def get_value(self, x):
        x = np.asarray(x, dtype=float)
        return (self.constant + self.amplitude0 * np.sin(self.frequency * x + self.phase0)
                + self.amplitude1 * np.sin(2 * self.frequency * x + self.phase1))


class LinearModel(Model):
    parameter_names = ("slope", "intercept")

    def get_value(self, x):
        return self.slope * np.asarray(x, dtype=float) + self.intercept

    def compute_gradient(self, x):
        x = np.asarray(x, dtype=float)
        return np.array([x, np.ones_like(x)])
