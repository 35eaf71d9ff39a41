"""Same export surface as ``/root/reference/mind_the_gaps/models/__init__.py:1-2``, plus the
analytic PSD models the reference's notebooks import from here (``tutorial_ppp.ipynb`` cell 0)."""
from .mean_models import LinearModel, GaussianModel, SineModel
from .celerite_models import Lorentzian, Cosinus, DampedRandomWalk, BendingPowerlaw
from . import psd_models

__all__ = ["LinearModel", "GaussianModel", "SineModel", "Lorentzian", "Cosinus", "DampedRandomWalk",
           "BendingPowerlaw", "psd_models"]
