"""``GappyLightcurve`` container (reference ``mind_the_gaps/lightcurves/__init__.py:1-4``; the
instrument-specific file readers are out of scope)."""
from .gappylightcurve import GappyLightcurve, ExposureTimeError
from .regular import Lightcurve

__all__ = ["GappyLightcurve", "ExposureTimeError", "Lightcurve"]
