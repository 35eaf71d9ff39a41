"""Lightcurve containers and file readers (reference ``mind_the_gaps/lightcurves/__init__.py:1-4``)."""
from .gappylightcurve import GappyLightcurve, ExposureTimeError
from .simplelightcurve import SimpleLightcurve
from .swiftlightcurve import SwiftLightcurve
from .fermilightcurve import FermiLightcurve
from .regular import Lightcurve

__all__ = ["GappyLightcurve", "ExposureTimeError", "SimpleLightcurve", "SwiftLightcurve", "FermiLightcurve", "Lightcurve"]
