"""pyfftw stand-in: pyfftw.interfaces.numpy_fft is a drop-in for numpy.fft (golden generation only)."""
import types
import numpy as _np

interfaces = types.ModuleType("pyfftw.interfaces")
interfaces.numpy_fft = _np.fft
interfaces.cache = types.SimpleNamespace(enable=lambda: None, disable=lambda: None)
