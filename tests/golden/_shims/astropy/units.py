"""astropy.units stand-in (golden generation only): time units with ``array << unit``, ``.to(unit).value``."""
import numpy as np


class Quantity:
    def __init__(self, value, unit):
        self.value = np.asarray(value, dtype=float)
        self.unit = unit

    def to(self, unit):
        if self.unit.scale is None or unit.scale is None:
            raise ValueError("not a time unit")
        # astropy multiplies by the float ratio of the unit scales
        return Quantity(self.value * (self.unit.scale / unit.scale), unit)

    def __len__(self):
        return len(self.value)


class _Unit:
    __array_ufunc__ = None          # let ``ndarray << unit`` reach __rlshift__

    def __init__(self, scale=None):
        self.scale = scale           # seconds, for the time units

    def __rlshift__(self, o): return Quantity(o, self)
    def __rmul__(self, o): return o
    def __mul__(self, o): return o
    def __truediv__(self, o): return self
    def __rtruediv__(self, o): return o


s = _Unit(1.0)
d = _Unit(86400.0)
Hz = ct = erg = cm = _Unit()
