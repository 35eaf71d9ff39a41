class Quantity(float):
    pass


class _Unit:
    def __rmul__(self, o): return o
    def __mul__(self, o): return o
    def __truediv__(self, o): return self
    def __rtruediv__(self, o): return o


s = d = Hz = ct = erg = cm = _Unit()
