"""``FermiLightcurve``: a ``GappyLightcurve`` from a comma-separated Fermi-LAT light-curve table
(reference ``mind_the_gaps/lightcurves/fermilightcurve.py:12-44``).

Columns: the first is the time, the second the flux ``<name>``; the asymmetric errors are the columns
``<name>_err_neg`` and ``<name>_err_pos`` and the symmetric error handed on is their mean magnitude.
The reference's day/second test (``"MJD" in data`` on the structured array, fermilightcurve.py:38) never
matches a column name, so times are always taken as seconds -- kept, because a drop-in must return the
same numbers; convert beforehand if the file is in MJD.  No exposures: Gaussian noise in the simulator.
"""
import numpy as np

from ._table import read_named, to_seconds
from .gappylightcurve import GappyLightcurve


class FermiLightcurve(GappyLightcurve):
    def __init__(self, input_file):
        time, y, yerr = self.readdata(input_file)
        super().__init__(time, y, yerr)

    def readdata(self, input_file):
        table = read_named(input_file, delimiter=",")
        names = table.dtype.names
        flux = names[1]
        yerr = (np.abs(table["%s_err_neg" % flux]) + table["%s_err_pos" % flux]) / 2
        return to_seconds(table[names[0]], False), table[flux], yerr
