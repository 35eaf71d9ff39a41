"""``SimpleLightcurve``: a ``GappyLightcurve`` from a delimited text file with a header line
(reference ``mind_the_gaps/lightcurves/simplelightcurve.py:12-59``; the 6-column tab format is what
``GappyLightcurve.to_csv`` writes, ``gappylightcurve.py:256-262``).

Columns by position: time, rate, error[, exposure[, bkg_rate, bkg_rate_err]].  A time column called
``mjd``, ``jd`` or ``day`` is in days, anything else in seconds.
"""
import warnings

import numpy as np

from ._table import read_named, to_seconds
from .gappylightcurve import GappyLightcurve

DAY_COLUMNS = ("mjd", "jd", "day")


class SimpleLightcurve(GappyLightcurve):
    def __init__(self, input_file, skip_header=0, delimiter=None):
        time, y, yerr, exposures, bkg_rate, bkg_err = self.readdata(input_file, skip_header, delimiter)
        super().__init__(time, y, yerr, exposures, bkg_rate, bkg_err)

    def readdata(self, input_file, skip_header=0, delimiter=None):
        """``(time [s], rate, error, exposures, bkg_rate, bkg_rate_err)``; missing columns come back as zeros."""
        table = read_named(input_file, skip_header=skip_header, delimiter=delimiter)
        names = table.dtype.names
        if len(names) < 3:
            raise ValueError("%s: need at least time, rate and error columns (found %d)" % (input_file, len(names)))
        in_days = names[0] in DAY_COLUMNS
        print("Time in days" if in_days else "Time in seconds")
        time = to_seconds(table[names[0]], in_days)
        n = len(time)
        exposures, bkg_rate, bkg_err = np.zeros(n), np.zeros(n), np.zeros(n)
        if len(names) > 3:
            print("Reading exposures")
            exposures = table[names[3]]
            if len(names) >= 6:
                print("Reading background rates")
                bkg_rate, bkg_err = table[names[4]], table[names[5]]
        else:
            warnings.warn("Lightcurve has no exposures!")
        return time, table[names[1]], table[names[2]], exposures, bkg_rate, bkg_err
