"""Delimited-text helpers shared by the lightcurve file readers (host side, numpy only)."""
import numpy as np

SECONDS_PER_DAY = 86400.0


def read_named(path, **kwargs):
    """Structured array with the column names of the file's header line (``numpy.genfromtxt(names=True)``,
    which is what the reference's readers call: name sanitising and comment handling are numpy's)."""
    return np.atleast_1d(np.genfromtxt("%s" % path, names=True, **kwargs))


def to_seconds(values, in_days):
    values = np.asarray(values, dtype=np.float64)
    return values * SECONDS_PER_DAY if in_days else values * 1.0
