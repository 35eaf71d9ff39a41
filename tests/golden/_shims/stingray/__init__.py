"""stingray stand-in: the Lightcurve members mind_the_gaps/simulator.py touches (golden generation only).

Restated from stingray 2.x ``lightcurve.py``: ``countrate = counts / dt`` for ``input_counts=True``;
``meanrate = mean(countrate)``; ``tstart = time[0] - dt / 2``; ``truncate(start, stop, method="time")``
slices ``[time.searchsorted(start), time.searchsorted(stop))``; ``shift(x)`` adds ``x`` to the times."""
import copy
import numpy as np


class Lightcurve:
    def __init__(self, time, counts, input_counts=True, skip_checks=True, dt=None, err_dist="gauss", **kw):
        self.time = np.asarray(time, dtype=float)
        self.dt = float(dt) if dt is not None else float(np.median(np.diff(self.time)))
        counts = np.asarray(counts, dtype=float)
        self._countrate = counts / self.dt if input_counts else counts
        self.tstart = self.time[0] - 0.5 * self.dt

    @property
    def countrate(self):
        return self._countrate

    @countrate.setter
    def countrate(self, v):
        self._countrate = np.asarray(v, dtype=float)

    @property
    def counts(self):
        return self._countrate * self.dt

    @property
    def meanrate(self):
        return np.mean(self._countrate)

    @property
    def n(self):
        return self.time.shape[0]

    def truncate(self, start=0, stop=None, method="index"):
        if method == "time":
            i0 = self.time.searchsorted(start)
            i1 = self.time.searchsorted(stop)
        else:
            i0, i1 = start, stop
        new = copy.copy(self)
        new.time = self.time[i0:i1]
        new._countrate = self._countrate[i0:i1]
        new.tstart = new.time[0] - 0.5 * self.dt
        return new

    def shift(self, time_shift):
        new = copy.copy(self)
        new.time = self.time + time_shift
        new.tstart = self.tstart + time_shift
        return new
