"""Minimal regularly-sampled lightcurve: the five ``stingray.Lightcurve`` semantics the reference's
simulator relies on (``/root/reference/mind_the_gaps/simulator.py:391-393, 412-414, 533, 539``):
``countrate = counts / dt``, ``meanrate``, ``tstart = time[0] - dt/2``, ``truncate(method="time")`` by
left-sided ``searchsorted`` and ``shift``."""
import numpy as np


class Lightcurve(object):
    def __init__(self, time, counts, input_counts=True, dt=None, skip_checks=True, err_dist="gauss", **kwargs):
        self.time = np.asarray(time, dtype=np.float64)
        self.dt = float(dt) if dt is not None else float(np.median(np.diff(self.time)))
        v = np.asarray(counts, dtype=np.float64)
        self._countrate = v / self.dt if input_counts else v.copy()

    @property
    def countrate(self):
        return self._countrate

    @countrate.setter
    def countrate(self, value):
        self._countrate = np.asarray(value, dtype=np.float64)

    @property
    def counts(self):
        return self._countrate * self.dt

    @property
    def n(self):
        return len(self.time)

    @property
    def meanrate(self):
        return np.mean(self.counts) / self.dt

    @property
    def tstart(self):
        return self.time[0] - 0.5 * self.dt

    @property
    def tseg(self):
        return self.time[-1] - self.time[0] + self.dt

    def truncate(self, start=0, stop=None, method="index"):
        if method == "time":
            i0 = int(np.searchsorted(self.time, start))
            i1 = int(np.searchsorted(self.time, stop)) if stop is not None else len(self.time)
        else:
            i0, i1 = int(start), (len(self.time) if stop is None else int(stop))
        return Lightcurve(self.time[i0:i1], self._countrate[i0:i1], input_counts=False, dt=self.dt)

    def shift(self, time_shift):
        return Lightcurve(self.time + time_shift, self._countrate, input_counts=False, dt=self.dt)
