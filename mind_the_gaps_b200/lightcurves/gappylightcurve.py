"""Irregularly sampled lightcurve container (reference
``mind_the_gaps/lightcurves/gappylightcurve.py:20-171, 265-290``)."""
import numpy as np


class ExposureTimeError(Exception):
    pass


class GappyLightcurve(object):
    """times [s], rates ``y``, 1-sigma errors ``dy``, exposures [s], background rate and its error."""

    def __init__(self, times, y, dy=None, exposures=None, bkg_rate=None, bkg_rate_err=None):
        self._times = np.asarray(times)
        self._y = np.asarray(y)
        self._dy = None if dy is None else np.asarray(dy)
        n = len(self._times)
        if exposures is not None:
            self._exposures = np.full(n, exposures) if np.isscalar(exposures) else np.asarray(exposures)
            # same consistency check as the reference (gappylightcurve.py:60-64), epsilon = 1.01
            wrong = np.count_nonzero(np.diff(self._times) < self._exposures[:-1] * 1.01 / 2)
            if wrong > 0:
                raise ExposureTimeError("Some timestamps (%d) have a spacing below the exposure sampling time!" % wrong)
        else:
            self._exposures = np.zeros(n)
        self._bkg_rate = np.zeros(n) if bkg_rate is None else np.asarray(bkg_rate)
        self._bkg_rate_err = np.zeros(n) if bkg_rate_err is None else np.asarray(bkg_rate_err)

    times = property(lambda self: self._times)
    y = property(lambda self: self._y)
    dy = property(lambda self: self._dy)
    exposures = property(lambda self: self._exposures)
    bkg_rate = property(lambda self: self._bkg_rate)
    bkg_rate_err = property(lambda self: self._bkg_rate_err)

    @property
    def n(self):
        return len(self._times)

    @property
    def duration(self):
        return self._times[-1] - self._times[0]

    @property
    def mean(self):
        return np.mean(self._y)

    def _masked(self, mask):
        return GappyLightcurve(self._times[mask], self._y[mask], None if self._dy is None else self._dy[mask],
                               self._exposures[mask], self._bkg_rate[mask], self._bkg_rate_err[mask])

    def truncate(self, tmin=-np.inf, tmax=np.inf):
        if tmin >= tmax:
            raise ValueError("Minimum truncation time (%.2es) is greater than or equal to maximum truncation time (%.3es)!" % (tmin, tmax))
        if tmax < self._times[0]:
            raise ValueError("Maximum truncation time (%.2f) is lower than initial lightcurve time (%.2f)" % (tmax, self._times[0]))
        return self._masked((self._times >= tmin) & (self._times <= tmax))

    def split(self, interval):
        cuts = np.append(np.where(np.diff(self._times) > interval)[0], -1)
        out, j = [], 0
        for i in cuts:
            out.append(self.truncate(self._times[j], self._times[i]))
            j = i + 1
        return out

    def rand_remove(self, points_remove):
        if points_remove > self.n:
            raise ValueError("Number of points to remove (%d) is greater than number of lightcurve datapoints (%d)" % (points_remove, self.n))
        mask = np.ones(self.n, dtype=bool)
        mask[np.random.choice(self.n, points_remove, replace=False)] = False
        return self._masked(mask)

    def to_csv(self, outname):
        outputs = np.array([self._times, self._y, self._dy, self._exposures, self._bkg_rate, self._bkg_rate_err])
        np.savetxt(outname, outputs.T, fmt="%.8e\t%.5f\t%.5f\t%.3f\t%.5f\t%.5f", header="t\trate\terror\texposure\tbkg_rate\tbkg_rate_err")

    def get_simulator(self, psd_model, pdf="gaussian", **kwargs):
        """A ``Simulator`` for this observing window (gappylightcurve.py:265-290)."""
        from ..simulator import Simulator
        return Simulator(psd_model, self._times, self._exposures, self.mean, pdf, self._bkg_rate, self._bkg_rate_err, **kwargs)
