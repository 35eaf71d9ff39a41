"""Batched periodograms of lightcurve ensembles on the GPU.

The reference computes them one lightcurve at a time in Python loops -- ``numpy.fft.rfft`` in its simulator tests
(``tests/simulator_test.py:24-35``) and ``astropy.timeseries.LombScargle(t, rate, fit_mean=True, center_data=True,
normalization=...).power(frequencies)`` in ``docs/notebooks/lomb_scargle_biases.ipynb`` -- before handing them to
``fitting.fit_psd_powerlaw`` / ``fit_lomb_scargle``.  Here a whole ensemble sharing its timestamps is one device call
(``csrc/mtg_periodogram.cu``)."""
import numpy as np

from . import _native

__all__ = ["power_spectrum", "lomb_scargle"]

_NORMALIZATIONS = {"standard": 0, "model": 1, "log": 2, "psd": 3}


def power_spectrum(timestamps, rates):
    """``(frequencies, powers)`` of regularly sampled lightcurves ``rates`` (``(n,)`` or ``(L, n)``): ``|rfft|^2`` without
    the zero frequency, and without the Nyquist frequency when the number of frequencies is even -- exactly the helper of
    the reference's tests (``tests/simulator_test.py:24-35``; Vaughan 2005)."""
    _native.require_device()
    timestamps = np.asarray(timestamps, dtype=np.float64)
    r = np.ascontiguousarray(np.atleast_2d(rates), dtype=np.float64)
    L, n = r.shape
    if n != len(timestamps):
        raise ValueError("rates and timestamps differ in length")
    dt = np.mean(np.diff(timestamps))
    freqs = np.fft.rfftfreq(n, dt)
    out = np.empty((L, n // 2 + 1))
    _native.check(_native.lib().mtg_power_spectrum_host(L, n, _native.ptr(r), _native.ptr(out)))
    sl = slice(1, -1) if len(freqs) % 2 == 0 else slice(1, None)
    p = out[:, sl]
    return freqs[sl], (p[0] if np.ndim(rates) == 1 else p)


def lomb_scargle(t, y, frequencies, fit_mean=True, center_data=True, normalization="standard"):
    """Lomb-Scargle powers of ``y`` (``(N,)`` or ``(L, N)``, all on the times ``t``) at ``frequencies``: the values of
    ``astropy.timeseries.LombScargle(t, y_l, fit_mean=fit_mean, center_data=center_data,
    normalization=normalization).power(frequencies)`` for every row ``l`` (unit errors; generalised / floating-mean
    periodogram when ``fit_mean``)."""
    _native.require_device()
    if normalization not in _NORMALIZATIONS:
        raise ValueError("normalization='%s' not recognized" % normalization)
    t = np.ascontiguousarray(t, dtype=np.float64)
    Y = np.ascontiguousarray(np.atleast_2d(y), dtype=np.float64)
    f = np.ascontiguousarray(np.atleast_1d(frequencies), dtype=np.float64)
    L, N = Y.shape
    if N != len(t):
        raise ValueError("y and t differ in length")
    out = np.empty((L, len(f)))
    _native.check(_native.lib().mtg_lombscargle_host(N, L, len(f), _native.ptr(t), _native.ptr(Y), _native.ptr(f), int(bool(fit_mean)),
                                                     int(bool(center_data)), _NORMALIZATIONS[normalization], _native.ptr(out)))
    return out[0] if np.ndim(y) == 1 else out
