"""Oracle: Lomb-Scargle periodogram, the O(N M) textbook evaluation.  TEST INFRASTRUCTURE ONLY.

Follows ``astropy.timeseries.periodograms.lombscargle.implementations.slow_impl.lombscargle_slow`` (un-vendored
dependency of the reference's ``docs/notebooks/lomb_scargle_biases.ipynb``; algorithm published in Zechmeister & Kuerster
2009 and VanderPlas 2018, sec. 6) with unit errors.  astropy is not importable here: "parity unpinned vs astropy"; the
restatement is pinned on ``scipy.signal.lombscargle(..., floating_mean=...)`` (tests/test_periodogram_gpu.py /
test_oracle.py), an independent implementation of the same estimator."""
import numpy as np


def lombscargle(t, y, frequency, fit_mean=True, center_data=True, normalization="standard"):
    t = np.asarray(t, dtype=float); y = np.asarray(y, dtype=float)
    w = np.full(len(t), 1.0 / len(t))
    if fit_mean or center_data:
        y = y - np.dot(w, y)
    omega = 2 * np.pi * np.asarray(frequency, dtype=float)[None, :]
    tt, yy, ww = t[:, None], y[:, None], w[:, None]
    s, c = np.sin(omega * tt), np.cos(omega * tt)
    S2 = 2 * np.dot(ww.T, s * c)
    C2 = 2 * np.dot(ww.T, 0.5 - s ** 2)
    if fit_mean:
        S, C = np.dot(ww.T, s), np.dot(ww.T, c)
        S2 -= 2 * S * C
        C2 -= C * C - S * S
    wtt = omega * tt - 0.5 * np.arctan2(S2, C2)
    st, ct = np.sin(wtt), np.cos(wtt)
    Y = np.dot(ww.T, yy)
    wy = ww * yy
    YC, YS = np.dot(wy.T, ct), np.dot(wy.T, st)
    CC, SS = np.dot(ww.T, ct * ct), np.dot(ww.T, st * st)
    if fit_mean:
        Ct, St = np.dot(ww.T, ct), np.dot(ww.T, st)
        YC -= Y * Ct; YS -= Y * St
        CC -= Ct * Ct; SS -= St * St
    p = YC * YC / CC + YS * YS / SS
    YY = np.dot(ww.T, yy * yy)
    if normalization == "standard":
        p /= YY
    elif normalization == "model":
        p /= YY - p
    elif normalization == "log":
        p = -np.log(1 - p / YY)
    elif normalization == "psd":
        p *= 0.5 * len(t)
    else:
        raise ValueError(normalization)
    return p.ravel()
