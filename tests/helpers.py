"""Shared synthetic inputs for the tests (SURVEY.md 8d window generator)."""
import numpy as np


def window(N, seed):
    """Swift-like gapped observing window: seasonal gaps, random daily sampling.  Times in seconds."""
    rng = np.random.default_rng(seed)
    base = np.arange(0, 4 * N)
    base = base[(base % 365) < 240]
    picked = np.sort(rng.choice(base, size=N, replace=False))
    t = (picked + rng.uniform(-0.3, 0.3, N)) * 86400.0
    exposures = rng.uniform(500.0, 2000.0, N)
    return t, exposures


def fake_lightcurve(N, seed, mean=0.5):
    t, exposures = window(N, seed)
    rng = np.random.default_rng(seed + 1000)
    y = mean * (1.0 + 0.1 * np.cumsum(rng.normal(0, 0.1, N))) + rng.normal(0, 0.02, N)
    dy = rng.uniform(0.01, 0.03, N)
    return t, y, dy, exposures


def c2_truth():
    """DRW + Lorentzian truth of config C2 (tutorial_ppp.ipynb cell 9), per-second units."""
    mean = 0.5
    day = 86400.0
    var = (0.1 * mean) ** 2
    w_bend = 2 * np.pi / (20 * day)
    w_qpo = 2 * np.pi / (10 * day)
    return np.array([np.log(var), np.log(w_bend), np.log(var), np.log(80.0), np.log(w_qpo)])


C2_BOUNDS = [(-30.0, 30.0), (-25.0, 5.0), (-30.0, 30.0), (-5.0, 10.0), (-25.0, 5.0)]
