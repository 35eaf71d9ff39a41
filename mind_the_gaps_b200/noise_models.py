"""Noise models (``/root/reference/mind_the_gaps/noise_models.py``): same classes, same constructor
arguments, ``add_noise(rates) -> (noisy_rates, dy)``; the arithmetic runs in the ``k_noise`` CUDA kernel
(``csrc/mtg_sim.cu``) through ``mtg_noise_host``.

Randomness: the reference draws from NumPy's global MT19937; here each call takes a 63-bit Philox seed
from ``np.random`` (so ``np.random.seed`` still makes runs reproducible) unless the caller injects the
random inputs (``counts=`` / ``normals=``) for parity tests.
"""
import numpy as np

from . import _native

__all__ = ["BaseNoise", "PoissonNoise", "KraftNoise", "GaussianNoise"]


def _seed():
    return int(np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(np.random.randint(0, 2 ** 31 - 1))


def _run(kind, rates, exposures, bkg_counts, bkg_rate_err, sigma, kraft_counts, seed, counts, normals):
    _native.require_device()
    r = np.ascontiguousarray(rates, dtype=np.float64)
    squeeze = r.ndim == 1
    r2 = np.atleast_2d(r)
    L, N = r2.shape
    f = lambda a: None if a is None else np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (N,)))
    ex, bk, er = f(exposures), f(bkg_counts), f(bkg_rate_err)
    c = None if counts is None else np.ascontiguousarray(np.atleast_2d(counts), dtype=np.int64)
    z = None if normals is None else np.ascontiguousarray(np.atleast_2d(normals), dtype=np.float64)
    out = np.empty((L, N)); dy = np.empty((L, N)); status = np.empty((L, N), dtype=np.int32)
    rc = _native.lib().mtg_noise_host(kind, L, N, _native.ptr(r2), _native.ptr(ex), _native.ptr(bk), _native.ptr(er),
                                      float(sigma), float(kraft_counts), int(seed if seed is not None else _seed()),
                                      _native.ptr(c), _native.ptr(z), _native.ptr(out), _native.ptr(dy), _native.ptr(status))
    _native.check(rc)
    if np.any(status == 1):
        raise ValueError("lam < 0 or lam contains NaNs")       # what np.random.poisson raises in the reference
    return (out[0], dy[0]) if squeeze else (out, dy)


class BaseNoise(object):
    def __init__(self, name):
        self.name = name

    def add_noise(self, rates):
        raise NotImplementedError("This method should be implemented by subclasses")


class PoissonNoise(BaseNoise):
    """Poisson counts on ``rate * exposure + background`` (noise_models.py:29-78)."""

    def __init__(self, exposures, background_counts=None, bkg_rate_err=None):
        super().__init__(name="Poisson")
        self.exposures = exposures
        self.background_counts = np.zeros(len(exposures), dtype=int) if background_counts is None else background_counts
        self.bkg_rate_err = np.zeros(len(exposures)) if bkg_rate_err is None else bkg_rate_err

    def add_noise(self, rates, seed=None, counts=None):
        return _run(_native.NOISE_POISSON, rates, self.exposures, self.background_counts, self.bkg_rate_err, 0.0, 15.0,
                    seed, counts, None)


class KraftNoise(PoissonNoise):
    """Poisson noise with the Kraft, Burrows & Nousek (1991) posterior median / 68 % interval for bins
    below ``kraft_counts`` total counts (noise_models.py:81-150)."""

    def __init__(self, exposures, background_counts=None, bkg_rate_err=None, kraft_counts=15):
        super().__init__(exposures, background_counts, bkg_rate_err)
        self.name = "Kraft"
        self.kraft_counts = kraft_counts

    def add_noise(self, rates, seed=None, counts=None):
        return _run(_native.NOISE_KRAFT, rates, self.exposures, self.background_counts, self.bkg_rate_err, 0.0,
                    self.kraft_counts, seed, counts, None)


class GaussianNoise(BaseNoise):
    """Zero-mean Gaussian white noise of standard deviation ``sigma_noise`` (noise_models.py:152-184)."""

    def __init__(self, exposures, sigma_noise):
        super().__init__(name="Gaussian")
        self.sigma_noise = sigma_noise

    def add_noise(self, rates, seed=None, normals=None):
        return _run(_native.NOISE_GAUSSIAN, rates, None, None, None, self.sigma_noise, 15.0, seed, None, normals)
