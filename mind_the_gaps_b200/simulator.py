"""Lightcurve simulator (``/root/reference/mind_the_gaps/simulator.py``) with the arithmetic on the GPU.

Same constructor, attributes and methods as the reference ``Simulator``
(``sim_dt, sim_duration, sim_timestamps, fftndatapoints, strategy, noise, simulator, psd_model``;
``generate_lightcurve, simulate_regularly_sampled, downsample, add_noise, set_psd_params``) plus the
batched ``generate_lightcurves(nsims, parameters=None)`` the LRT uses.  Timmer & Koenig synthesis,
the random segment, the Emmanoulopoulos PDF adjustment, exposure binning and the noise models run in
``csrc/mtg_sim.cu`` / ``csrc/mtg_e13.cu`` behind ``mtg_sim_*`` (``include/mtg_b200.h``).

Randomness: Philox4x32-10 streams on the device, keyed by a seed drawn per call from
``np.random`` (or from ``RandomState(random_state)`` when given) -- the reference uses NumPy's global
MT19937, so sequences differ but ``np.random.seed`` still makes a run reproducible.  Every random
input can also be injected (``normals=``, ``u_segment=``, ``pdf_variates=``) for parity tests.
"""
import ctypes
import warnings

import numpy as np
from scipy.stats import norm

from . import _native
from .lightcurves.regular import Lightcurve
from .noise_models import GaussianNoise, KraftNoise, PoissonNoise
from .stats import create_log_normal, create_uniform_distribution

__all__ = ["Simulator", "TK95Simulator", "E13Simulator", "add_poisson_noise", "get_fft", "cut_random_segment", "get_segment"]


class BaseSimulatorMethod(object):
    def __init__(self, mean):
        self.meanrate = mean

    def adjust_pdf(self, segment):
        raise NotImplementedError("This method should be implemented by subclasses")


class TK95Simulator(BaseSimulatorMethod):
    """Gaussian PDF: nothing to adjust (simulator.py:34-40)."""

    def __init__(self, mean, random_state=None):
        super().__init__(mean)

    def adjust_pdf(self, segment):
        return segment


class E13Simulator(BaseSimulatorMethod):
    """Emmanoulopoulos et al. (2013) PDF adjustment (simulator.py:43-140)."""

    def __init__(self, mean, pdf, max_iter=1000, random_state=None):
        super().__init__(mean)
        self.pdf = pdf
        self.max_iter = max_iter
        if pdf == "lognormal":
            self.pdfmethod = create_log_normal
        elif pdf == "uniform":
            self.pdfmethod = create_uniform_distribution
        elif pdf == "gaussian":
            self.pdfmethod = norm
        else:
            raise ValueError("pdf must be one of 'lognormal', 'uniform', 'gaussian'")

    def adjust_lightcurve_pdf(self, lc, pdf, max_iter=400, xsim=None):
        """Iterate amplitude / rank adjustment on the device; ``xsim`` injects the ``pdf.rvs`` draws."""
        _native.require_device()
        n = lc.n
        if xsim is None:
            xsim = pdf.rvs(size=n)
        xsim = np.ascontiguousarray(xsim, dtype=np.float64)
        rate = np.ascontiguousarray(lc.countrate, dtype=np.float64)
        out = np.empty(n); iters = np.zeros(1, dtype=np.int32)
        _native.check(_native.lib().mtg_e13_adjust_host(n, _native.ptr(rate), _native.ptr(xsim), int(max_iter),
                                                        _native.ptr(out), _native.ptr(iters)))
        if iters[0] == max_iter:
            warnings.warn("Lightcurve did not converge after %d iterations, PDF might be inaccurate. "
                          "Try increase the maximum number of iterations" % iters[0])
        lc.countrate = out
        self.last_iterations = int(iters[0])
        return lc

    def adjust_pdf(self, segment):
        sample_std = np.std(segment.countrate)
        return self.adjust_lightcurve_pdf(segment, self.pdfmethod(self.meanrate, sample_std), max_iter=self.max_iter)


def _largest_prime_factor(v):
    f, big = 2, 1
    while f * f <= v:
        while v % f == 0:
            big, v = f, v // f
        f += 1
    return max(big, v) if v > 1 else big


def _fast_length(n):
    """Smallest ``N1 * N2 >= n`` with ``N1 = 2^p`` (``p <= 13``; ``N2 >= 256`` where ``n`` allows, so the padding stays
    below half a percent) and no prime factor of ``N2`` above 127.  This is the grid shape of the fused TK95 kernels
    (``csrc/mtg_tk95.cu``: column FFTs of length ``N1`` in shared memory, a direct sum over ``N2``); the bound on the
    prime factors keeps cuFFT, which the E13 / regular-lightcurve paths still call on this length, off Bluestein."""
    n = int(n)
    p = max(3, min(13, int(np.floor(np.log2(max(n, 2048) / 256.0)))))
    n1 = 1 << p
    n2 = -(-n // n1)
    while _largest_prime_factor(n2) > 127:
        n2 += 1
    return n1 * n2


_PDF = {"gaussian": _native.PDF_GAUSSIAN, "lognormal": _native.PDF_LOGNORMAL, "uniform": _native.PDF_UNIFORM}


class Simulator(object):
    """Simulate lightcurves with a given PSD and flux PDF on an arbitrary (gapped) observing window."""

    def __init__(self, psd_model, times, exposures, mean, pdf="gaussian", bkg_rate=None, bkg_rate_err=None,
                 sigma_noise=None, aliasing_factor=2, extension_factor=10, epsilon=1.001, max_iter=400,
                 random_state=None, fft_length="exact"):
        if extension_factor < 1:
            raise ValueError("Extension factor must be greater than 1")
        if epsilon < 1:
            raise ValueError("Epsilon needs to be greater than 1!")
        if np.any(np.asarray(exposures) == 0):
            raise ValueError("Some exposure times are 0!")
        times = np.asarray(times, dtype=np.float64)
        self._exposures = np.full(len(times), exposures, dtype=np.float64) if np.isscalar(exposures) \
            else np.asarray(exposures, dtype=np.float64)
        if pdf.lower() not in _PDF:
            raise ValueError("%s not implemented! Currently implemented: Gaussian, Uniform or Lognormal" % pdf)
        if pdf.lower() == "gaussian":
            self.simulator = TK95Simulator(mean)
        else:
            self.simulator = E13Simulator(mean, pdf.lower(), max_iter=max_iter)
        self.random_state = np.random.RandomState(random_state) if random_state is not None else None
        self.sim_dt = np.min(self._exposures) / aliasing_factor
        dt = np.diff(times)
        wrong = np.count_nonzero(dt < self.sim_dt * 0.99)
        if wrong > 0:
            raise ValueError("%d timestamps differences are below the exposure integration time! Either reduce "
                             "the exposure times, or space your observations" % wrong)
        start_time = times[0] - dt[0] / 1.99
        end_time = times[-1] + dt[-1]
        self.sim_duration = end_time - start_time
        duration = (times[-1] - times[0]) * extension_factor
        self.sim_timestamps = np.arange(start_time - self.sim_dt, start_time + duration + self.sim_dt, self.sim_dt)
        if fft_length == "fast" or isinstance(fft_length, (int, np.integer)):
            # NOT the reference's behaviour (opt-in): lengthen the regular grid to a length the device transforms
            # without Bluestein's algorithm (a real window easily gives a prime length) -- 2^p x N2, the shape of the
            # fused synthesis + binning kernels -- or to an explicit length (tests)
            n = _fast_length(len(self.sim_timestamps)) if fft_length == "fast" else int(fft_length)
            if n < len(self.sim_timestamps):
                raise ValueError("fft_length must be at least the reference's grid length %d" % len(self.sim_timestamps))
            self.sim_timestamps = (start_time - self.sim_dt) + np.arange(n) * self.sim_dt
        elif fft_length != "exact":
            raise ValueError("fft_length must be 'exact', 'fast' or an explicit length")
        self.fftndatapoints = len(self.sim_timestamps)
        self._lengthened = fft_length != "exact"
        self.pdf = pdf
        self.max_iter = max_iter
        self.psd_model = psd_model
        self._times = times
        if sigma_noise is None:
            if bkg_rate is None or np.all(np.asarray(bkg_rate) == 0):
                self.noise = PoissonNoise(self._exposures)
            else:
                self.noise = KraftNoise(self._exposures, np.asarray(bkg_rate) * self._exposures, bkg_rate_err)
        else:
            self.noise = GaussianNoise(self._exposures, sigma_noise)
        half_bins = self._exposures / 2 * epsilon
        self.strategy = [(time - half_bin, time + half_bin) for time, half_bin in zip(times, half_bins)]
        self.mean = mean
        self._handle = None
        self._keep = None

    def __str__(self):
        return "Simulator(\n  PSD Model: %s\n  PDF: %s\n) Noise: %s" % (self._psd_model, self.pdf, self.noise.name)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if getattr(self, "_handle", None):
            _native.lib().mtg_sim_destroy(self._handle)
            self._handle = None

    # -- PSD -------------------------------------------------------------------------------------------
    @property
    def psd_model(self):
        return self._psd_model

    @psd_model.setter
    def psd_model(self, new_psd_model):
        if not callable(new_psd_model):
            raise ValueError("PSD model must be callable (e.g., a function or Astropy model).")
        self._psd_model = new_psd_model

    def set_psd_params(self, psd_params):
        for par in psd_params.keys():
            setattr(self._psd_model, par, psd_params[par])

    def _psd_source(self, parameters=None):
        """``(model, params, table)``: a celerite kernel PSD is evaluated on the device from the kernel's
        parameters; anything else is evaluated once on the host and uploaded as a table."""
        from .celerite import GP
        from .celerite.terms import Term
        f = self._psd_model
        owner = getattr(f, "__self__", None)
        if isinstance(owner, Term) and getattr(f, "__name__", "") == "get_psd" and \
                all(t._mtg_kind is not None for t in owner.terms):
            model = GP(owner, mean=0.0).device_model()
            p = owner.get_parameter_vector() if parameters is None else parameters
            return model, np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64), None
        if parameters is not None:
            raise ValueError("per-lightcurve parameters need a celerite kernel PSD (kernel.get_psd)")
        M = self.fftndatapoints // 2 + 1
        omega = np.fft.rfftfreq(self.fftndatapoints, self.sim_dt) * 2 * np.pi
        table = np.zeros(M)
        table[1:] = np.asarray(f(omega[1:]), dtype=np.float64)
        return None, None, np.ascontiguousarray(table)

    # -- device object -----------------------------------------------------------------------------------
    def _sim(self):
        if self._handle is None:
            _native.require_device()
            cfg = _native.MtgSimConfig()
            n = len(self._times)
            grid = np.ascontiguousarray(self.sim_timestamps, dtype=np.float64)
            lo = np.ascontiguousarray([s[0] for s in self.strategy], dtype=np.float64)
            hi = np.ascontiguousarray([s[1] for s in self.strategy], dtype=np.float64)
            ex = np.ascontiguousarray(self._exposures, dtype=np.float64)
            cfg.n_obs = n; cfg.n_fft = self.fftndatapoints; cfg.sim_dt = float(self.sim_dt)
            cfg.sim_duration = float(self.sim_duration); cfg.mean = float(self.mean)
            cfg.pdf = _PDF[self.pdf.lower()]; cfg.max_iter = int(self.max_iter)
            cfg.flags = _native.SIM_E13_PADDED if self._lengthened else 0
            keep = [grid, lo, hi, ex]
            if isinstance(self.noise, GaussianNoise):
                cfg.noise = _native.NOISE_GAUSSIAN; cfg.sigma_noise = float(self.noise.sigma_noise)
            else:
                cfg.noise = _native.NOISE_KRAFT if isinstance(self.noise, KraftNoise) else _native.NOISE_POISSON
                cfg.kraft_counts = float(getattr(self.noise, "kraft_counts", 15))
                bk = np.ascontiguousarray(np.broadcast_to(np.asarray(self.noise.background_counts, dtype=np.float64), (n,)))
                er = np.ascontiguousarray(np.broadcast_to(np.asarray(self.noise.bkg_rate_err, dtype=np.float64), (n,)))
                keep += [bk, er]
                cfg.bkg_counts = bk.ctypes.data; cfg.bkg_rate_err = er.ctypes.data
            cfg.grid = grid.ctypes.data; cfg.bin_lo = lo.ctypes.data; cfg.bin_hi = hi.ctypes.data; cfg.exposures = ex.ctypes.data
            h = ctypes.c_void_p()
            _native.check(_native.lib().mtg_sim_create(ctypes.byref(cfg), ctypes.byref(h)))
            self._handle = h
        return self._handle

    @property
    def fused(self):
        """True when binned Gaussian-PDF lightcurves of this simulator come from the fused synthesis + binning kernels
        (``csrc/mtg_tk95.cu``) rather than the spectrum -> cuFFT -> binning chain."""
        return bool(_native.lib().mtg_sim_is_fused(self._sim()))

    def _seed(self):
        rs = self.random_state if self.random_state is not None else np.random
        return int(rs.randint(0, 2 ** 31 - 1)) * (2 ** 31) + int(rs.randint(0, 2 ** 31 - 1))

    # -- generation ------------------------------------------------------------------------------------------
    def generate_lightcurves(self, nsims, parameters=None, seed=None, normals=None, u_segment=None,
                             pdf_variates=None, return_info=False):
        """``nsims`` noiseless lightcurves binned to the observing window: array ``(nsims, N)``.

        ``parameters``: optional ``(nsims, ndim)`` kernel parameters, one PSD per lightcurve (the
        ``generate_from_posteriors`` path).  ``normals (nsims, 2, n_fft//2+1)``, ``u_segment (nsims,)`` and
        ``pdf_variates (nsims, n_fft)`` inject the random inputs."""
        h = self._sim()
        nsims = int(nsims)
        if parameters is not None:
            parameters = np.ascontiguousarray(np.atleast_2d(parameters), dtype=np.float64)
            if parameters.shape[0] != nsims:
                raise ValueError("parameters must have one row per simulation")
        model, params, table = self._psd_source(parameters)
        if model is not None and params.shape[0] == 1 and nsims > 1:
            params = np.ascontiguousarray(np.repeat(params, nsims, axis=0))
        c = lambda a, shape: None if a is None else _native.as_f64(a, shape)
        M = self.fftndatapoints // 2 + 1
        normals = c(normals, (nsims, 2, M)); u_segment = c(u_segment, (nsims,)); pdf_variates = c(pdf_variates, (nsims, self.fftndatapoints))
        out = np.empty((nsims, len(self._times))); seg = np.empty((nsims, 2), dtype=np.int32); its = np.zeros(nsims, dtype=np.int32)
        rc = _native.lib().mtg_sim_generate_host(h, nsims, ctypes.byref(model) if model is not None else None,
                                                 _native.ptr(params), _native.ptr(table),
                                                 int(self._seed() if seed is None else seed), _native.ptr(normals),
                                                 _native.ptr(u_segment), _native.ptr(pdf_variates), _native.ptr(out),
                                                 _native.ptr(seg), _native.ptr(its))
        _native.check(rc)
        if self.pdf.lower() != "gaussian" and np.any(its >= self.max_iter):
            warnings.warn("%d lightcurve(s) did not converge after %d iterations, PDF might be inaccurate. Try increase "
                          "the maximum number of iterations" % (int(np.sum(its >= self.max_iter)), self.max_iter))
        if return_info:
            return out, {"segment": seg, "e13_iterations": its}
        return out

    def generate_lightcurve(self):
        """One realisation: the rates at the input timestamps (simulator.py:397-420)."""
        return self.generate_lightcurves(1)[0]

    def simulate_regularly_sampled(self, seed=None, normals=None):
        """The long, finely and regularly sampled TK95 lightcurve (simulator.py:369-394)."""
        h = self._sim()
        model, params, table = self._psd_source()
        M = self.fftndatapoints // 2 + 1
        normals = None if normals is None else _native.as_f64(normals, (1, 2, M))
        out = np.empty((1, self.fftndatapoints))
        rc = _native.lib().mtg_sim_regular_host(h, 1, ctypes.byref(model) if model is not None else None, _native.ptr(params),
                                                _native.ptr(table), int(self._seed() if seed is None else seed),
                                                _native.ptr(normals), _native.ptr(out))
        _native.check(rc)
        return Lightcurve(self.sim_timestamps, out[0], input_counts=False, dt=self.sim_dt)

    def downsample(self, lc):
        """Mean count rate of ``lc`` inside every exposure bin of the sampling pattern (simulator.py:340-367)."""
        _native.require_device()
        t = np.ascontiguousarray(lc.time, dtype=np.float64)
        r = np.ascontiguousarray(lc.countrate, dtype=np.float64)
        lo = np.ascontiguousarray([s[0] for s in self.strategy], dtype=np.float64)
        hi = np.ascontiguousarray([s[1] for s in self.strategy], dtype=np.float64)
        out = np.empty(len(lo))
        _native.check(_native.lib().mtg_downsample_host(len(t), _native.ptr(t), _native.ptr(r), len(lo), _native.ptr(lo),
                                                        _native.ptr(hi), _native.ptr(out), None, None))
        return list(out)

    def downsample_indices(self, lc):
        """(first index, count) of the samples of ``lc`` falling in every bin."""
        _native.require_device()
        t = np.ascontiguousarray(lc.time, dtype=np.float64)
        r = np.ascontiguousarray(lc.countrate, dtype=np.float64)
        lo = np.ascontiguousarray([s[0] for s in self.strategy], dtype=np.float64)
        hi = np.ascontiguousarray([s[1] for s in self.strategy], dtype=np.float64)
        out = np.empty(len(lo)); first = np.empty(len(lo), dtype=np.int32); count = np.empty(len(lo), dtype=np.int32)
        _native.check(_native.lib().mtg_downsample_host(len(t), _native.ptr(t), _native.ptr(r), len(lo), _native.ptr(lo),
                                                        _native.ptr(hi), _native.ptr(out), _native.ptr(first), _native.ptr(count)))
        return first, count

    def add_noise(self, rates, **inject):
        """``(noisy_rates, dy)`` for one lightcurve ``(N,)`` or a batch ``(L, N)`` (simulator.py:300-338)."""
        if "seed" not in inject:
            inject["seed"] = self._seed()
        return self.noise.add_noise(rates, **inject)


def add_poisson_noise(rates, exposures, background_counts=None, bkg_rate_err=None):
    """Legacy module-level helper of the reference (simulator.py:423-465): Poisson-distributed total counts drawn with
    ``np.random.poisson`` on the host (so a seeded NumPy run reproduces the reference's numbers), net rates and their
    uncertainties.  ``Simulator.add_noise`` / ``noise_models.PoissonNoise`` are the device path."""
    rates = np.asarray(rates, dtype=np.float64)
    if background_counts is None:
        background_counts = np.zeros(len(rates), dtype=int)
    if bkg_rate_err is None:
        bkg_rate_err = np.zeros(len(rates), dtype=int)
    drawn = np.random.poisson(rates * exposures + background_counts)
    dy = np.sqrt((np.sqrt(drawn) / exposures) ** 2 + np.asarray(bkg_rate_err) ** 2)
    return (drawn - background_counts) / exposures, dy


def get_fft(N, dt, model, real=None, im=None):
    """Host restatement of the TK95 half-spectrum (simulator.py:468-501), kept for API compatibility;
    the GPU path builds it in ``k_spectrum``."""
    freqs = np.fft.rfftfreq(N, dt) * 2 * np.pi
    if real is None:
        real, im = np.random.normal(0, size=(2, N // 2 + 1))
    complex_fft = np.empty(len(freqs), dtype=complex)
    complex_fft[1:] = (real + im * 1j)[1:] * np.sqrt(.5 * model(freqs[1:]))
    complex_fft[0] = 1e6
    if N % 2 == 0:
        complex_fft[-1] = np.real(complex_fft[-1])
    return complex_fft


def get_segment(lc, duration, N):
    if N < 0:
        raise ValueError("N must be a non-negative integer.")
    start = lc.time[0] + duration * N
    return lc.truncate(start=start, stop=start + duration, method="time")


def cut_random_segment(lc, duration):
    shift = np.random.uniform(lc.time[0], lc.time[-1] - duration)
    return lc.truncate(start=shift, stop=shift + duration, method="time")
