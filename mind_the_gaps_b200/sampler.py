"""Affine-invariant ensemble sampler with the slice of the ``emcee`` 3.1.4 API that ``GPModelling``
uses (``/root/reference/mind_the_gaps/gpmodelling.py:247-264, 284-286, 397-401``):
``EnsembleSampler(nwalkers, ndim, log_prob_fn, pool=None)``, ``sample(p0, iterations, progress)``,
``iteration``, ``get_autocorr_time(tol=0)``, ``get_chain / get_log_prob(discard, thin, flat)``.

Algorithm: Goodman & Weare (2010) stretch move, red/blue split as in emcee's ``RedBlueMove`` with
``nsplits=2`` and ``a=2``: each half of the ensemble is updated against the other half,
and the whole half is handed to the log-probability function in ONE call -- with
``vectorize=True`` a ``(B, ndim)`` array goes straight to the batched GPU kernel (B = nwalkers / 2).
emcee is not importable in this environment; this is a restatement of its published algorithm, not
a copy, and the random stream differs (NumPy ``Generator`` instead of emcee's ``RandomState``).
"""
import ctypes

import numpy as np

from . import _native

__all__ = ["EnsembleSampler", "DeviceEnsembleSampler", "integrated_time", "AutocorrError"]


class AutocorrError(Exception):
    pass


def _next_pow_two(n):
    i = 1
    while i < n:
        i <<= 1
    return i


def _acf_1d(x):
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    return acf / acf[0]


def _mean_acf(x):
    """Walker-averaged normalised autocorrelation function of every dimension, ``(steps, ndim)``: emcee's ``function_1d``
    for all walkers and dimensions (host, NumPy)."""
    n_t = x.shape[0]
    n = _next_pow_two(n_t)
    y = np.ascontiguousarray(np.moveaxis(x, 0, -1))                  # (walkers, ndim, steps): transforms along contiguous rows
    y = y - y.mean(axis=-1, keepdims=True)
    f = np.fft.rfft(y, n=2 * n, axis=-1)
    acf = np.fft.irfft(f.real ** 2 + f.imag ** 2, n=2 * n, axis=-1)[..., :n_t]
    acf /= acf[..., :1]
    return acf.mean(axis=0).T


def _mean_acf_device(x):
    """The same on the GPU (``mtg_autocorr_host``: batched cuFFT): what ``DeviceEnsembleSampler`` uses -- at ~25 000 sampler
    steps/s the host FFTs of the whole chain at every convergence check were the bottleneck of ``derive_posteriors``."""
    n_t, n_w, n_d = x.shape
    rows = np.ascontiguousarray(np.moveaxis(x, 0, -1).reshape(n_w * n_d, n_t), dtype=np.float64)
    out = np.empty_like(rows)
    _native.check(_native.lib().mtg_autocorr_host(n_w * n_d, n_t, _native.ptr(rows), _native.ptr(out)))
    return out.reshape(n_w, n_d, n_t).mean(axis=0).T


def integrated_time(x, c=5, tol=50, quiet=False, acf=None):
    """Integrated autocorrelation time per dimension with Sokal's automatic window (emcee
    ``autocorr.integrated_time``): ``x`` is ``(steps, walkers, ndim)``.  ``acf``: function returning the walker-averaged
    autocorrelation functions ``(steps, ndim)`` (default: NumPy on the host)."""
    x = np.atleast_1d(x)
    if x.ndim == 1:
        x = x[:, None, None]
    if x.ndim == 2:
        x = x[:, :, None]
    n_t, n_w, n_d = x.shape
    tau = np.empty(n_d)
    windows = np.empty(n_d, dtype=int)
    mean_acf = (acf or _mean_acf)(np.asarray(x, dtype=np.float64))
    for d in range(n_d):
        taus = 2.0 * np.cumsum(mean_acf[:, d]) - 1.0
        m = np.arange(len(taus)) < c * taus
        windows[d] = np.argmin(m) if np.any(~m) else len(taus) - 1
        tau[d] = taus[windows[d]]
    if tol > 0 and np.any(tol * tau > n_t) and not quiet:
        raise AutocorrError("The chain is shorter than %d times the integrated autocorrelation time" % tol)
    return tau


class EnsembleSampler(object):
    def __init__(self, nwalkers, ndim, log_prob_fn, pool=None, moves=None, args=None, kwargs=None,
                 vectorize=False, a=2.0, seed=None):
        if nwalkers < 2 * ndim:
            raise ValueError("the number of walkers must be at least twice the dimension (got %d for ndim=%d)" % (nwalkers, ndim))
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)
        self.log_prob_fn = log_prob_fn
        self.pool = pool
        self.vectorize = vectorize
        self.args = args or []
        self.kwargs = kwargs or {}
        self.a = float(a)
        self._rng = np.random.default_rng(seed if seed is not None else np.random.randint(0, 2 ** 31 - 1))
        self.reset()

    def reset(self):
        self.iteration = 0
        self._chain = np.empty((0, self.nwalkers, self.ndim))
        self._log_prob = np.empty((0, self.nwalkers))
        self._accepted = np.zeros(self.nwalkers)

    # -- log probability of a batch of positions -------------------------------------------------------
    def compute_log_prob(self, coords):
        p = np.asarray(coords, dtype=np.float64)
        if np.any(~np.isfinite(p)):
            raise ValueError("At least one parameter value was infinite or NaN")
        if self.vectorize:
            lp = np.asarray(self.log_prob_fn(p, *self.args, **self.kwargs), dtype=np.float64)
        else:
            if self.pool is not None:
                vals = self.pool.map(self.log_prob_fn, list(p))
            else:
                vals = [self.log_prob_fn(q, *self.args, **self.kwargs) for q in p]
            lp = np.array([float(v) for v in vals], dtype=np.float64)
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return lp

    def _grow(self, n):
        self._chain = np.concatenate([self._chain, np.empty((n, self.nwalkers, self.ndim))], axis=0)
        self._log_prob = np.concatenate([self._log_prob, np.empty((n, self.nwalkers))], axis=0)

    def sample(self, initial_state, iterations=1, progress=False, store=True):
        p = np.array(initial_state, dtype=np.float64)
        if p.shape != (self.nwalkers, self.ndim):
            raise ValueError("incompatible input dimensions %r" % (p.shape,))
        if self.nwalkers > 1 and np.linalg.cond(np.atleast_2d(np.cov(p, rowvar=False))) > 1e8:
            raise ValueError("Initial state has a large condition number. Make sure that your walkers are linearly independent "
                             "for the best performance")
        lp = self.compute_log_prob(p)
        if store:
            self._grow(int(iterations))
        nw, nd = self.nwalkers, self.ndim
        for _ in range(int(iterations)):
            perm = self._rng.permutation(nw)
            halves = [perm[: nw // 2], perm[nw // 2:]]
            for h in range(2):
                S, C = halves[h], halves[1 - h]
                ns = len(S)
                zz = ((self.a - 1.0) * self._rng.random(ns) + 1.0) ** 2 / self.a
                partner = p[C[self._rng.integers(len(C), size=ns)]]
                q = partner - (partner - p[S]) * zz[:, None]
                lq = self.compute_log_prob(q)
                with np.errstate(invalid="ignore"):
                    lnpdiff = (nd - 1.0) * np.log(zz) + lq - lp[S]
                accept = np.log(self._rng.random(ns)) < lnpdiff
                idx = S[accept]
                p[idx] = q[accept]
                lp[idx] = lq[accept]
                self._accepted[idx] += 1
            if store:
                self._chain[self.iteration] = p
                self._log_prob[self.iteration] = lp
            self.iteration += 1
            yield p, lp

    def run_mcmc(self, initial_state, nsteps, **kwargs):
        out = None
        for out in self.sample(initial_state, iterations=nsteps, **kwargs):
            pass
        return out

    # -- results ---------------------------------------------------------------------------------------------
    def _get(self, arr, discard, thin, flat):
        v = arr[: self.iteration][discard + thin - 1:: thin]
        if flat:
            return v.reshape((-1,) + v.shape[2:])
        return v

    def get_chain(self, discard=0, thin=1, flat=False):
        return self._get(self._chain, int(discard), int(thin), flat)

    def get_log_prob(self, discard=0, thin=1, flat=False):
        return self._get(self._log_prob, int(discard), int(thin), flat)

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        return thin * integrated_time(self.get_chain(discard=discard, thin=thin), **kwargs)

    @property
    def acceptance_fraction(self):
        return self._accepted / max(self.iteration, 1)


class DeviceEnsembleSampler(EnsembleSampler):
    """The same sampler with the whole loop on the GPU: proposals, the celerite log-probability, acceptance and chain
    storage run inside one kernel launch per block of steps, for ``E`` independent ensembles in lock step.  Two kernels
    share the algorithm and the random streams: K8 (``csrc/mtg_ensemble.cu``, one thread per proposal: thousands of
    ensembles) and K10 (``csrc/mtg_scan.cu``, one CTA per proposal with the parallel-in-time likelihood and a barrier
    across CTAs per half step: the one-lightcurve case of ``derive_posteriors``); ``kernel="auto"`` lets the library
    choose (K10 when all ``E * ceil(W/2)`` CTAs fit the device and the model has <= 4 state dimensions).

    ``gp``: a computed ``mind_the_gaps_b200.celerite.GP`` (its kernel / mean / bounds define the log-probability,
    i.e. ``GPModelling._log_probability``); ``y``: ``(N,)`` one lightcurve shared by every ensemble, or ``(E, N)`` one
    per ensemble with ``yerr`` ``(E, N)`` (the simulated lightcurves of the LRT).  With ``E == 1`` the emcee-shaped
    results (``get_chain`` -> ``(steps, walkers, ndim)``) are unchanged; with ``E > 1`` they gain a leading
    ensemble axis.  ``block``: steps per kernel launch of ``sample`` (the generator yields once per block, with
    ``iteration`` advanced by the block; ``GPModelling.derive_posteriors`` sets it to its convergence interval)."""

    def __init__(self, nwalkers, gp, y, yerr=None, ensembles=1, a=2.0, seed=None, block=100, frozen_mean=None, kernel="auto"):
        if not gp.device_prior_supported():
            raise NotImplementedError("a kernel term overrides log_prior in a way the device sampler does not restate; "
                                      "use the host sampler (derive_posteriors(device_sampler=False))")
        model = gp.device_model()
        if kernel not in ("auto", "k8", "scan"):
            raise ValueError("kernel must be 'auto', 'k8' or 'scan'")
        self._kernel_flag = {"auto": 0, "k8": _native.FLAG_NO_SCAN, "scan": _native.FLAG_SCAN}[kernel]
        ndim = int(model.n_free)
        if nwalkers < 2 * ndim:
            raise ValueError("the number of walkers must be at least twice the dimension (got %d for ndim=%d)" % (nwalkers, ndim))
        _native.require_device()
        self.nwalkers, self.ndim = int(nwalkers), ndim
        self.a = float(a)
        self.block = max(1, int(block))
        self._model = model
        self._t = np.ascontiguousarray(gp._t, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        if y.ndim == 1:
            self._y = y[None, :]
            self._diag = np.ascontiguousarray(gp._diag, dtype=np.float64)[None, :]
            self.E = int(ensembles)
        else:
            if yerr is None:
                raise ValueError("per-ensemble lightcurves need their own yerr (E, N)")
            self._y = y
            self._diag = np.ascontiguousarray((np.asarray(yerr, dtype=np.float64) + 0.0) ** 2)
            self.E = y.shape[0]
            if ensembles not in (1, self.E):
                raise ValueError("ensembles must equal the number of lightcurves")
        if self._y.shape[1] != len(self._t) or self._diag.shape != self._y.shape:
            raise ValueError("dimension mismatch")
        self._means = None
        self._flags = 0
        if frozen_mean is not None:
            self._means = np.ascontiguousarray(np.broadcast_to(np.asarray(frozen_mean, dtype=np.float64), (self._y.shape[0],)))
            self._flags = _native.FLAG_DS_MEAN
        self._seed = int(seed if seed is not None else np.random.randint(0, 2 ** 31 - 1)) * (2 ** 31) + \
            int(np.random.randint(0, 2 ** 31 - 1))
        self.reset()

    def reset(self):
        self.iteration = 0
        shape = (self.E,) if self.E > 1 else ()
        self._chain = np.empty(shape + (0, self.nwalkers, self.ndim))
        self._log_prob = np.empty(shape + (0, self.nwalkers))
        self._accepted_i = np.zeros((self.E, self.nwalkers), dtype=np.int32)

    def _run(self, pos, lp, init, nsteps, store, chain=None, clp=None):
        E, W, nd = self.E, self.nwalkers, self.ndim
        if store and chain is None:
            chain = np.empty((E, nsteps, W, nd)); clp = np.empty((E, nsteps, W))
        rc = _native.lib().mtg_ensemble_host(ctypes.byref(self._model), len(self._t), self._y.shape[0], E, W, _native.ptr(self._t),
                                             _native.ptr(self._y), _native.ptr(self._diag), _native.ptr(self._means),
                                             _native.ptr(pos), _native.ptr(lp), int(init), int(nsteps), self.a, self._seed,
                                             int(self.iteration), _native.ptr(chain), _native.ptr(clp),
                                             _native.ptr(self._accepted_i), self._flags | self._kernel_flag)
        _native.check(rc)
        return chain, clp

    def sample(self, initial_state, iterations=1, progress=False, store=True):
        E, W, nd = self.E, self.nwalkers, self.ndim
        p = np.array(initial_state, dtype=np.float64)
        if p.shape == (W, nd):
            p = np.repeat(p[None], E, axis=0) if E > 1 else p[None]
        if p.shape != (E, W, nd):
            raise ValueError("incompatible input dimensions %r" % (p.shape,))
        if np.any(~np.isfinite(p)):
            raise ValueError("At least one parameter value was infinite or NaN")
        for e in range(E if E <= 8 else 1):
            if W > 1 and np.linalg.cond(np.atleast_2d(np.cov(p[e], rowvar=False))) > 1e8:
                raise ValueError("Initial state has a large condition number. Make sure that your walkers are linearly independent "
                                 "for the best performance")
        p = np.ascontiguousarray(p)
        lp = np.zeros((E, W))
        init = True
        done = 0
        iterations = int(iterations)
        if store and E == 1:
            # one ensemble: grow the emcee-shaped arrays once and let every launch write its block in place (K10 runs
            # ~25 000 steps/s: re-concatenating the chain per block would cost more than the sampling)
            have = self._chain.shape[0]
            if have < self.iteration + iterations:
                extra = self.iteration + iterations - have
                self._chain = np.concatenate([self._chain, np.empty((extra, W, nd))], axis=0)
                self._log_prob = np.concatenate([self._log_prob, np.empty((extra, W))], axis=0)
        while done < iterations:
            n = min(self.block, iterations - done)
            if store and E == 1:
                i0 = self.iteration
                self._run(p, lp, init, n, True, self._chain[None, i0:i0 + n], self._log_prob[None, i0:i0 + n])
            else:
                chain, clp = self._run(p, lp, init, n, store)
                if store:
                    self._chain = np.concatenate([self._chain, chain], axis=1)
                    self._log_prob = np.concatenate([self._log_prob, clp], axis=1)
            init = False
            self.iteration += n
            done += n
            yield (p if E > 1 else p[0]), (lp if E > 1 else lp[0])

    def _get(self, arr, discard, thin, flat):
        if self.E == 1:
            return super()._get(arr, discard, thin, flat)
        v = arr[:, : self.iteration][:, discard + thin - 1:: thin]
        if flat:
            return v.reshape((v.shape[0], -1) + v.shape[3:])
        return v

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        ch = self.get_chain(discard=discard, thin=thin)
        # short chains: the host FFTs cost less than the trip to the device
        if ch.shape[-3] * ch.shape[-2] * ch.shape[-1] >= 60000:
            kwargs.setdefault("acf", _mean_acf_device)
        if self.E == 1:
            return thin * integrated_time(ch, **kwargs)
        return np.array([thin * integrated_time(ch[e], **kwargs) for e in range(self.E)])

    @property
    def acceptance_fraction(self):
        f = self._accepted_i / max(self.iteration, 1)
        return f if self.E > 1 else f[0]
