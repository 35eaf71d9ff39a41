"""``celerite.GP`` as used by ``GPModelling`` -- evaluated by libmtg_b200 on a B200.

Surface mirrored (call sites in ``/root/reference/mind_the_gaps/gpmodelling.py``):
``GP(kernel, mean=Model, fit_mean=bool)`` (:51), ``compute(t, yerr)`` (:54),
``get_parameter_vector()`` (:55), ``set_parameter_vector(p)`` (:147, :168, :533), ``log_prior()``
(:149), ``log_likelihood(y)`` (:152, :169), ``get_parameter_bounds()`` (:193, :239),
``get_parameter_names()`` (:454), ``predict(y, return_var=True, return_cov=False)`` (:366),
``.kernel.jitter`` (:368), ``.kernel.get_psd`` (:509, :535).

Added for the device: ``log_likelihood_batch`` / ``log_probability_batch`` /
``grad_log_likelihood_batch`` evaluate B parameter vectors in one kernel launch (the seam emcee's
``vectorize=True`` and the batched optimiser use).
"""
import ctypes

import numpy as np

from .. import _native
from .modeling import ConstantModel, Model, ModelSet
from .terms import Term

__all__ = ["GP", "LinAlgError"]


class LinAlgError(np.linalg.LinAlgError):
    """Raised where celerite raises ``celerite.solver.LinAlgError`` (matrix not positive definite)."""


_MEAN_KINDS = {"ConstantModel": (_native.MEAN_CONSTANT, ("value",)),
               "LinearModel": (_native.MEAN_LINEAR, ("slope", "intercept")),
               "GaussianModel": (_native.MEAN_GAUSSIAN, ("mean", "sigma", "amplitude", "constant"))}


class GP(ModelSet):
    def __init__(self, kernel, mean=0.0, fit_mean=False):
        if not isinstance(kernel, Term):
            raise TypeError("kernel must be a mind_the_gaps_b200.celerite.terms.Term")
        if not isinstance(mean, Model):
            mean = ConstantModel(float(mean))
        if not fit_mean:
            mean.freeze_all_parameters()
        super(GP, self).__init__([("kernel", kernel), ("mean", mean)])
        self._t = None
        self._yerr = None
        self._diag = None

    @property
    def kernel(self):
        return self.models["kernel"]

    @property
    def mean(self):
        return self.models["mean"]

    @property
    def computed(self):
        return self._t is not None

    # -- model compilation ----------------------------------------------------------------------
    def device_model(self):
        """Compile kernel + mean + frozen mask + bounds into an ``mtg_model_t``."""
        m = _native.MtgModel()
        terms = self.kernel.terms
        if len(terms) > _native.MTG_MAX_TERMS:
            raise ValueError("at most %d kernel terms are supported" % _native.MTG_MAX_TERMS)
        m.n_terms = len(terms)
        off = 0
        for k, term in enumerate(terms):
            kind = term._mtg_kind
            if kind is None:
                raise NotImplementedError(
                    "%s has no device coefficient map (a user-defined Term): its likelihood runs on the GPU through the "
                    "raw-coefficient entry -- GP.log_likelihood(_batch) / log_probability_batch, GPModelling.fit and the host "
                    "sampler -- but the device-resident optimiser, sampler, predict and LRT need one of the built-in terms"
                    % type(term).__name__)
            m.kind[k] = int(kind)
            m.extra[k] = float(term._mtg_extra)
            for i in range(term.full_size):
                m.pidx[k][i] = off + i
            off += term.full_size
        mean = self.mean
        mk = None
        for cls in type(mean).__mro__:
            if cls.__name__ in _MEAN_KINDS:
                mk = _MEAN_KINDS[cls.__name__]
                break
        if mk is None:
            raise NotImplementedError("mean model %s is not supported on the GPU path" % type(mean).__name__)
        m.mean_kind = mk[0]
        names = mean.get_parameter_names(include_frozen=True)
        for i, pname in enumerate(mk[1]):
            # the reference builds GaussianModel with 3 values for 4 names (gpmodelling.py:118 vs
            # mean_models.py:7); tolerate a missing trailing "constant" by pointing at a zero default
            m.mean_pidx[i] = off + names.index(pname) if pname in names else -1
        n_full = off + mean.full_size
        if n_full > _native.MTG_MAX_PARAMS:
            raise ValueError("too many parameters (%d)" % n_full)
        full = self.get_parameter_vector(include_frozen=True)
        for i in range(n_full):
            m.full_default[i] = float(full[i])
        for i in range(4):
            if m.mean_pidx[i] < 0:
                m.mean_pidx[i] = n_full
                m.full_default[n_full] = 0.0
                n_full += 1
        m.n_full = n_full
        mask = self.unfrozen_mask
        bounds = self.get_parameter_bounds(include_frozen=True)
        kfree = 0
        for i in range(len(mask)):
            if mask[i]:
                lo, hi = bounds[i]
                m.free_idx[kfree] = i
                m.lo[kfree] = -np.inf if lo is None else float(lo)
                m.hi[kfree] = np.inf if hi is None else float(hi)
                kfree += 1
        m.n_free = kfree
        return m

    # -- user-defined terms: coefficient map on the host, everything else on the device --------------------------
    def needs_host_coefficients(self):
        """True when a kernel term has no device coefficient map: a user subclass of ``Term`` with its own
        ``get_real_coefficients`` / ``get_complex_coefficients`` (the reference's kernels are written this way,
        models/celerite_models.py:7-90).  Its coefficients are then evaluated in Python once per parameter set and the
        likelihood runs on the device from them (``MTG_TERM_RAW_*``)."""
        return any(t._mtg_kind is None for t in self.kernel.terms)

    def _raw_model_and_rows(self, params):
        """``(mtg_model_t of raw terms, rows (B, n_raw))`` for the parameter vectors ``params (B, ndim)``: each row holds the
        celerite coefficients of the kernel at that parameter vector -- (a, c) per real term, (a, b, c, d) per complex
        term, the jitter -- followed by the free mean parameters.  Degenerate real terms (a = c = 0, what the reference's
        ``Lorentzian`` returns) are dropped: they cannot influence the likelihood."""
        keep = self.get_parameter_vector()
        nk = self.kernel.vector_size
        rows, layout = [], None
        try:
            for row in params:
                self.set_parameter_vector(row)
                ar, cr, ac, bc, cc, dc = self.kernel.get_all_coefficients()
                live = ~((ar == 0.0) & (cr == 0.0))
                lay = (tuple(live), len(ac))
                if layout is None:
                    layout = lay
                elif lay != layout:
                    raise ValueError("the number of real / complex celerite terms changes with the parameters: not supported "
                                     "for user-defined terms")
                rows.append(np.concatenate([np.column_stack([ar[live], cr[live]]).ravel(), np.column_stack([ac, bc, cc, dc]).ravel(),
                                            [self.kernel.jitter], row[nk:]]))
        finally:
            self.set_parameter_vector(keep)
        n_r, n_c = int(np.sum(layout[0])), layout[1]
        if n_r + n_c + 1 > _native.MTG_MAX_TERMS:
            raise ValueError("at most %d celerite terms are supported" % (_native.MTG_MAX_TERMS - 1))
        m = _native.MtgModel()
        m.n_terms = n_r + n_c + 1
        off = 0
        for k in range(n_r):
            m.kind[k] = 10; m.pidx[k][0] = off; m.pidx[k][1] = off + 1; off += 2                      # MTG_TERM_RAW_REAL
        for k in range(n_r, n_r + n_c):
            m.kind[k] = 11                                                                               # MTG_TERM_RAW_COMPLEX
            for i in range(4):
                m.pidx[k][i] = off + i
            off += 4
        m.kind[n_r + n_c] = 12; m.pidx[n_r + n_c][0] = off; off += 1                                     # MTG_TERM_RAW_JITTER
        n_coef = off
        # mean model: the free mean parameters follow the coefficients in the rows, the frozen ones keep their values
        mean = self.mean
        mk = None
        for cls in type(mean).__mro__:
            if cls.__name__ in _MEAN_KINDS:
                mk = _MEAN_KINDS[cls.__name__]
                break
        if mk is None:
            raise NotImplementedError("mean model %s is not supported on the GPU path" % type(mean).__name__)
        m.mean_kind = mk[0]
        names = list(mean.get_parameter_names(include_frozen=True))
        mfull = mean.get_parameter_vector(include_frozen=True)
        mmask = np.atleast_1d(mean.unfrozen_mask)
        n_full = n_coef + len(names)
        if n_full + 1 > _native.MTG_MAX_PARAMS:
            raise ValueError("too many parameters (%d)" % n_full)
        for i in range(n_coef):
            m.full_default[i] = 0.0
        for i, v in enumerate(mfull):
            m.full_default[n_coef + i] = float(v)
        for i, pname in enumerate(mk[1]):
            if pname in names:
                m.mean_pidx[i] = n_coef + names.index(pname)
            else:                                       # GaussianModel built with 3 of its 4 names (see device_model)
                m.mean_pidx[i] = n_full; m.full_default[n_full] = 0.0; n_full += 1
        m.n_full = n_full
        kfree = 0
        for i in range(n_coef):
            m.free_idx[kfree] = i; m.lo[kfree] = -np.inf; m.hi[kfree] = np.inf; kfree += 1
        for i in range(len(names)):
            if mmask[i]:
                m.free_idx[kfree] = n_coef + i; m.lo[kfree] = -np.inf; m.hi[kfree] = np.inf; kfree += 1
        m.n_free = kfree
        return m, np.ascontiguousarray(rows, dtype=np.float64)

    # -- data -------------------------------------------------------------------------------------
    def compute(self, t, yerr=1.123e-12, check_sorted=True):
        t = np.atleast_1d(np.asarray(t, dtype=np.float64))
        if t.ndim != 1:
            raise ValueError("dimension mismatch")
        if check_sorted and np.any(np.diff(t) < 0.0):
            raise ValueError("the input coordinates must be sorted")
        yerr = np.asarray(yerr, dtype=np.float64)
        self._t = np.ascontiguousarray(t)
        self._yerr = np.ascontiguousarray(np.broadcast_to(yerr, t.shape))
        self._diag = np.ascontiguousarray(self._yerr ** 2)
        self.__dict__["_dirty"] = True

    def _check_y(self, y):
        if self._t is None:
            raise RuntimeError("you must call 'compute' first")
        y = np.ascontiguousarray(y, dtype=np.float64)
        if y.shape[-1] != len(self._t):
            raise ValueError("dimension mismatch")
        return y

    # -- batched device evaluation ------------------------------------------------------------------
    def _batch_params(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.ndim == 1:
            p = p[None, :]
        if p.ndim != 2 or p.shape[1] != self.vector_size:
            raise ValueError("expected parameters of shape (B, %d)" % self.vector_size)
        return p

    _KERNEL_FLAGS = {"auto": 0, "k1": _native.FLAG_NO_SCAN, "scan": _native.FLAG_SCAN}

    def log_likelihood_batch(self, y, params, apply_prior=False, return_status=False, out=None, kernel="auto"):
        """log-likelihood of ``y`` under B parameter vectors ``params[B, ndim]`` in one launch.

        ``kernel``: ``"auto"`` (the library picks: the parallel-in-time scan kernel K9 for small batches of narrow
        kernels, the one-thread-per-series kernel K1 otherwise), ``"k1"`` or ``"scan"`` to force one."""
        y = self._check_y(y)
        p = self._batch_params(params)
        B = p.shape[0]
        _native.require_device()
        if self.needs_host_coefficients():
            if apply_prior:
                raise ValueError("a user-defined term's prior is evaluated on the host: use log_probability_batch")
            if B == 0:
                return (np.empty(0), np.empty(0, dtype=np.int32)) if return_status else np.empty(0)
            model, p = self._raw_model_and_rows(p)
        else:
            model = self.device_model()
        ll = out if out is not None else np.empty(B, dtype=np.float64)
        # the status array is only copied back when asked for: a device->host copy into pageable memory blocks the
        # calling thread, which would serialise the chunk pipeline of large host calls (csrc/mtg_capi.cu host_eval)
        status = np.empty(B, dtype=np.int32) if return_status else None
        flags = (_native.FLAG_APPLY_PRIOR if apply_prior else 0) | self._KERNEL_FLAGS[kernel]
        rc = _native.lib().mtg_loglike_host(ctypes.byref(model), len(self._t), 1, max(B, 1), B,
                                            _native.ptr(self._t), _native.ptr(y), _native.ptr(self._diag), None,
                                            _native.ptr(p), _native.ptr(ll), _native.ptr(status) if return_status else None, flags)
        _native.check(rc)
        return (ll, status) if return_status else ll

    _GRAD_KERNEL_FLAGS = {"auto": 0, "k2": _native.FLAG_NO_SCAN, "scan": _native.FLAG_SCAN}

    def grad_log_likelihood_batch(self, y, params, apply_prior=False, return_status=False, kernel="auto"):
        """``(loglike[B], grad[B, ndim])`` by reverse-mode differentiation on the device.

        ``kernel``: ``"auto"`` (the library picks: the parallel-in-time scan gradient K11 -- one CTA per series -- for small
        batches of narrow kernels, the one-thread-per-series kernel K2 otherwise), ``"k2"`` or ``"scan"`` to force one."""
        y = self._check_y(y)
        p = self._batch_params(params)
        B = p.shape[0]
        _native.require_device()
        model = self.device_model()          # raises for user-defined terms: their map to coefficients is arbitrary Python
        ll = np.empty(B, dtype=np.float64)
        grad = np.empty((B, p.shape[1]), dtype=np.float64)
        status = np.empty(B, dtype=np.int32)
        flags = (_native.FLAG_APPLY_PRIOR if apply_prior else 0) | self._GRAD_KERNEL_FLAGS[kernel]
        rc = _native.lib().mtg_loglike_grad_host(ctypes.byref(model), len(self._t), 1, max(B, 1), B,
                                                 _native.ptr(self._t), _native.ptr(y), _native.ptr(self._diag), None,
                                                 _native.ptr(p), _native.ptr(ll), _native.ptr(grad),
                                                 _native.ptr(status), flags)
        _native.check(rc)
        return (ll, grad, status) if return_status else (ll, grad)

    def device_prior_supported(self):
        """True when the device restates this model's ``log_prior`` exactly: the box of every parameter plus the two
        overrides the kernels know (``ComplexTerm`` with ``log_b``: ``a c >= b d``; ``BendingPowerlaw``:
        ``log_S0 >= log_Q``, celerite_models.py:85-90).  A user subclass that overrides ``log_prior`` -- even one that
        inherits a device coefficient map -- makes this False, and the prior is then evaluated on the host."""
        from ..models.celerite_models import BendingPowerlaw
        from .terms import ComplexTerm, TermSum
        if self.needs_host_coefficients():
            return False
        known = {Model.log_prior, ModelSet.log_prior, TermSum.log_prior, ComplexTerm.log_prior, BendingPowerlaw.log_prior}
        parts = list(self.kernel.terms) + [self.kernel, self.mean]
        return all(type(m).log_prior in known for m in parts)

    def log_probability_batch(self, y, params, kernel="auto"):
        """``log_prior + log_likelihood`` for B parameter vectors: ``GPModelling._log_probability``
        (gpmodelling.py:127-152) vectorised; ``-inf`` outside the prior or when not positive definite."""
        if self.device_prior_supported():
            return self.log_likelihood_batch(y, params, apply_prior=True, kernel=kernel)
        # a log_prior the device does not know: the prior row by row on the host (arbitrary Python, as the reference
        # runs it), the likelihood of the rows inside it in one launch
        p = self._batch_params(params)
        keep = self.get_parameter_vector()
        lp = np.empty(p.shape[0])
        try:
            for i, row in enumerate(p):
                self.set_parameter_vector(row)
                lp[i] = self.log_prior()
        finally:
            self.set_parameter_vector(keep)
        out = np.full(p.shape[0], -np.inf)
        ok = np.isfinite(lp)
        if np.any(ok):
            out[ok] = lp[ok] + self.log_likelihood_batch(y, p[ok], apply_prior=False, kernel=kernel)
        return out

    # -- celerite single-point API ---------------------------------------------------------------------
    def log_likelihood(self, y, _const=None, quiet=False):
        ll, status = self.log_likelihood_batch(y, self.get_parameter_vector()[None, :], return_status=True)
        if status[0] == 1 and not quiet:
            raise LinAlgError("failed to factorize or solve matrix")
        if not np.isfinite(ll[0]):
            return -np.inf
        return float(ll[0])

    def grad_log_likelihood(self, y, quiet=False):
        ll, g, status = self.grad_log_likelihood_batch(y, self.get_parameter_vector()[None, :], return_status=True)
        if status[0] == 1 and not quiet:
            raise LinAlgError("failed to factorize or solve matrix")
        if not np.isfinite(ll[0]):
            return -np.inf, np.zeros(self.vector_size)
        return float(ll[0]), g[0]

    def predict(self, y, t=None, return_cov=True, return_var=False):
        from .predict import predict as _predict
        return _predict(self, y, t=t, return_cov=return_cov, return_var=return_var)
