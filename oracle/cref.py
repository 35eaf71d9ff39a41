"""Oracle: ctypes loader for ``celerite_ref.c`` (OpenMP CPU restatement).  TEST INFRASTRUCTURE ONLY.

Used by tests as a second checker and by ``bench.py`` as the timed CPU baseline
(``cpu_baseline.kind = "port"``: celerite itself is not importable, see ``oracle/__init__.py``).
"""
import ctypes
import os
import subprocess

import numpy as np

from . import terms as _terms

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "celerite_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        try:
            build()
            _lib = ctypes.CDLL(_SO)
        except (OSError, subprocess.CalledProcessError):
            build(force=True)
            _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        _lib.oracle_loglike_batch.argtypes = [ctypes.c_int, dp, dp, dp, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              dp, dp, dp, dp, dp, dp, dp, dp, dp, ip, ctypes.c_int]
        _lib.oracle_loglike_batch.restype = ctypes.c_int
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(lib().oracle_max_threads())


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def loglike_coeffs(t, y, diag, ar, cr, ac, bc, cc, dc, jitter=None, mean=None, group=None, nthreads=0):
    """``y, diag``: ``(S, N)`` or ``(N,)``; coefficient arrays ``(B, Jr)`` / ``(B, Jc)``."""
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    t = c(t); y = c(np.atleast_2d(y)); diag = c(np.atleast_2d(diag))
    ar, cr, ac, bc, cc, dc = [c(np.atleast_2d(x)) for x in (ar, cr, ac, bc, cc, dc)]
    B = max(ar.shape[0], ac.shape[0])
    S, N = y.shape
    if group is None:
        group = B // S
    jitter = c(np.broadcast_to(0.0 if jitter is None else jitter, (B,)))
    mean_ = None if mean is None else c(np.broadcast_to(mean, (B,)))
    out = np.empty(B); status = np.empty(B, dtype=np.int32)
    rc = lib().oracle_loglike_batch(N, _p(t), _p(y), _p(diag), S, group, B, ar.shape[1], ac.shape[1],
                                    _p(ar), _p(cr), _p(ac), _p(bc), _p(cc), _p(dc), _p(jitter), _p(mean_),
                                    _p(out), status.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), nthreads)
    if rc != 0:
        raise RuntimeError("oracle_loglike_batch failed rc=%d" % rc)
    return out, status


def loglike_params(spec, params, t, y, dy, mean=None, group=None, nthreads=0):
    """Same contract as ``oracle.gp.loglike_params`` (frozen mean), evaluated by the C code."""
    p = np.atleast_2d(np.asarray(params, dtype=float))
    y = np.atleast_2d(np.asarray(y, dtype=float))
    diag = (np.atleast_2d(np.asarray(dy, dtype=float)) + 1e-12) ** 2
    B = p.shape[0]; S = y.shape[0]
    if group is None:
        group = B // S
    co = _terms.coefficients(spec, p)
    if mean is None:
        mean = np.repeat(y.mean(axis=1), group)[:B]
    ll, _ = loglike_coeffs(t, y, diag, *co[:6], jitter=co[6], mean=mean, group=group, nthreads=nthreads)
    return ll
