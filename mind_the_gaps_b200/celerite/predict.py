"""``celerite.GP.predict`` at the training times, on the device (``mtg_predict_host``).

The reference only calls ``gp.predict(y, return_var=True, return_cov=False)``
(``/root/reference/mind_the_gaps/gpmodelling.py:366``).  Mean and variance follow from quantities the
reverse-mode kernel already forms: ``alpha = K^-1 r = -d loglike/d r`` and
``diag(K^-1) = alpha^2 - 2 d loglike/d diag``, so ``mu = mean + r - Sigma alpha`` and
``var = Sigma - Sigma^2 diag(K^-1)`` with ``Sigma = yerr^2 + jitter`` -- O(N J^2) instead of celerite's
dense O(N^2 J) variance.  Prediction at new times / full covariance is not on the hot path.
"""
import ctypes

import numpy as np

from .. import _native


def predict(gp, y, t=None, return_cov=True, return_var=False):
    if t is not None:
        raise NotImplementedError("prediction at new times is outside the mind_the_gaps hot path")
    y = gp._check_y(y)
    _native.require_device()
    model = gp.device_model()
    N = len(gp._t)
    mu = np.empty(N); var = np.empty(N)
    p = np.ascontiguousarray(gp.get_parameter_vector(), dtype=np.float64)
    rc = _native.lib().mtg_predict_host(ctypes.byref(model), N, _native.ptr(gp._t), _native.ptr(y), _native.ptr(gp._diag),
                                        _native.ptr(p) if len(p) else None, _native.ptr(mu), _native.ptr(var))
    _native.check(rc)
    if return_var:
        return mu, var
    if return_cov:
        raise NotImplementedError("the full predictive covariance is outside the mind_the_gaps hot path; use return_var=True")
    return mu
