"""GP stand-in: dense Cholesky of K_ij = k(|t_i - t_j|) + delta_ij (yerr_i^2 + jitter)."""
import numpy as np
from scipy import linalg
from .modeling import ModelSet, ConstantModel


class LinAlgError(Exception):
    pass


class GP(ModelSet):
    def __init__(self, kernel, mean=0.0, fit_mean=False, log_white_noise=None, fit_white_noise=False):
        self._t = None
        self._yerr = None
        try:
            float(mean)
        except TypeError:
            pass
        else:
            mean = ConstantModel(float(mean))
        if not fit_mean:
            mean.freeze_all_parameters()
        super(GP, self).__init__([("kernel", kernel), ("mean", mean)])

    @property
    def mean(self):
        return self.models["mean"]

    @property
    def kernel(self):
        return self.models["kernel"]

    def compute(self, t, yerr=1.123e-12, check_sorted=True):
        t = np.atleast_1d(np.asarray(t, dtype=float))
        if check_sorted and np.any(np.diff(t) < 0.0):
            raise ValueError("the input coordinates must be sorted")
        if t.ndim > 1:
            raise ValueError("dimension mismatch")
        self._t = t
        self._yerr = np.empty_like(t)
        self._yerr[:] = yerr

    def _factor(self):
        tau = np.abs(self._t[:, None] - self._t[None, :])
        K = self.kernel.get_value(tau)
        K[np.diag_indices_from(K)] += self._yerr ** 2 + self.kernel.jitter
        try:
            return linalg.cho_factor(K, lower=True)
        except linalg.LinAlgError:
            raise LinAlgError("failed to factorize or solve matrix")

    def log_likelihood(self, y, _const=np.log(2.0 * np.pi), quiet=False):
        y = np.asarray(y, dtype=float)
        if y.shape[0] != len(self._t):
            raise ValueError("dimension mismatch")
        resid = y - self.mean.get_value(self._t)
        try:
            cf = self._factor()
        except LinAlgError:
            if quiet:
                return -np.inf
            raise
        logdet = 2.0 * np.sum(np.log(np.diag(cf[0])))
        if not np.isfinite(logdet):
            return -np.inf
        loglike = -0.5 * (np.dot(resid, linalg.cho_solve(cf, resid)) + logdet + len(y) * _const)
        if not np.isfinite(loglike):
            return -np.inf
        return loglike

    def predict(self, y, t=None, return_cov=True, return_var=False):
        y = np.asarray(y, dtype=float)
        resid = y - self.mean.get_value(self._t)
        cf = self._factor()
        alpha = linalg.cho_solve(cf, resid)
        if t is None:
            xs = self._t
        else:
            xs = np.asarray(t, dtype=float)
        Kxs = self.kernel.get_value(np.abs(xs[:, None] - self._t[None, :]))
        mu = Kxs @ alpha + self.mean.get_value(xs)
        if not (return_var or return_cov):
            return mu
        KinvKxsT = linalg.cho_solve(cf, Kxs.T)
        if return_var:
            var = self.kernel.get_value(0.0) - np.sum(Kxs.T * KinvKxsT, axis=0)
            return mu, var
        cov = self.kernel.get_value(np.abs(xs[:, None] - xs[None, :])) - Kxs @ KinvKxsT
        return mu, cov
