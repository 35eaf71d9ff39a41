"""``GPModelling``: the Gaussian-process modelling interface of mind_the_gaps
(``/root/reference/mind_the_gaps/gpmodelling.py``) on top of the B200 kernels.

Same constructor, methods, keyword arguments, properties and error behaviour as the reference
class.  What changed underneath:

* ``self.gp`` is ``mind_the_gaps_b200.celerite.GP`` (K1/K2 kernels) instead of ``celerite.GP``;
* ``fit`` runs the device L-BFGS with analytic gradients (one launch of the persistent kernel for a single lightcurve;
  the same optimiser fits the thousands of simulated lightcurves of the LRT); ``optimizer="scipy"`` keeps the reference's
  scipy L-BFGS-B call with every finite-difference gradient request done as one batched launch;
* ``derive_posteriors`` runs emcee's stretch-move loop inside ONE kernel launch per convergence interval
  (``sampler.DeviceEnsembleSampler``: kernel K10 for the one-lightcurve case, K8 for thousands of ensembles;
  ``device_sampler=False`` keeps the host loop with one batched launch per half step) -- ``cores`` is accepted and
  ignored, there is no process pool;
* ``generate_from_posteriors`` simulates all ``nsims`` lightcurves in one batched device call
  (``cpus`` accepted and ignored).
"""
import ctypes
import warnings

import numpy as np
from scipy.optimize import OptimizeResult

from . import _native, fitting
from .celerite import GP
from .celerite.modeling import ConstantModel
from .lightcurves import GappyLightcurve
from .models import GaussianModel, LinearModel
from .sampler import DeviceEnsembleSampler, EnsembleSampler

__all__ = ["GPModelling"]


class GPModelling(object):
    """The interface for Gaussian Process (GP) modelling."""
    meanmodels = ["linear", "constant", "gaussian"]

    def __init__(self, lightcurve, kernel, mean_model=None):
        self._lightcurve = lightcurve
        meanmodel, fit_mean = self._build_mean_model(mean_model)
        self.gp = GP(kernel, mean=meanmodel, fit_mean=fit_mean)
        self.gp.compute(self._lightcurve.times, self._lightcurve.dy + 1e-12)      # gpmodelling.py:54
        self.initial_params = self.gp.get_parameter_vector()
        self._ndim = len(self.initial_params)
        self._autocorr = []
        self._loglikelihoods = None
        self._mcmc_samples = None
        self._sampler = None

    def _build_mean_model(self, meanmodel=None):
        lc = self._lightcurve
        maxy = np.max(lc.y)
        if meanmodel is None:
            return ConstantModel(lc.mean, bounds=[(np.min(lc.y), maxy)]), False
        if meanmodel.lower() not in GPModelling.meanmodels:
            raise ValueError("Input mean model %s not implemented! Only \n %s \n are available"
                             % (meanmodel, "\t".join(GPModelling.meanmodels)))
        if meanmodel.lower() == "constant":
            return ConstantModel(lc.mean, bounds=[(np.min(lc.y), maxy)]), True
        if meanmodel.lower() == "linear":
            return LinearModel(0, 1.5, bounds=[(-np.inf, np.inf), (-np.inf, np.inf)]), True
        # gaussian (gpmodelling.py:114-122).  The reference passes three values for GaussianModel's four
        # parameter names (a latent bug that raises inside celerite); the missing ``constant`` is started
        # at the lightcurve mean here.
        sigma_guess = lc.duration / 2
        amplitude_guess = (maxy - np.min(lc.y)) * np.sqrt(2 * np.pi) * sigma_guess
        mean_guess = lc.times[len(lc.times) // 2]
        amp_lo = maxy * np.sqrt(2 * np.pi) * lc.duration
        amp_hi = 50 * maxy * np.sqrt(2 * np.pi) * lc.duration
        amplitude_guess = min(max(amplitude_guess, amp_lo), amp_hi)
        return GaussianModel(mean_guess, sigma_guess, amplitude_guess, lc.mean,
                             bounds=[(lc.times[0], lc.times[-1]), (0, lc.duration), (amp_lo, amp_hi),
                                     (np.min(lc.y), maxy)]), True

    # -- likelihood ----------------------------------------------------------------------------------------
    def _log_probability(self, params):
        """log posterior at ``params`` (``(ndim,)``) or at a batch ``(B, ndim)`` -> ``(B,)``."""
        params = np.asarray(params, dtype=np.float64)
        if params.ndim == 2:
            return self.gp.log_probability_batch(self._lightcurve.y, params)
        self.gp.set_parameter_vector(params)
        lp = self.gp.log_prior()
        if not np.isfinite(lp):
            return -np.inf
        return lp + self.gp.log_likelihood(self._lightcurve.y)

    def _neg_log_like(self, params):
        self.gp.set_parameter_vector(params)
        return -self.gp.log_likelihood(self._lightcurve.y)

    def fit(self, initial_params=None, n_starts=1, optimizer="auto", **opts):
        """Maximum-likelihood fit within ``gp.get_parameter_bounds()``; returns an ``OptimizeResult`` like
        ``scipy.optimize.minimize``.

        ``optimizer="device"``: the device L-BFGS with analytic gradients that also serves the LRT -- for one lightcurve the
        whole optimisation is ONE launch of the persistent kernel (``csrc/mtg_lbfgs.cu``: parallel-in-time gradient K11,
        ~30 us per iteration); with ``n_starts > 1`` the best of the starts is returned.  ``optimizer="scipy"``: the
        reference's own call -- scipy's L-BFGS-B with 2-point finite differences (gpmodelling.py:192-193) -- every
        function-and-gradient request being one batched device launch (``fitting.scipy_lbfgsb``): same iterates as the
        reference.  ``"auto"`` (default): the device optimiser when the parallel-in-time kernels cover the model (<= 4 state
        dimensions, series fits an SM's shared memory) or ``n_starts > 1``, scipy otherwise."""
        if initial_params is None:
            initial_params = self.initial_params
        if optimizer == "auto":
            lib = _native.lib()
            scan_ok = (not self.gp.needs_host_coefficients()) and lib.mtg_device_count() > 0 and \
                bool(lib.mtg_scan_grad_supported(ctypes.byref(self.gp.device_model()), int(self._lightcurve.n)))
            if self.gp.needs_host_coefficients() and int(n_starts) > 1:
                raise NotImplementedError("multi-start fits run on the device optimiser, which needs built-in kernel terms")
            optimizer = "device" if (scan_ok or int(n_starts) > 1) else "scipy"
        if optimizer not in ("scipy", "device"):
            raise ValueError("optimizer must be 'auto', 'scipy' or 'device'")
        lc = self._lightcurve
        if optimizer == "scipy":
            solution = fitting.scipy_lbfgsb(self.gp, lc.y, initial_params, **opts)
            if np.all(np.isfinite(solution.x)):
                self.gp.set_parameter_vector(solution.x)
            return solution
        model = self.gp.device_model()
        starts = fitting.multistart_points(np.asarray(initial_params, dtype=np.float64), self.gp.get_parameter_bounds(),
                                           int(n_starts), np.random.default_rng(np.random.randint(0, 2 ** 31 - 1)))
        res = fitting.lbfgs_batch(model, lc.times, lc.y, lc.dy, starts, int(n_starts), ds_mean=None, **opts)
        fun = np.where(np.isfinite(res["fun"]), res["fun"], np.inf)
        i = int(np.argmin(fun))
        st = int(res["status"][i])
        solution = OptimizeResult(x=res["x"][i].copy(), fun=float(res["fun"][i]), nit=int(res["nit"][i]), status=st,
                                  success=(st == 0), message=fitting._MESSAGES.get(st, "?"))
        if np.all(np.isfinite(solution.x)):
            self.gp.set_parameter_vector(solution.x)
        return solution

    # -- posteriors -----------------------------------------------------------------------------------------
    def derive_posteriors(self, initial_chain_params=None, fit=True, converge=True, max_steps=10000,
                          convergence_steps=500, walkers=12, cores=6, progress=True, device_sampler=True):
        if initial_chain_params is None:
            if not fit:
                initial_params = self.initial_params
            else:
                initial_params = self.fit(self.initial_params).x
            initial_chain_params = self.spread_walkers(walkers, initial_params, np.array(self.gp.get_parameter_bounds()))
        every_samples = convergence_steps
        old_tau = np.inf
        self.converged = False
        y = self._lightcurve.y
        if device_sampler and not self.gp.device_prior_supported():
            warnings.warn("a kernel term overrides log_prior: sampling with the host loop (prior on the host, likelihood batches on the device)")
            device_sampler = False
        if device_sampler:
            # the whole stretch-move loop on the GPU (kernel K8), one launch per convergence interval
            sampler = DeviceEnsembleSampler(walkers, self.gp, y, block=every_samples)
        else:
            # host loop, one batched K1 launch per half step (emcee's vectorize=True seam)
            sampler = EnsembleSampler(walkers, self._ndim, lambda p: self.gp.log_probability_batch(y, p), vectorize=True)
        tau = None
        for sample in sampler.sample(initial_chain_params, iterations=max_steps, progress=progress):
            if sampler.iteration % every_samples:
                continue
            tau = sampler.get_autocorr_time(tol=0)
            self.autocorr.append(np.mean(tau))
            if np.all(tau * 100 < sampler.iteration) and np.all(np.abs(old_tau - tau) / tau < 0.01) and converge:
                print("Convergence reached after %d samples!" % sampler.iteration)
                self.converged = True
                break
            old_tau = tau
        if tau is None:
            # the reference hits a NameError here (max_steps < convergence_steps); estimate once instead
            tau = sampler.get_autocorr_time(tol=0)
        self._tau = tau
        mean_tau = np.mean(tau)
        if not self.converged:
            warnings.warn(f"The chains did not converge after {sampler.iteration} iterations!")
            thin = int(mean_tau / 4)
            discard = int(mean_tau) * 5
        else:
            discard = int(mean_tau * 40)
            if discard > max_steps:
                discard = int(mean_tau * 10)
            thin = int(mean_tau / 2)
        thin = max(thin, 1)                     # the reference passes thin=0 to emcee for tau < 2 (slice error)
        discard = min(discard, max(sampler.iteration - thin, 0))
        self._loglikelihoods = sampler.get_log_prob(discard=discard, thin=thin, flat=True)
        self._mcmc_samples = sampler.get_chain(discard=discard, thin=thin, flat=True)
        self._sampler = sampler

    def spread_walkers(self, walkers, parameters, bounds, percent=0.1, max_attempts=20):
        """Gaussian ball around ``parameters`` (sigma = ``percent`` * |parameter|) inside ``bounds``: walkers
        falling outside are redrawn up to ``max_attempts`` times, then set to 1.05 / 0.95 of the violated
        bound (gpmodelling.py:289-350)."""
        if percent < 0 or percent > 1:
            raise ValueError("The 'percent' parameter must be between 0 and 1 (inclusive).")
        parameters = np.asarray(parameters, dtype=np.float64)
        std = np.abs(parameters) * percent
        initial_samples = np.random.normal(parameters, std, size=(walkers, len(parameters)))
        bounds = np.array([(-np.inf if lower is None else lower, np.inf if upper is None else upper)
                           for lower, upper in bounds], dtype=np.float64)
        factors_lower = np.where(bounds[:, 0] > 0, 1.05, 0.95)
        factors_upper = np.where(bounds[:, 1] > 0, 0.95, 1.05)
        inside = lambda v: np.all(np.logical_and(bounds[:, 0] <= v, v <= bounds[:, 1]))
        for i in range(walkers):
            attempt = 0
            for attempt in range(max_attempts):
                if inside(initial_samples[i]):
                    break
                initial_samples[i] = np.random.normal(parameters, std)
            if attempt == max_attempts - 1:
                warnings.warn("Some walkers are out of bounds! Setting them to values close to the bounds")
                low = initial_samples[i] < bounds[:, 0]
                up = initial_samples[i] > bounds[:, 1]
                with np.errstate(invalid="ignore"):
                    initial_samples[i][low] = (bounds[:, 0] * factors_lower)[low]
                    initial_samples[i][up] = (bounds[:, 1] * factors_upper)[up]
        return initial_samples

    def standarized_residuals(self, include_noise=True):
        pred_mean, pred_var = self.gp.predict(self._lightcurve.y, return_var=True, return_cov=False)
        if include_noise:
            pred_var += self.gp.kernel.jitter
        return (self._lightcurve.y - pred_mean) / np.sqrt(pred_var)

    def get_rstat(self, burnin=None):
        if self._sampler is None:
            raise ValueError("Posteriors have not been derived. Please run derive_posteriors prior to populate the attributes.")
        if burnin is None:
            burnin = int(np.mean(self.tau)) * 10
        samples = self._sampler.get_chain(discard=burnin)
        whithin_chain_variances = np.var(samples, axis=0)
        samples = self._sampler.get_chain(flat=True, discard=burnin)
        between_chain_variances = np.var(samples, axis=0)
        return whithin_chain_variances / between_chain_variances[np.newaxis, :]

    # -- results ----------------------------------------------------------------------------------------------
    def _need(self, value):
        if value is None:
            raise AttributeError("Posteriors have not been derived. Please run derive_posteriors prior to populate the attributes.")
        return value

    @property
    def loglikelihoods(self):
        return self._need(self._loglikelihoods)

    @property
    def autocorr(self):
        return self._autocorr

    @property
    def sampler(self):
        self._need(self._loglikelihoods)
        return self._sampler

    @property
    def mcmc_samples(self):
        return self._need(self._mcmc_samples)

    @property
    def max_loglikelihood(self):
        return np.max(self._need(self._loglikelihoods))

    @property
    def max_parameters(self):
        return self._need(self._mcmc_samples)[np.argmax(self._loglikelihoods)]

    @property
    def median_parameters(self):
        return np.median(self._need(self._mcmc_samples), axis=0)

    @property
    def parameter_names(self):
        return self.gp.get_parameter_names()

    @property
    def k(self):
        return self._ndim

    @property
    def tau(self):
        self._need(self._mcmc_samples)
        return self._tau

    def gather_posteriors(self, device=None):
        """Pool the posterior draws of replica chains over the ranks of an initialised ``torch.distributed`` group.

        The ensemble of ONE lightcurve does not shard (its half steps are sequential), so on several GPUs every rank
        runs ``derive_posteriors`` on its own replica (give the ranks different ``np.random`` seeds) and this call
        all-gathers ``mcmc_samples`` and ``loglikelihoods`` (NCCL when ``device`` is a CUDA device, gloo on CPU): the only
        communication of the sampling path.  Afterwards the result properties describe the pooled sample on every rank.
        Returns the number of draws contributed by each rank."""
        import torch.distributed as dist
        from .lrt import all_gather_varlen
        self._need(self._mcmc_samples)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return [len(self._mcmc_samples)]
        rows = np.concatenate([self._mcmc_samples, self._loglikelihoods[:, None]], axis=1)
        pooled, counts = all_gather_varlen(dist, rows, device)
        self._mcmc_samples = np.ascontiguousarray(pooled[:, :-1])
        self._loglikelihoods = np.ascontiguousarray(pooled[:, -1])
        return counts

    # -- simulation from the posteriors -------------------------------------------------------------------------
    def generate_from_posteriors(self, nsims=10, cpus=8, pdf="Gaussian", extension_factor=2, sigma_noise=None, fft_length="fast"):
        """``nsims`` lightcurves, each simulated from one random posterior sample of the kernel parameters
        with the observing window, exposures and noise model of the data (gpmodelling.py:478-539).

        ``fft_length``: only the binned lightcurves leave this call, so the internal regular grid is lengthened (by < 0.5 %) to
        the next 7-smooth length by default -- the same process, statistically (tests/test_simulator_stats_gpu.py), without
        the Bluestein FFT a real window's grid length usually needs; ``"exact"`` keeps the reference's own grid length."""
        if self._mcmc_samples is None:
            raise RuntimeError("Posteriors have not been derived. Please run derive_posteriors prior to calling this method.")
        if nsims >= len(self._mcmc_samples):
            warnings.warn("The number of simulation requested (%d) is higher than the number of posterior samples (%d), "
                          "so many samples will be drawn more than once" % (nsims, len(self._mcmc_samples)))
        param_samples = self._mcmc_samples[np.random.randint(len(self._mcmc_samples), size=nsims)]
        simulator = self._lightcurve.get_simulator(self.gp.kernel.get_psd, pdf, sigma_noise=sigma_noise,
                                                   extension_factor=extension_factor, fft_length=fft_length)
        nk = self.gp.kernel.vector_size
        rates = simulator.generate_lightcurves(nsims, parameters=param_samples[:, :nk])
        noisy, dy = simulator.add_noise(rates)
        simulator.close()
        return [GappyLightcurve(self._lightcurve.times, noisy[i], dy[i]) for i in range(nsims)]

    def _generate_lc_from_params(self, parameters, simulator):
        """One lightcurve from one parameter vector with ``simulator`` (gpmodelling.py:516-539): the unit of work the
        reference maps over a process pool; ``generate_from_posteriors`` above does all of them in one batched call."""
        self.gp.set_parameter_vector(parameters)
        simulator.psd_model = self.gp.kernel.get_psd
        rates = simulator.generate_lightcurve()
        noisy_rates, dy = simulator.add_noise(rates)
        return GappyLightcurve(self._lightcurve.times, noisy_rates, dy)
