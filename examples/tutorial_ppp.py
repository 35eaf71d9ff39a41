#!/usr/bin/env python
"""The reference's tutorial (``docs/notebooks/tutorial_ppp.ipynb``: is the QPO of a lightcurve significant?  Protassov
likelihood-ratio test) with ``mind_the_gaps_b200`` -- same calls, cell by cell, plots left out.

    python examples/tutorial_ppp.py [--nsims 1000] [--max-steps 20000] [--qpo]      (needs a B200)

Cells 3-9 (simulate an input lightcurve, posteriors of the null and the alternative model) are the notebook's lines with
the import lines changed.  Cells 11-15 are shown twice: the notebook's own loop over the simulated lightcurves (a
``GPModelling`` per lightcurve and model: ``--loop`` of them) and the batched driver ``lrt.simulate_and_fit`` that does the
same work for all ``--nsims`` at once."""
import argparse
import os
import sys
import time
import warnings

import numpy as np
from scipy.stats import percentileofscore

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mind_the_gaps_b200 import celerite, lrt                                    # noqa: E402  (notebook: import celerite)
from mind_the_gaps_b200.gpmodelling import GPModelling                           # noqa: E402
from mind_the_gaps_b200.lightcurves import GappyLightcurve                       # noqa: E402
from mind_the_gaps_b200.models.psd_models import BendingPowerlaw, Lorentzian     # noqa: E402
from mind_the_gaps_b200.simulator import Simulator                               # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--nsims", type=int, default=1000, help="simulated lightcurves of the LRT (the notebook: 100, 'typically 10,000')")
    ap.add_argument("--max-steps", type=int, default=20000, help="max_steps of derive_posteriors (the notebook: 50000)")
    ap.add_argument("--loop", type=int, default=3, help="simulated lightcurves also treated with the notebook's per-lightcurve loop")
    ap.add_argument("--points", type=int, default=1000)
    ap.add_argument("--qpo", action="store_true", help="second half of the notebook: the input lightcurve does contain a QPO")
    ap.add_argument("--seed", type=int, default=10)
    args = ap.parse_args(argv)
    np.random.seed(args.seed)
    t_start = time.perf_counter()

    # ---- cells 3-5 (19-20 with --qpo): the input lightcurve ------------------------------------------------------
    times = np.arange(0, args.points)
    dt = np.diff(times)[0]
    mean = 100
    variance_drw = (mean * 0.1) ** 2
    w_bend = 2 * np.pi / 20
    P, Q = 10, 80
    w_qpo = 2 * np.pi / P
    psd_model = BendingPowerlaw(variance_drw, w_bend)
    if args.qpo:
        psd_model = Lorentzian(variance_drw, Q, w_qpo) + psd_model
    simulator = Simulator(psd_model, times, np.ones(len(times)) * dt, mean, pdf="Gaussian", extension_factor=2)
    countrates = simulator.generate_lightcurve()
    noisy_countrates, dy = simulator.add_noise(countrates)
    input_lc = GappyLightcurve(times, noisy_countrates, dy, exposures=dt)

    # ---- cell 7: null model ---------------------------------------------------------------------------------------
    bounds_drw = dict(log_a=(-10, 50), log_c=(-10, 10))
    null_kernel = celerite.terms.RealTerm(log_a=np.log(variance_drw), log_c=np.log(w_bend), bounds=bounds_drw)
    null_model = GPModelling(input_lc, null_kernel)
    print("Deriving posteriors for null model")
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        null_model.derive_posteriors(max_steps=args.max_steps, fit=True, cores=15, progress=False)
    print("  %d steps in %.2f s, converged: %s, median parameters %s" % (null_model.sampler.iteration, time.perf_counter() - t0,
                                                                       null_model.converged, np.round(null_model.median_parameters, 3)))

    # ---- cell 9: alternative model -----------------------------------------------------------------------------------
    log_c, log_d = np.log(0.5 * w_qpo / Q), np.log(w_qpo)
    bounds_qpo = dict(log_a=(-10, 50), log_c=(-10, 10), log_d=(-5, 5))
    alternative_kernel = celerite.terms.ComplexTerm(log_a=np.log(variance_drw), log_c=log_c, log_d=log_d, bounds=bounds_qpo) \
        + celerite.terms.RealTerm(log_a=np.log(variance_drw), log_c=np.log(w_bend), bounds=bounds_drw)
    alternative_model = GPModelling(input_lc, alternative_kernel)
    print("Deriving posteriors for alternative model")
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        alternative_model.derive_posteriors(max_steps=args.max_steps, fit=True, cores=15, progress=False)
    print("  %d steps in %.2f s, converged: %s" % (alternative_model.sampler.iteration, time.perf_counter() - t0, alternative_model.converged))
    T_obs = -2 * (null_model.max_loglikelihood - alternative_model.max_loglikelihood)

    # ---- cells 11-13, the notebook's way: one GPModelling per simulated lightcurve and model ------------------------
    T_loop = []
    if args.loop > 0:
        lcs = null_model.generate_from_posteriors(args.loop, cpus=15)
        t0 = time.perf_counter()
        for lc in lcs:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                null_modelling = GPModelling(lc, null_kernel)
                null_modelling.derive_posteriors(fit=True, cores=15, walkers=30, max_steps=500, progress=False)
                alternative_modelling = GPModelling(lc, alternative_kernel)
                alternative_modelling.derive_posteriors(fit=True, cores=15, walkers=30, max_steps=500, progress=False)
            T_loop.append(-2 * (null_modelling.max_loglikelihood - alternative_modelling.max_loglikelihood))
        print("notebook loop: %d lightcurves in %.2f s, T = %s" % (args.loop, time.perf_counter() - t0, np.round(T_loop, 2)))

    # ---- cells 11-13, batched: all simulations and all fits in lock step on the device ----------------------------------
    t0 = time.perf_counter()
    draws = null_model.mcmc_samples[np.random.randint(len(null_model.mcmc_samples), size=args.nsims)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = lrt.simulate_and_fit(input_lc, null_kernel, alternative_kernel, draws, refine_walkers=30, refine_steps=500)
    print("batched: %d simulations (L-BFGS + 30 walkers x 500 steps per model) in %.2f s" % (args.nsims, time.perf_counter() - t0))

    # ---- cell 15: the test ----------------------------------------------------------------------------------------
    T_dist = res.T
    perc = percentileofscore(T_dist, T_obs)
    print("Observed LRT_stat: %.3f" % T_obs)
    print("p-value: %.4f   (T_dist: median %.2f, 95 %% %.2f, 99.7 %% %.2f)" % (1 - perc / 100, np.median(T_dist), np.percentile(T_dist, 95),
                                                                              np.percentile(T_dist, 99.7)))
    print("total %.1f s" % (time.perf_counter() - t_start))
    return dict(T_obs=T_obs, T_dist=T_dist, p=1 - perc / 100, T_loop=np.array(T_loop), null=null_model, alternative=alternative_model)


if __name__ == "__main__":
    main()
