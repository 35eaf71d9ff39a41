"""Diagnostics behind tests/test_lrt_null_gpu.py -- GPU box only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_lrt_null_gpu as T
from mind_the_gaps_b200 import lrt
from mind_the_gaps_b200.celerite import GP
from mind_the_gaps_b200.fitting import multistart_points
n_sims, ns0, ns1, seed = 200, 2, 8, 11
lc, truth, null_kernel, alt_kernel = T._problem()
samples = truth[None, :] + np.random.default_rng(5).normal(0, 0.15, (n_sims, 2))
res = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, samples, n_starts_null=ns0, n_starts_alt=ns1, seed=seed, return_lightcurves=True, chunk=n_sims, pipeline=1)
Y, DY = res.lightcurves
gp1 = GP(alt_kernel, mean=0.0)
b0 = [(-15.0, 5.0), (-20.0, -5.0)]; b1 = [(-20.0, 5.0), (0.5, 7.0), (-16.0, -9.0), (-15.0, 5.0), (-20.0, -5.0)]
rng = np.random.default_rng([seed, 0, 0])
x0 = multistart_points(samples, b0, ns0, rng).reshape(n_sims, ns0, 2)
lo1, hi1 = lrt._bounds_arrays(gp1)
x1 = lrt.alternative_starts(gp1.get_parameter_vector().copy(), res.params_null, lrt.infer_null_slots(null_kernel, alt_kernel), lo1, hi1, ns1, rng)
l0 = np.array([T._cpu_best([("drw",)], x0[i], b0, lc.times, Y[i], DY[i]) for i in range(n_sims)])
l1 = np.array([T._cpu_best([("lorentzian",), ("drw",)], x1[i], b1, lc.times, Y[i], DY[i]) for i in range(n_sims)])
Tc = -2 * (l0 - l1)
d = res.T - Tc
print("null: max |dL| %.2e ; device-cpu alt dL: pcts" % np.max(np.abs(res.loglike_null - l0)), np.percentile(res.loglike_alt - l1, [0, 1, 5, 25, 50, 75, 95, 99, 100]))
for th in (1e-6, 1e-4, 1e-3, 1e-2, 1e-1): print("frac |dT| < %g: %.3f" % (th, np.mean(np.abs(d) < th)))
print("device better by > 1e-3: %.3f ; cpu better by > 1e-3: %.3f" % (np.mean(d > 1e-3), np.mean(d < -1e-3)))
print("medians", np.median(res.T), np.median(Tc), "p95", np.percentile(res.T, 95), np.percentile(Tc, 95))
from scipy.stats import ks_2samp
print("KS p", ks_2samp(res.T, Tc).pvalue)
