"""Tiny K9 / K10 / K8 workload for compute-sanitizer (memcheck / racecheck / synccheck) -- GPU box only.

    compute-sanitizer --tool racecheck python tools/sanitize_scan.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import C2_BOUNDS, c2_truth, fake_lightcurve
from mind_the_gaps_b200.celerite import GP
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from mind_the_gaps_b200.sampler import DeviceEnsembleSampler
t, y, dy, _ = fake_lightcurve(300, 1)
p0 = c2_truth()
kernel = DampedRandomWalk(p0[0], p0[1], bounds=C2_BOUNDS[:2]) + Lorentzian(p0[2], p0[3], p0[4], bounds=C2_BOUNDS[2:])
gp = GP(kernel, mean=float(np.mean(y)))
gp.compute(t, dy + 1e-12)
p = p0[None, :] + np.random.default_rng(0).normal(0, 0.05, (12, 5))
a = gp.log_likelihood_batch(y, p, kernel="scan"); b = gp.log_likelihood_batch(y, p, kernel="k1")
print("K9 vs K1 max rel diff %.1e" % np.max(np.abs(a - b) / np.abs(b)))
for kern in ("scan", "k8"):
    np.random.seed(1)
    s = DeviceEnsembleSampler(12, gp, y, seed=3, block=6, kernel=kern)
    s.run_mcmc(p, 6)
    print(kern, "chain checksum %.12f" % float(np.sum(s.get_chain())))
# round 2: the parallel-in-time gradient K11, K2 beside it, and the persistent L-BFGS kernel (a small multi-start fit)
_, g11 = gp.grad_log_likelihood_batch(y, p, kernel="scan"); _, g2 = gp.grad_log_likelihood_batch(y, p, kernel="k2")
print("K11 vs K2 gradient max rel diff %.1e" % (np.max(np.abs(g11 - g2)) / np.max(np.abs(g2))))
from mind_the_gaps_b200 import fitting
res = fitting.fit_lightcurves(gp, t, np.stack([y, y[::-1]]), np.stack([dy, dy]), x0=p0, n_starts=3, rng=np.random.default_rng(0), max_iter=8)
print("persistent L-BFGS: fun", [round(r.fun, 6) for r in res], "nit", [int(r.nit) for r in res])
