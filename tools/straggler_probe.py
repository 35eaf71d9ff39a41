"""Which L-BFGS starts are the stragglers of the LRT alternative fits?  GPU box only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mind_the_gaps_b200 import fitting, lrt
from mind_the_gaps_b200.celerite import GP
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from helpers import c2_truth, window

n_sims = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
N = 1000
tt, exposures = window(N, 2)
lc = GappyLightcurve(tt, np.full(N, 0.5), np.full(N, 0.02), exposures)
tr = c2_truth()
b_drw = dict(log_S0=(-15, 5), log_omega0=(-20, -5))
null_k = DampedRandomWalk(tr[0], tr[1], bounds=b_drw)
alt_k = Lorentzian(tr[2] - 2, np.log(20.0), tr[4], bounds=dict(log_S0=(-20, 5), log_Q=(0.5, 7), log_omega0=(-16, -9))) + \
    DampedRandomWalk(tr[0], tr[1], bounds=b_drw)
samples = tr[None, :2] + np.random.default_rng(5).normal(0, 0.1, (n_sims, 2))
res = lrt.simulate_and_fit(lc, null_k, alt_k, samples, n_starts_null=2, n_starts_alt=8, seed=3, distributed=False,
                           return_lightcurves=True, simulator_kwargs={"fft_length": "fast"})
Y, DY = res.lightcurves
gp1 = GP(alt_k, mean=0.0)
lo, hi = lrt._bounds_arrays(gp1)
rng = np.random.default_rng(1)
x1 = lrt.alternative_starts(gp1.get_parameter_vector().copy(), res.params_null, [3, 4], lo, hi, 8, rng).reshape(n_sims * 8, 5)
model = gp1.device_model()
out = {}
for mi in (25, 50, 100, 200, 400):
    out[mi] = fitting.lbfgs_batch(model, tt, Y, DY, x1, 8, ds_mean=Y.mean(axis=1), max_iter=mi)
r = out[400]
print("nit percentiles (max_iter 400):", np.percentile(r["nit"], [50, 90, 99, 99.9, 100]), "status counts", np.bincount(r["status"], minlength=3))
best400 = np.where(np.isfinite(r["fun"]), r["fun"], np.inf).reshape(n_sims, 8).min(axis=1)
for mi in (25, 50, 100, 200):
    b = np.where(np.isfinite(out[mi]["fun"]), out[mi]["fun"], np.inf).reshape(n_sims, 8).min(axis=1)
    d = b - best400
    print("max_iter %3d: group-best objective minus the max_iter=400 one: max %.3g  p99 %.3g  #>1e-3: %d  #>1e-2: %d   (T changes by 2x this)"
          % (mi, d.max(), np.percentile(d, 99), np.sum(d > 1e-3), np.sum(d > 1e-2)))
slow = np.flatnonzero(out[200]["nit"] >= 100)
print("%d starts with nit >= 100 at max_iter 200" % len(slow))
for i in slow[:12]:
    g = i // 8
    print("start %d (group %d, k=%d): x0=%s" % (i, g, i % 8, np.round(x1[i], 2)))
    for mi in (25, 50, 100, 200, 400):
        print("   max_iter %3d: nit %3d status %d fun %.6f x=%s   group best %.6f" % (mi, out[mi]["nit"][i], out[mi]["status"][i], out[mi]["fun"][i],
              np.round(out[mi]["x"][i], 3), best400[g]))
