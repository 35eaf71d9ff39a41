"""K9 vs K1 on a four-state-dimension kernel (two SHO terms), 32 series, through the host API -- GPU box only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import fake_lightcurve
from mind_the_gaps_b200.celerite import GP, terms
t, y, dy, _ = fake_lightcurve(1000, 2)
gp = GP(terms.SHOTerm(-6.0, np.log(3.0), -12.0) + terms.SHOTerm(-7.0, np.log(2.0), -11.0), mean=float(np.mean(y)))
gp.compute(t, dy + 1e-12)
p = gp.get_parameter_vector()[None, :] + np.random.default_rng(0).normal(0, 0.05, (32, 6))
for kern in ("k1", "scan"):
    gp.log_likelihood_batch(y, p, kernel=kern)
    t0 = time.perf_counter()
    for _ in range(200):
        gp.log_likelihood_batch(y, p, kernel=kern)
    print("J=4 B=32 %s: %.1f us per host call" % (kern, (time.perf_counter() - t0) / 200 * 1e6))
