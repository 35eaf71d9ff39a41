"""GPModelling.fit on ONE lightcurve: device (persistent L-BFGS kernel) vs scipy-driven -- GPU box only."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from mind_the_gaps_b200.gpmodelling import GPModelling
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from helpers import C2_BOUNDS, c2_truth, fake_lightcurve
for N in (200, 1000):
    t, y, dy, _ = fake_lightcurve(N, 2)
    tr = c2_truth()
    for label, mk in (("null DRW", lambda: DampedRandomWalk(tr[0], tr[1], bounds=C2_BOUNDS[:2])),
                      ("alt DRW+Lorentzian", lambda: DampedRandomWalk(tr[0], tr[1], bounds=C2_BOUNDS[:2]) + Lorentzian(tr[2], tr[3], tr[4], bounds=C2_BOUNDS[2:]))):
        g = GPModelling(GappyLightcurve(t, y, dy), mk())
        for opt in ("device", "scipy"):
            g.fit(g.initial_params, optimizer=opt)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); sol = g.fit(g.initial_params, optimizer=opt); ts.append(time.perf_counter() - t0)
            print("N=%d %-20s %-7s fit %.3f ms (min of 5)  nit %d  fun %.6f  success %s" % (N, label, opt, min(ts) * 1e3, sol.nit, sol.fun, sol.success), flush=True)
