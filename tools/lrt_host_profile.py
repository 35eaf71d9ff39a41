"""Where does the host time of a 10^4-simulation LRT go?  GPU box only.  python tools/lrt_host_profile.py [n_sims] [pipeline]"""
import cProfile, os, pstats, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native, lrt
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from helpers import c2_truth, window
n_sims = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
pipeline = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = 1000
tt, exposures = window(N, 2)
lc = GappyLightcurve(tt, np.full(N, 0.5), np.full(N, 0.02), exposures)
tr = c2_truth()
b_drw = dict(log_S0=(-15, 5), log_omega0=(-20, -5))
null_k = DampedRandomWalk(tr[0], tr[1], bounds=b_drw)
alt_k = Lorentzian(tr[2] - 2, np.log(20.0), tr[4], bounds=dict(log_S0=(-20, 5), log_Q=(0.5, 7), log_omega0=(-16, -9))) + \
    DampedRandomWalk(tr[0], tr[1], bounds=b_drw)
samples = tr[None, :2] + np.random.default_rng(5).normal(0, 0.1, (n_sims, 2))
def run():
    return lrt.simulate_and_fit(lc, null_k, alt_k, samples, n_starts_null=2, n_starts_alt=8, seed=3, distributed=False, pipeline=pipeline,
                                simulator_kwargs={"fft_length": "fast"})
for rep in range(3):
    t0 = time.perf_counter(); res = run(); wall = time.perf_counter() - t0
    print("rep %d pipeline %d: wall %.3f s  timings %s" % (rep, pipeline, wall, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.timings.items() if k != 'note'}), flush=True)
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
