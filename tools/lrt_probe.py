"""Full-size LRT probe (config C4) -- GPU box only.

    python tools/lrt_probe.py [n_sims] [fft_length=fast|exact] [chunk]
Prints wall time, per-stage device times and the number of K2 rounds of each fit."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native, lrt
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from helpers import c2_truth, window

n_sims = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
chunk = int(sys.argv[3]) if len(sys.argv) > 3 and int(sys.argv[3]) > 0 else None
pipeline = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rw = int(sys.argv[5]) if len(sys.argv) > 5 else 0
rs = int(sys.argv[6]) if len(sys.argv) > 6 else 0
rb = int(sys.argv[7]) if len(sys.argv) > 7 else 25
N = 1000
tt, exposures = window(N, 2)
lc = GappyLightcurve(tt, np.full(N, 0.5), np.full(N, 0.02), exposures)
tr = c2_truth()
b_drw = dict(log_S0=(-15, 5), log_omega0=(-20, -5))
null_k = DampedRandomWalk(tr[0], tr[1], bounds=b_drw)
alt_k = Lorentzian(tr[2] - 2, np.log(20.0), tr[4], bounds=dict(log_S0=(-20, 5), log_Q=(0.5, 7), log_omega0=(-16, -9))) + \
    DampedRandomWalk(tr[0], tr[1], bounds=b_drw)
samples = tr[None, :2] + np.random.default_rng(5).normal(0, 0.1, (n_sims, 2))
lib = _native.lib()
for rep in range(2):
    l0 = lib.mtg_launch_count()
    t0 = time.perf_counter()
    res = lrt.simulate_and_fit(lc, null_k, alt_k, samples, n_starts_null=2, n_starts_alt=8, seed=3, distributed=False, chunk=chunk, pipeline=pipeline, refine_walkers=rw, refine_steps=rs, refine_block=rb,
                               simulator_kwargs={"fft_length": mode})
    wall = time.perf_counter() - t0
    print("rep %d: %d sims  wall %.3f s  %.0f sims/s  launches %d  timings %s  T median %.3f p95 %.3f  status_alt %s" % (
        rep, n_sims, wall, n_sims / wall, lib.mtg_launch_count() - l0, {k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.timings.items() if k != 'note'},
        np.median(res.T), np.percentile(res.T, 95), np.bincount(res.status_alt, minlength=3)), flush=True)
