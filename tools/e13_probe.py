"""E13 (log-normal PDF) simulator stage of config C3 -- GPU box only.  python tools/e13_probe.py [n_lc]"""
import os, sys, time, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import psd_models
from helpers import window
n_lc = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 1000
day = 86400.0
tt_, exposures = window(N, 3)
lc2 = GappyLightcurve(tt_, np.full(N, 0.01), np.full(N, 0.002), exposures, bkg_rate=np.full(N, 0.001), bkg_rate_err=np.full(N, 1e-4))
psd2 = psd_models.BendingPowerlaw(S0=(0.3 * 0.01) ** 2, omega0=2 * np.pi / (20 * day))
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
sim = lc2.get_simulator(psd2, "Lognormal", extension_factor=2, fft_length=mode, max_iter=400)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    sim.generate_lightcurves(4, seed=1)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rates, info = sim.generate_lightcurves(n_lc, seed=4 + rep, return_info=True)
        torch.cuda.synchronize(); wall = time.perf_counter() - t0
        import hashlib
        print("%d lcs: %.3f s  %.1f lc/s  iterations mean %.1f max %d  n_fft %d  sha1(rates) %s" % (n_lc, wall, n_lc / wall, np.mean(info["e13_iterations"]),
              np.max(info["e13_iterations"]), sim.fftndatapoints, hashlib.sha1(np.ascontiguousarray(rates).tobytes()).hexdigest()[:12]), flush=True)
