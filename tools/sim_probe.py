"""Simulator stage only (config C3, TK95 + Poisson) -- GPU box only.  python tools/sim_probe.py [n_lc] [fast|exact]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import psd_models
from helpers import window
n_lc = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
N = 1000
day = 86400.0
tt_, exposures = window(N, 3)
lc = GappyLightcurve(tt_, np.full(N, 0.5), np.full(N, 0.02), exposures)
psd = psd_models.BendingPowerlaw(S0=(0.1 * 0.5) ** 2, omega0=2 * np.pi / (20 * day))
sim = lc.get_simulator(psd, "Gaussian", extension_factor=2, fft_length=mode)
sim.generate_lightcurves(64, seed=1)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rates = sim.generate_lightcurves(n_lc, seed=4)
    torch.cuda.synchronize(); wall = time.perf_counter() - t0
    print("%d lcs n_fft %d fused %s: %.3f s  %.0f lc/s" % (n_lc, sim.fftndatapoints, sim.fused, wall, n_lc / wall), flush=True)
