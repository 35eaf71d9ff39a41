"""Device-resident simulator stage (mtg_sim_generate_dev, CUDA events on the launching stream) -- GPU box only.
    python tools/sim_dev_probe.py [n_lc] [fast|exact] [celerite|table]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models import DampedRandomWalk, psd_models
from helpers import window
n_lc = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
mode = sys.argv[2] if len(sys.argv) > 2 else "fast"
src = sys.argv[3] if len(sys.argv) > 3 else "celerite"
N = 1000
day = 86400.0
tt_, exposures = window(N, 3)
lc = GappyLightcurve(tt_, np.full(N, 0.5), np.full(N, 0.02), exposures)
if src == "celerite":
    k = DampedRandomWalk(np.log((0.1 * 0.5) ** 2), np.log(2 * np.pi / (20 * day)))
    psd = k.get_psd
else:
    psd = psd_models.BendingPowerlaw(S0=(0.1 * 0.5) ** 2, omega0=2 * np.pi / (20 * day))
sim = lc.get_simulator(psd, "Gaussian", extension_factor=2, fft_length=mode)
h = sim._sim()
model, params, table = sim._psd_source(None)
dev = torch.device("cuda:0")
lib = _native.lib()
rates = torch.empty((N, n_lc), dtype=torch.float64, device=dev)
p_dev = t_dev = None
if model is not None:
    p_dev = torch.from_numpy(np.ascontiguousarray(np.repeat(params, n_lc, axis=0))).to(dev)
else:
    t_dev = torch.from_numpy(table).to(dev)
stream = torch.cuda.current_stream().cuda_stream
def run(seed):
    _native.check(lib.mtg_sim_generate_dev(h, n_lc, ctypes.byref(model) if model is not None else None,
                                           p_dev.data_ptr() if p_dev is not None else None,
                                           t_dev.data_ptr() if t_dev is not None else None, seed, rates.data_ptr(), stream))
run(1); torch.cuda.synchronize()
for rep in range(3):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(2 + rep); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("%s %s: %d lcs n_fft %d fused %s: %.2f ms  %.0f lc/s  %.2f us/lc  finite %s std %.4f" % (
        mode, src, n_lc, sim.fftndatapoints, sim.fused, ms, n_lc / ms * 1e3, ms * 1e3 / n_lc,
        bool(torch.isfinite(rates).all()), float(rates.std())), flush=True)
