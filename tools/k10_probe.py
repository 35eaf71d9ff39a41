"""K10 (scan sampler) vs K8 for few ensembles -- GPU box only."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native
import bench
lib = _native.lib(); dev = torch.device("cuda", 0)
t, y, dy, p, bounds = bench.workload(64 * 8, 1)
gp = bench.make_gp(t, y, dy, bounds); model = gp.device_model()
d_t = torch.from_numpy(t).to(dev); d_y = torch.from_numpy(y).to(dev); d_d = torch.from_numpy((dy + 1e-12) ** 2).to(dev)
d_p = torch.from_numpy(p).to(dev)
s = torch.cuda.current_stream()
for W, E, nsteps in ((64, 1, 400), (12, 1, 400), (30, 1, 400), (64, 4, 400), (12, 24, 400), (128, 1, 200)):
    for name, flag in (("k8", _native.FLAG_NO_SCAN), ("k10", _native.FLAG_SCAN)):
        pos = d_p[:E * W].clone().reshape(E, W, 5).contiguous()
        lpw = torch.empty((E, W), dtype=torch.float64, device=dev)
        acc = torch.zeros((E, W), dtype=torch.int32, device=dev)
        def run():
            _native.check(lib.mtg_ensemble_dev(ctypes.byref(model), 1000, 1, E, W, d_t.data_ptr(), d_y.data_ptr(), d_d.data_ptr(), None,
                                               pos.data_ptr(), lpw.data_ptr(), 1, nsteps, 2.0, 12345, 0, None, None, acc.data_ptr(), None, None, flag, s.cuda_stream))
        run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(s); run(); e1.record(s); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("%s W=%d E=%d steps=%d: %.2f ms  %.0f steps/s  %.3g evals/s  acc %.2f" % (name, W, E, nsteps, ms, nsteps / ms * 1e3, E * W * (nsteps + 1) / ms * 1e3,
              float(acc.double().mean().item()) / (2 * nsteps)), flush=True)
