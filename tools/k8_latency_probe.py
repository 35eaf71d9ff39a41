"""K8 in the latency regime (hundreds to a few thousand ensembles of 30 walkers): ms per step against the ensemble count -- GPU box only.
(Round 2 used it to A/B a software-pipelined variant of the likelihood loop, measured slower and removed: DESIGN.md section 4.)"""
import ctypes, os, sys, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native
import bench
lib = _native.lib(); dev = torch.device("cuda", 0)
t, y, dy, p, bounds = bench.workload(64 * 2368, 1)
gp = bench.make_gp(t, y, dy, bounds); model = gp.device_model()
d_t = torch.from_numpy(t).to(dev); d_y = torch.from_numpy(y).to(dev); d_d = torch.from_numpy((dy + 1e-12) ** 2).to(dev)
d_p = torch.from_numpy(p).to(dev)
s = torch.cuda.current_stream()
for W, E, nsteps in ((30, 160, 50), (30, 625, 50), (30, 1250, 50), (30, 2500, 50), (30, 5000, 25)):
    pos = d_p[:E * W].clone().reshape(E, W, 5).contiguous()
    lpw = torch.empty((E, W), dtype=torch.float64, device=dev)
    def run():
        _native.check(lib.mtg_ensemble_dev(ctypes.byref(model), 1000, 1, E, W, d_t.data_ptr(), d_y.data_ptr(), d_d.data_ptr(), None,
                                           pos.data_ptr(), lpw.data_ptr(), 1, nsteps, 2.0, 12345, 0, None, None, None, None, None, 0, s.cuda_stream))
    run(); torch.cuda.synchronize()
    pos.copy_(d_p[:E * W].reshape(E, W, 5))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s); run(); e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    h = hashlib.sha1(pos.cpu().numpy().tobytes()).hexdigest()[:10]
    print("K8 W=%d E=%d steps=%d: %.2f ms  %.3f ms/step  %.3g evals/s  sha1(pos) %s" % (W, E, nsteps, ms, ms / nsteps, E * W * (nsteps + 1) / ms * 1e3, h), flush=True)
