"""K11 (scan gradient) vs K2 latency over batch sizes -- GPU box only.  python tools/k11_probe.py [N]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
lib = _native.lib(); dev = torch.device("cuda", 0)
bench.N_POINTS = N
t, y, dy, p, bounds = bench.workload(8192, 1)
gp = bench.make_gp(t, y, dy, bounds); model = gp.device_model()
d_t = torch.from_numpy(t).to(dev); d_y = torch.from_numpy(y).to(dev); d_d = torch.from_numpy((dy + 1e-12) ** 2).to(dev)
d_p = torch.from_numpy(p).to(dev)
s = torch.cuda.current_stream()
for B in (1, 32, 148, 296, 592, 1184, 2368, 4736, 8192):
    ll = torch.empty(B, dtype=torch.float64, device=dev); st = torch.empty(B, dtype=torch.int32, device=dev)
    g = torch.empty((B, 5), dtype=torch.float64, device=dev)
    res = {}
    for name, flag in (("k2", _native.FLAG_NO_SCAN), ("scan", _native.FLAG_SCAN)):
        def run():
            _native.check(lib.mtg_loglike_grad_dev(ctypes.byref(model), N, 1, B, B, d_t.data_ptr(), d_y.data_ptr(), d_d.data_ptr(), None,
                                                   d_p.data_ptr(), ll.data_ptr(), g.data_ptr(), st.data_ptr(), flag, s.cuda_stream))
        run(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(s)
        for _ in range(reps): run()
        e1.record(s); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps
        res[name + "_g"] = g.clone()
    err = float((res["k2_g"] - res["scan_g"]).abs().max() / res["k2_g"].abs().max())
    print("N=%d B=%5d  k2 %.4f ms  scan-grad %.4f ms  speedup %.2fx  max rel diff %.1e" % (N, B, res["k2"], res["scan"], res["k2"] / res["scan"], err), flush=True)
