"""Time K2 (loglike + gradient) for a few batch sizes -- GPU box only."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native
import bench
B_list = [int(x) for x in (sys.argv[1:] or ["16384", "65536", "151552"])]
lib = _native.lib()
dev = torch.device("cuda", 0)
for B in B_list:
    t, y, dy, p, bounds = bench.workload(B, 1)
    gp = bench.make_gp(t, y, dy, bounds); model = gp.device_model()
    d_t = torch.from_numpy(t).to(dev); d_y = torch.from_numpy(y).to(dev); d_d = torch.from_numpy((dy + 1e-12) ** 2).to(dev)
    d_p = torch.from_numpy(p).to(dev)
    ll = torch.empty(B, dtype=torch.float64, device=dev); g = torch.empty((B, 5), dtype=torch.float64, device=dev)
    st = torch.empty(B, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream()
    def run():
        _native.check(lib.mtg_loglike_grad_dev(ctypes.byref(model), 1000, 1, B, B, d_t.data_ptr(), d_y.data_ptr(), d_d.data_ptr(), None,
                                               d_p.data_ptr(), ll.data_ptr(), g.data_ptr(), st.data_ptr(), 0, s.cuda_stream))
    run(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(3): run()
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("K2 B=%d  %.3f ms  %.3g grad-evals/s  tape %.0f GB/s" % (B, ms, B / ms * 1e3, 2 * B * 1000 * 14 * 8 / ms / 1e6), flush=True)
