import ctypes, os, sys, subprocess, time, threading
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native
import bench
lib = _native.lib(); dev = torch.device("cuda", 0)
t, y, dy, p, bounds = bench.workload(64 * 148, 1)
gp = bench.make_gp(t, y, dy, bounds); model = gp.device_model()
d_t = torch.from_numpy(t).to(dev); d_y = torch.from_numpy(y).to(dev); d_d = torch.from_numpy((dy + 1e-12) ** 2).to(dev)
d_p = torch.from_numpy(p).to(dev)
s = torch.cuda.current_stream()
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
rows = []
threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True).start()
def ev_time(fn, reps=1):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps): fn()
    e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ll = torch.empty(32, dtype=torch.float64, device=dev); st = torch.empty(32, dtype=torch.int32, device=dev)
def k1_small():
    _native.check(lib.mtg_loglike_dev(ctypes.byref(model), 1000, 1, 32, 32, d_t.data_ptr(), d_y.data_ptr(), d_d.data_ptr(), None,
                                      d_p.data_ptr(), ll.data_ptr(), st.data_ptr(), 1, s.cuda_stream))
print("K1 B=32 x1000: %.4f ms each" % ev_time(k1_small, 1000), "clocks", rows[-3:])
for W, E, nsteps in ((64, 1, 400), (64, 1, 400), (64, 148, 100)):
    pos = d_p[:E * W].clone().reshape(E, W, 5).contiguous(); lpw = torch.empty((E, W), dtype=torch.float64, device=dev)
    def run():
        _native.check(lib.mtg_ensemble_dev(ctypes.byref(model), 1000, 1, E, W, d_t.data_ptr(), d_y.data_ptr(), d_d.data_ptr(), None,
                                           pos.data_ptr(), lpw.data_ptr(), 1, nsteps, 2.0, 12345, 0, None, None, None, None, None, 0, s.cuda_stream))
    ms = ev_time(run)
    print("K8 W=%d E=%d steps=%d: %.2f ms  %.3f ms/half-step" % (W, E, nsteps, ms, ms / (2 * nsteps + 2)), "clocks", rows[-3:], flush=True)
proc.terminate()
