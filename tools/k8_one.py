import ctypes, os, sys
import numpy as np
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native
import bench
lib = _native.lib(); dev = torch.device("cuda", 0)
t, y, dy, p, bounds = bench.workload(64, 1)
gp = bench.make_gp(t, y, dy, bounds); model = gp.device_model()
d_t = torch.from_numpy(t).to(dev); d_y = torch.from_numpy(y).to(dev); d_d = torch.from_numpy((dy + 1e-12) ** 2).to(dev)
d_p = torch.from_numpy(p).to(dev)
s = torch.cuda.current_stream()
W, E, nsteps = 64, 1, 10
pos = d_p[:E * W].clone().reshape(E, W, 5).contiguous()
lpw = torch.empty((E, W), dtype=torch.float64, device=dev)
for _ in range(2):
    _native.check(lib.mtg_ensemble_dev(ctypes.byref(model), 1000, 1, E, W, d_t.data_ptr(), d_y.data_ptr(), d_d.data_ptr(), None,
                                       pos.data_ptr(), lpw.data_ptr(), 1, nsteps, 2.0, 12345, 0, None, None, None, None, None, 0, s.cuda_stream))
torch.cuda.synchronize()
