"""Config C5 sweep (N = 1e5, 3 x SHO + Real, J = 7) -- GPU box only.  python tools/c5_probe.py [n_sets]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mind_the_gaps_b200 import _native
import bench
n_sets = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
lib = _native.lib(); dev = torch.device("cuda", 0); stream = torch.cuda.current_stream()
def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
pk = ctypes.c_double(); ms = ctypes.c_double(); lib.mtg_fp64_peak_probe(ctypes.byref(pk), ctypes.byref(ms))
r = bench.c5_measurement(torch, lib, _native, dev, stream, timed, 0, 1, n_sets=n_sets)
print("C5: %d sets  %.1f ms  %.3g evals/s  %.2f TF algorithmic  frac %.3f of %.1f TF  finite %s" % (
    r["parameter_sets_per_gpu"], r["ms"], r["evals_per_s_per_gpu"], r["tflops_algorithmic_per_gpu"], r["tflops_algorithmic_per_gpu"] / pk.value, pk.value, r["all_finite"]))
