import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from scipy.optimize import minimize
from mind_the_gaps_b200 import fitting
from mind_the_gaps_b200.celerite import GP
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from oracle import gp as ogp
from test_lbfgs_gpu import make_set, DAY
S, N = 6, 300
t, Y, DY, truth = make_set(S, N, 42)
bounds = [(-15.0, 5.0), (-20.0, -5.0), (-15.0, 5.0), (-2.0, 8.0), (-20.0, -5.0)]
x0 = np.array([truth[0], truth[1], truth[0] - 2.0, np.log(10.0), np.log(2 * np.pi / (10 * DAY))])
kernel = DampedRandomWalk(x0[0], x0[1], bounds=bounds[:2]) + Lorentzian(x0[2], x0[3], x0[4], bounds=bounds[2:])
gp = GP(kernel, mean=0.0)
spec = [("drw",), ("lorentzian",)]
res = fitting.fit_lightcurves(gp, t, Y, DY, n_starts=1)
for s in range(S):
    f = lambda p: -ogp.loglike_params(spec, p[None, :], t, Y[s], DY[s])[0]
    g = lambda p: -ogp.grad_complex_step(spec, p[None, :], t, Y[s], DY[s])[0]
    ref = minimize(f, x0, method="L-BFGS-B", bounds=bounds)
    refg = minimize(f, x0, jac=g, method="L-BFGS-B", bounds=bounds)
    cont = minimize(f, res[s].x, jac=g, method="L-BFGS-B", bounds=bounds)
    print(s, "gpu", res[s].fun, res[s].nit, res[s].status, np.round(res[s].x, 3))
    print("   scipy-fd", ref.fun, ref.nit, np.round(ref.x, 3))
    print("   scipy-jac", refg.fun, refg.nit, np.round(refg.x, 3))
    print("   scipy continued from gpu x:", cont.fun, cont.nit, np.round(cont.x, 3), "grad at gpu x", np.round(g(res[s].x), 4))
