"""The command-line workflow end to end on a small simulated lightcurve -- GPU box only."""
import os, sys, tempfile, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mind_the_gaps_b200 import workflow
from mind_the_gaps_b200.lightcurves import GappyLightcurve
from mind_the_gaps_b200.models.psd_models import BendingPowerlaw
from mind_the_gaps_b200.simulator import Simulator
warnings.simplefilter("ignore")
np.random.seed(3)
d = tempfile.mkdtemp()
times = np.arange(0, 300) * 86400.0 + 5e8
sim = Simulator(BendingPowerlaw(100.0, 2 * np.pi / (20 * 86400.0)), times, np.full(300, 2000.0), 100, pdf="Gaussian", extension_factor=2)
r = sim.generate_lightcurve(); y, dy = sim.add_noise(r)
GappyLightcurve(times, y, dy, exposures=2000.0).to_csv(os.path.join(d, "input.dat"))
open(os.path.join(d, "null.config"), "w").write("DampedRandomWalk log_S0=4.6:-10:15 log_omega0=-12.5:-20:-5\n")
open(os.path.join(d, "alt.config"), "w").write("Lorentzian log_S0=3:-10:15 log_Q=3:0.5:7 log_omega0=-11.8:-16:-9\nDampedRandomWalk log_S0=4.6:-10:15 log_omega0=-12.5:-20:-5\n")
a = ["--outdir", d]
f_null = workflow.fit_lightcurve(os.path.join(d, "input.dat"), os.path.join(d, "null.config"), "output", True, max_steps=3000, outdir=d)
sims = workflow.generate_lightcurves(f_null, os.path.join(d, "input.dat"), os.path.join(d, "null.config"), n=24, outdir=d)
files = [os.path.join(sims, "lightcurves", f) for f in os.listdir(os.path.join(sims, "lightcurves"))]
rn = workflow.fit_lightcurves(files, os.path.join(d, "null.config"), "sims", d); ra = workflow.fit_lightcurves(files, os.path.join(d, "alt.config"), "sims", d)
dn = workflow.fit_lightcurves([os.path.join(d, "input.dat")], os.path.join(d, "null.config"), "data", d)
da = workflow.fit_lightcurves([os.path.join(d, "input.dat")], os.path.join(d, "alt.config"), "data", d)
out = workflow.ratio_test(*[os.path.join(x, "fits_results.dat") for x in (rn, ra, dn, da)])
print("five steps: T_obs %.3f p %.3f T_dist median %.2f (24 sims)" % (out["T_obs"], out["p_value"], np.median(out["T_dist"])))
workflow.main(["ppp", os.path.join(d, "input.dat"), "-c", os.path.join(d, "null.config"), "-a", os.path.join(d, "alt.config"), "-n", "200",
               "--max-steps", "3000", "--refine-steps", "100", "--outdir", d])
