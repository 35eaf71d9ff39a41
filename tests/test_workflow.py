"""The command-line workflow (mind_the_gaps_b200/workflow.py; reference docs/workflow.md): config parsing, the on-disk
artefacts of every step and the final ratio test.  CPU only: the numerical work of each step is injected (the real
steps are the GPU calls covered by tests/test_pipeline_gpu.py)."""
import os

import numpy as np
import pytest

from mind_the_gaps_b200 import workflow
from mind_the_gaps_b200.lightcurves import GappyLightcurve, SimpleLightcurve
from mind_the_gaps_b200.stats import lrt_pvalue

NULL_CFG = ["# null model", "DampedRandomWalk log_S0=-3:-10:5 log_omega0=-13:-20:-5"]
ALT_CFG = ["Lorentzian log_S0=-5:-15:5 log_Q=3:0.5:7 log_omega0 = -12:-16:-9   # the QPO",
           "DampedRandomWalk log_S0=-3:-10:5 log_omega0=-13:-20:-5", "mean = constant"]


def test_parse_config():
    kernel, mean, names = workflow.parse_config(ALT_CFG)
    assert names == ["Lorentzian", "DampedRandomWalk"] and mean == "constant"
    np.testing.assert_array_equal(kernel.get_parameter_vector(), [-5.0, 3.0, -12.0, -3.0, -13.0])
    assert list(kernel.get_parameter_bounds()) == [(-15.0, 5.0), (0.5, 7.0), (-16.0, -9.0), (-10.0, 5.0), (-20.0, -5.0)]
    assert workflow.model_label(names) == "Lorentzian_DampedRandomWalk"
    k2, mean2, _ = workflow.parse_config(["SHOTerm log_S0=-6 log_Q=1 log_omega0=-12", "JitterTerm log_sigma=-4:-10:0"])
    assert mean2 is None and len(k2.get_parameter_vector()) == 4
    for bad in (["Nonsense log_a=1"], ["DampedRandomWalk log_S0=-3"], ["DampedRandomWalk log_S0=-3:1 log_omega0=-13"], ["# nothing"]):
        with pytest.raises(ValueError):
            workflow.parse_config(bad)


def _write_lc(path, n=40, seed=0):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0, 3e7, n)) + 5e8
    GappyLightcurve(t, rng.uniform(0.1, 2, n), rng.uniform(0.01, 0.1, n), rng.uniform(300, 1500, n)).to_csv(path)
    return t


class FakeModelling(object):
    """stands in for GPModelling: deterministic 'posteriors' around the initial parameters"""
    def __init__(self, lc, kernel, mean_model=None):
        self.kernel = kernel; self.lc = lc; self.converged = True
        self.parameter_names = ["kernel:" + n for n in kernel.get_parameter_names()]

    def derive_posteriors(self, **kw):
        p0 = np.asarray(self.kernel.get_parameter_vector())
        rng = np.random.default_rng(len(p0))
        self.mcmc_samples = p0[None, :] + 0.01 * rng.normal(size=(25, len(p0)))
        self.loglikelihoods = -np.sum((self.mcmc_samples - p0) ** 2, axis=1) + 100.0 * len(p0)
        self.max_loglikelihood = self.loglikelihoods.max()
        self.max_parameters = self.mcmc_samples[np.argmax(self.loglikelihoods)]


def test_the_five_steps_write_the_documented_files(tmp_path):
    lc_file = str(tmp_path / "input_lightcurve.dat")
    t = _write_lc(lc_file)
    cfg_n, cfg_a = str(tmp_path / "null.config"), str(tmp_path / "alt.config")
    open(cfg_n, "w").write("\n".join(NULL_CFG)); open(cfg_a, "w").write("\n".join(ALT_CFG))
    # step 1: posteriors -> best_fit.dat, samples.dat
    d_null = workflow.fit_lightcurve(lc_file, cfg_n, "output", fit=True, outdir=str(tmp_path), modelling=FakeModelling)
    assert os.path.basename(d_null) == "celerite_output_t%d-%d_m_DampedRandomWalk_fit" % (int(t[0]), int(t[-1]))
    cols, best = workflow.read_table(os.path.join(d_null, "best_fit.dat"))
    assert cols[-1] == "loglikelihood" and best.shape == (1, 3)
    cols, samples = workflow.read_table(os.path.join(d_null, "samples.dat"))
    assert samples.shape == (25, 3) and np.all(np.abs(samples[:, 0] + 3) < 0.1)
    # step 2: synthetic lightcurves in the input format
    def simulate(lc, kernel, draws, pdf, ext):
        assert draws.shape == (7, 2) and pdf == "Lognormal" and ext == 2
        return np.outer(np.arange(1, 8), np.ones(lc.n)), np.full((7, lc.n), 0.05)
    np.random.seed(0)
    d_sim = workflow.generate_lightcurves(d_null, lc_file, cfg_n, n=7, pdf="Lognormal", outdir=str(tmp_path), simulate=simulate)
    files = sorted(os.listdir(os.path.join(d_sim, "lightcurves")))
    assert len(files) == 7
    sim3 = SimpleLightcurve(os.path.join(d_sim, "lightcurves", "lc3.dat"))
    np.testing.assert_allclose(sim3.times, t, rtol=1e-8); assert np.all(sim3.y == 4.0) and np.all(sim3.dy == 0.05)
    # steps 3-4: fits of both models to the simulations and to the data
    def fitter_for(offset):
        def fitter(lcs, kernel):
            ll = np.array([10.0 * l.y[0] + offset for l in lcs])
            return ll, np.tile(kernel.get_parameter_vector(), (len(lcs), 1))
        return fitter
    sims = [os.path.join(d_sim, "lightcurves", f) for f in files]
    r_ns = workflow.fit_lightcurves(sims, cfg_n, "sims", str(tmp_path), fitter=fitter_for(0.0))
    r_as = workflow.fit_lightcurves(sims, cfg_a, "sims", str(tmp_path), fitter=fitter_for(1.5))
    r_nd = workflow.fit_lightcurves([lc_file], cfg_n, "data", str(tmp_path), fitter=lambda l, k: (np.array([50.0]), np.atleast_2d(k.get_parameter_vector())))
    r_ad = workflow.fit_lightcurves([lc_file], cfg_a, "data", str(tmp_path), fitter=lambda l, k: (np.array([56.0]), np.atleast_2d(k.get_parameter_vector())))
    assert os.path.basename(r_as) == "fit_lcs_sims_m_Lorentzian_DampedRandomWalk"
    cols, tab = workflow.read_table(os.path.join(r_as, "fits_results.dat"))
    assert tab.shape == (7, 6) and cols[0] == "loglikelihood"
    # step 5: the ratio test
    out = workflow.ratio_test(*[os.path.join(d, "fits_results.dat") for d in (r_ns, r_as, r_nd, r_ad)])
    np.testing.assert_allclose(out["T_dist"], 3.0)
    assert out["T_obs"] == 12.0 and out["p_value"] == lrt_pvalue(np.full(7, 3.0), 12.0) == 0.0
    assert workflow.main(["ratio-test", "-n", os.path.join(r_ns, "fits_results.dat"), "-a", os.path.join(r_as, "fits_results.dat"),
                          "-dn", os.path.join(r_nd, "fits_results.dat"), "-da", os.path.join(r_ad, "fits_results.dat")]) == 0
