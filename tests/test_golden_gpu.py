"""GPU: the CUDA path (through the reference-shaped Python API -> C ABI) against outputs of the UNMODIFIED
reference on identical inputs and identical noise arrays (tests/golden/ref_*.npz, made by
tests/golden/make_golden.py; see tests/test_golden_oracle.py for what the fixtures pin).

Tolerances: log-likelihoods 1e-9 relative (north_star); index work and NumPy-order arithmetic bit-exact;
FFT-dependent values to cuFFT-vs-pocketfft rounding."""
import os
import warnings

import numpy as np
import pytest

from mind_the_gaps_b200 import celerite as mcel
from mind_the_gaps_b200.gpmodelling import GPModelling
from mind_the_gaps_b200.lightcurves import GappyLightcurve, Lightcurve
from mind_the_gaps_b200.models import BendingPowerlaw as BendingPowerlawTerm
from mind_the_gaps_b200.models import Cosinus, DampedRandomWalk, Lorentzian
from mind_the_gaps_b200.models import psd_models
from mind_the_gaps_b200.noise_models import GaussianNoise, KraftNoise, PoissonNoise
from mind_the_gaps_b200.simulator import Simulator
from mind_the_gaps_b200 import stats as mstats

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sim():
    return np.load(os.path.join(G, "ref_simulator.npz"))


@pytest.fixture(scope="module")
def noise():
    return np.load(os.path.join(G, "ref_noise.npz"))


@pytest.fixture(scope="module")
def gp():
    return np.load(os.path.join(G, "ref_gp.npz"))


b_drw = dict(log_S0=(-10.0, 5.0), log_omega0=(-15.0, -3.0))
b_lor = dict(log_S0=(-12.0, 5.0), log_Q=(0.5, 7.0), log_omega0=(-12.0, -5.0))
b_bpl = dict(log_S0=(-12.0, 5.0), log_Q=(-14.0, 5.0), log_omega0=(-15.0, -3.0))
# the same kernel constructions as tests/golden/make_golden.py::gen_gp, with this package's classes
BUILDERS = {
    "drw": lambda: DampedRandomWalk(-3.0, -8.0, bounds=b_drw),
    "drw_lorentzian": lambda: DampedRandomWalk(-3.0, -8.0, bounds=b_drw) + Lorentzian(-4.0, 3.0, -8.5, bounds=b_lor),
    "lorentzian_drw": lambda: Lorentzian(-4.0, 3.0, -8.5, bounds=b_lor) + DampedRandomWalk(-3.0, -8.0, bounds=b_drw),
    "cosinus_drw": lambda: Cosinus(-4.0, -8.5, bounds=dict(log_S0=(-12, 5), log_omega0=(-12, -5))) + DampedRandomWalk(-3.0, -8.0, bounds=b_drw),
    "bendingpl": lambda: BendingPowerlawTerm(-3.0, -5.0, -8.0, bounds=b_bpl),
    "sho_real": lambda: mcel.terms.SHOTerm(-3.0, 1.0, -8.0, bounds=dict(log_S0=(-12, 5), log_Q=(-3, 5), log_omega0=(-12, -5)))
    + mcel.terms.RealTerm(-3.5, -9.0, bounds=dict(log_a=(-12, 5), log_c=(-14, -4))),
    "matern_jitter": lambda: mcel.terms.Matern32Term(-2.0, 9.0, bounds=dict(log_sigma=(-8, 3), log_rho=(5, 12)))
    + mcel.terms.JitterTerm(-3.0, bounds=dict(log_sigma=(-8, 0))),
}
CASES = ["gp_drw_none", "gp_drw_constant", "gp_drw_linear", "gp_drw_lorentzian_none", "gp_drw_lorentzian_constant",
         "gp_drw_lorentzian_linear", "gp_lorentzian_drw_none", "gp_cosinus_drw_none", "gp_bendingpl_none",
         "gp_sho_real_none", "gp_matern_jitter_none"]


@pytest.mark.parametrize("tag", CASES)
def test_gpmodelling_log_probability(gp, tag):
    """GPModelling.__init__/_build_mean_model/_log_probability/_neg_log_like (gpmodelling.py:29-169)"""
    assert sorted(CASES) == sorted(gp["gp_cases"])
    name, mean = tag[len("gp_"):].rsplit("_", 1)
    lc = GappyLightcurve(gp["gp_t"], gp["gp_y"], gp["gp_dy"], gp["gp_exposures"])
    gpm = GPModelling(lc, BUILDERS[name](), mean_model=None if mean == "none" else mean)
    assert list(gpm.gp.get_parameter_names()) == list(gp[tag + "_names"])
    P = gp[tag + "_params"]; lp_ref = gp[tag + "_logprob"]; nll_ref = gp[tag + "_negloglike"]
    np.testing.assert_array_equal(np.asarray(gpm.initial_params), P[0])
    bounds = np.array([[-np.inf if b[0] is None else b[0], np.inf if b[1] is None else b[1]] for b in gpm.gp.get_parameter_bounds()], dtype=float)
    np.testing.assert_array_equal(bounds, gp[tag + "_bounds"])
    # batched (one K1 launch) and scalar (the reference's call shape) entry points
    lp = gpm._log_probability(P)
    inside = np.isfinite(lp_ref)
    assert np.array_equal(np.isfinite(lp), inside)
    np.testing.assert_allclose(lp[inside], lp_ref[inside], rtol=1e-9)
    for i in (0, 3, 10, 11):
        one = gpm._log_probability(P[i])
        if inside[i]:
            assert abs(one - lp_ref[i]) <= 1e-9 * abs(lp_ref[i])
            assert abs(gpm._neg_log_like(P[i]) - nll_ref[i]) <= 1e-9 * abs(nll_ref[i])
        else:
            assert one == -np.inf


def test_term_psds(gp):
    w = gp["psd_omega"]
    np.testing.assert_allclose(DampedRandomWalk(-3.0, -8.0).get_psd(w), gp["psd_term_drw"], rtol=1e-13)
    np.testing.assert_allclose(Lorentzian(-4.0, 3.0, -8.5).get_psd(w), gp["psd_term_lorentzian"], rtol=1e-13)
    np.testing.assert_allclose(psd_models.BendingPowerlaw(S0=np.exp(-3.0), omega0=np.exp(-8.0))(w), gp["psd_model_bendingpowerlaw"], rtol=1e-14)
    np.testing.assert_allclose(psd_models.Lorentzian(S0=np.exp(-4.0), Q=np.exp(3.0), omega0=np.exp(-8.5))(w), gp["psd_model_lorentzian"], rtol=1e-14)
    np.testing.assert_allclose(psd_models.SHO(S0=0.3, Q=4.0, omega0=2e-3)(w), gp["psd_model_sho"], rtol=1e-14)
    np.testing.assert_allclose(psd_models.Matern32(sigma=0.7, rho=900.0)(w), gp["psd_model_matern32"], rtol=1e-14)
    np.testing.assert_allclose(psd_models.Matern52(sigma=0.7, rho=900.0)(w), gp["psd_model_matern52"], rtol=1e-14)
    np.testing.assert_allclose(psd_models.Jitter(sigma=0.3)(w), gp["psd_model_jitter"], rtol=1e-14)


def test_spread_walkers_same_draws(gp):
    """spread_walkers consumes np.random exactly like the reference (gpmodelling.py:289-350)"""
    if "spread_walkers_out" not in gp.files:
        pytest.skip("fixture absent")
    lc = GappyLightcurve(gp["gp_t"], gp["gp_y"], gp["gp_dy"], gp["gp_exposures"])
    gpm = GPModelling(lc, BUILDERS["drw_lorentzian"]())
    np.random.seed(9)
    w0 = gpm.spread_walkers(32, gp["spread_walkers_p0"], gp["spread_walkers_bounds"], 0.1)
    np.testing.assert_array_equal(w0, gp["spread_walkers_out"])


# ---- simulator ----------------------------------------------------------------------------------------------

def test_grid_constants_exact(sim):
    for name in sim["grid_cases"]:
        t = sim["grid_%s_times" % name]; e = sim["grid_%s_exposures" % name]
        af, ef, eps = sim["grid_%s_kw" % name]
        s = Simulator(psd_models.BendingPowerlaw(S0=1.0, omega0=1e-3), t, float(e) if e.ndim == 0 else e, 1.0,
                      aliasing_factor=af, extension_factor=ef, epsilon=eps)
        sim_dt, sim_duration, n_fft, g0, g1 = sim["grid_%s_scalars" % name]
        assert s.sim_dt == sim_dt and s.sim_duration == sim_duration and s.fftndatapoints == int(n_fft)
        assert s.sim_timestamps[0] == g0 and s.sim_timestamps[-1] == g1
        assert np.array_equal(np.array(s.strategy), sim["grid_%s_strategy" % name])


@pytest.mark.parametrize("tag,pdf", [("tk95", "Gaussian"), ("e13_lognormal", "Lognormal"), ("e13_uniform", "Uniform")])
def test_generate_lightcurve_chain(sim, tag, pdf):
    """Simulator.generate_lightcurve of the reference (simulator.py:397-420) on its own recorded draws"""
    t = sim[tag + "_times"]; e = sim[tag + "_exposures"]
    S0, w0, mean = sim[tag + "_psd_S0_w0_mean"]
    psd = psd_models.BendingPowerlaw(S0=S0, omega0=w0)
    s = Simulator(psd, t, e, mean, pdf=pdf, aliasing_factor=2, extension_factor=2, max_iter=400)
    normals = sim[tag + "_normals"]
    # K3: the regularly sampled TK95 lightcurve (cuFFT vs pocketfft rounding)
    reg = s.simulate_regularly_sampled(normals=normals[None])
    ref_reg = sim[tag + "_regular_countrate"]
    scale = np.std(ref_reg)
    # (the reference's arbitrary DC term of 1e6 leaves ~1e6*eps/sqrt(N) of its own rounding noise in every sample)
    np.testing.assert_allclose(reg.countrate, ref_reg, rtol=0, atol=1e-10 * scale)
    draw, low, high = sim[tag + "_uniform"]
    u01 = (draw - low) / (high - low)
    kw = {}
    if pdf != "Gaussian":
        draws = sim[tag + "_pdf_draws"]; std = sim[tag + "_sample_std"][0]
        if pdf == "Lognormal":
            d = mstats.create_log_normal(mean, std)
            z = np.log(draws / d.kwds["scale"]) / d.args[0]
        else:
            d = mstats.create_uniform_distribution(mean, std)
            z = (draws - d.kwds["loc"]) / d.kwds["scale"]
        pv = np.zeros((1, s.fftndatapoints)); pv[0, :len(z)] = z
        kw["pdf_variates"] = pv
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rates, info = s.generate_lightcurves(1, normals=normals[None], u_segment=np.array([u01]), return_info=True, **kw)
    # K5: segment indices are integer work -> exact
    seg_t = sim[tag + "_segment_time"]
    i0 = int(np.searchsorted(s.sim_timestamps, draw))
    assert (int(info["segment"][0, 0]), int(info["segment"][0, 1])) == (i0, len(seg_t))
    ref = sim[tag + "_rates"]
    if pdf == "Gaussian":
        np.testing.assert_allclose(rates[0], ref, rtol=0, atol=1e-10 * scale)
    else:
        # E13 output = a permutation of the draws decided by ranks of FFT outputs: identical except where
        # cuFFT-vs-pocketfft rounding swaps near-tied ranks; bin means then differ by O(spacing of draws)
        close = np.isclose(rates[0], ref, rtol=1e-6, atol=0)
        assert close.mean() > 0.9
        np.testing.assert_allclose(rates[0], ref, rtol=2e-3)
    s.close()


@pytest.mark.parametrize("exposures", [0.5, 0.6, 0.1])
def test_reference_downsampling_cases(sim, exposures):
    timestamps = np.append(np.arange(1, 3.1, 2), np.arange(5, 7.1, 2))
    s = Simulator(psd_models.BendingPowerlaw(), timestamps, exposures, 1.0, aliasing_factor=1, extension_factor=1.0)
    lc = Lightcurve(np.arange(0.5, 10.1, 0.1), sim["downsample_%g_in" % exposures], input_counts=False, dt=0.1)
    np.testing.assert_array_equal(np.array(s.downsample(lc)), sim["downsample_%g_rates" % exposures])


def test_downsample_of_reference_segment_exact(sim):
    """K5 binning of the reference's own segment (time, countrate): bit-exact bin means"""
    for tag in ("tk95", "e13_lognormal"):
        s = Simulator(psd_models.BendingPowerlaw(), sim[tag + "_times"], sim[tag + "_exposures"], 2.0, aliasing_factor=2, extension_factor=2)
        lc = Lightcurve(sim[tag + "_segment_time"], sim[tag + "_segment_countrate"], input_counts=False, dt=s.sim_dt)
        np.testing.assert_array_equal(np.array(s.downsample(lc)), sim[tag + "_rates"])


# ---- noise ---------------------------------------------------------------------------------------------------

def test_poisson_noise_exact(noise):
    r, dy = PoissonNoise(noise["poisson_exposures"]).add_noise(noise["poisson_rates"], counts=noise["poisson_counts"].astype(np.int64))
    np.testing.assert_array_equal(r, noise["poisson_noisy"]); np.testing.assert_array_equal(dy, noise["poisson_dy"])


def test_gaussian_noise(noise):
    r, dy = GaussianNoise(noise["poisson_exposures"], noise["gaussian_sigma"][0]).add_noise(noise["poisson_rates"], normals=noise["gaussian_normals"])
    np.testing.assert_allclose(r, noise["gaussian_noisy"], rtol=1e-15); np.testing.assert_array_equal(dy, noise["gaussian_dy"])


def test_kraft_noise(noise):
    r, dy = KraftNoise(noise["kraft_exposures"], noise["kraft_bkg_counts"], noise["kraft_bkg_rate_err"]).add_noise(
        noise["kraft_rates"], counts=noise["kraft_counts"].astype(np.int64))
    low = noise["kraft_counts"] < 15
    np.testing.assert_array_equal(r[~low], noise["kraft_noisy"][~low]); np.testing.assert_array_equal(dy[~low], noise["kraft_dy"][~low])
    np.testing.assert_allclose(r[low], noise["kraft_noisy"][low], rtol=1e-8)       # scipy's numerical median of the reference's kraft_pdf
    np.testing.assert_allclose(dy[low], noise["kraft_dy"][low], rtol=1e-7)        # KBN interval (stand-in: parity unpinned vs astropy)


def test_pdf_factories(noise):
    for (m, s), (sg, sc), (loc, scale) in zip(noise["pdf_factory_mean_std"], noise["lognormal_s_scale"], noise["uniform_loc_scale"]):
        d = mstats.create_log_normal(m, s)
        assert (d.args[0], d.kwds["scale"]) == (sg, sc)
        d = mstats.create_uniform_distribution(m, s)
        assert (d.kwds["loc"], d.kwds["scale"]) == (loc, scale)
