"""CPU: pin the oracle on outputs of the UNMODIFIED reference (tests/golden/ref_*.npz).

The fixtures were produced by ``tests/golden/make_golden.py`` in the build container: the reference's own
simulator.py / noise_models.py / stats.py / gpmodelling.py / models/*.py executed as they are, with the
stand-ins of ``tests/golden/_shims`` for the un-vendored third-party imports (README there says what each
provides; ``celerite.GP`` is a dense Cholesky of the published kernel).  Every random draw the reference made is
stored next to its outputs, so the oracle (here) and the CUDA path (tests/test_golden_gpu.py) are checked on
identical noise arrays."""
import os

import numpy as np
import pytest

from oracle import gp as ogp
from oracle import simulator as osim
from oracle import terms as oterms

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


# kernel spec (oracle.terms) of each golden GP case, in the reference's parameter order
SPECS = {
    "drw": [("drw",)],
    "drw_lorentzian": [("drw",), ("lorentzian",)],
    "lorentzian_drw": [("lorentzian",), ("drw",)],
    "cosinus_drw": [("cosinus",), ("drw",)],
    "bendingpl": [("bendingpl",)],
    "sho_real": [("sho",), ("real",)],
    "matern_jitter": [("matern32",), ("jitter",)],
}


def split_case(tag):
    body = tag[len("gp_"):]
    name, mean = body.rsplit("_", 1)
    return name, mean


# ---- Simulator.__init__ grid (simulator.py:195-258) ---------------------------------------------------

def test_grid_constants_exact(sim):
    for name in sim["grid_cases"]:
        t = sim["grid_%s_times" % name]; e = sim["grid_%s_exposures" % name]
        af, ef, eps = sim["grid_%s_kw" % name]
        cfg = osim.setup(t, float(e) if e.ndim == 0 else e, aliasing_factor=af, extension_factor=ef, epsilon=eps)
        sim_dt, sim_duration, n_fft, g0, g1 = sim["grid_%s_scalars" % name]
        assert cfg["sim_dt"] == sim_dt and cfg["sim_duration"] == sim_duration
        assert cfg["n_fft"] == int(n_fft) and cfg["grid"][0] == g0 and cfg["grid"][-1] == g1
        assert np.array_equal(np.array(cfg["strategy"]), sim["grid_%s_strategy" % name])


@pytest.mark.parametrize("N", [257, 1024])
def test_get_fft_exact(sim, N):
    re, im = sim["getfft_%d_normals" % N]
    out = osim.get_fft(N, 3.5, lambda w: oterms.psd_lorentzian(w, S0=2.0, Q=12.0, omega0=0.2), re, im)
    np.testing.assert_allclose(out, sim["getfft_%d_out" % N], rtol=1e-15, atol=0)


@pytest.mark.parametrize("tag", ["tk95", "e13_lognormal", "e13_uniform"])
def test_generate_lightcurve_chain(sim, tag):
    """simulate_regularly_sampled -> cut_random_segment -> shift -> adjust_pdf -> downsample of the reference"""
    t = sim[tag + "_times"]; e = sim[tag + "_exposures"]
    S0, w0, mean = sim[tag + "_psd_S0_w0_mean"]
    cfg = osim.setup(t, e, aliasing_factor=2, extension_factor=2)
    psd = lambda w: oterms.psd_bendingpowerlaw(w, S0=S0, omega0=w0)
    re, im = sim[tag + "_normals"]
    full = osim.simulate_regular(cfg["n_fft"], cfg["sim_dt"], psd, mean, re, im)
    np.testing.assert_array_equal(full, sim[tag + "_regular_countrate"])          # same numpy FFT: bit-exact
    draw, low, high = sim[tag + "_uniform"]
    assert low == cfg["grid"][0] and high == cfg["grid"][-1] - cfg["sim_duration"]
    u01 = (draw - low) / (high - low)
    i0, i1, shift = osim.cut_segment(cfg["grid"], cfg["sim_dt"], cfg["sim_duration"], cfg["strategy"][0][0], u01)
    seg_t = cfg["grid"][i0:i1] + shift
    assert len(seg_t) == len(sim[tag + "_segment_time"])
    np.testing.assert_array_equal(seg_t, sim[tag + "_segment_time"])
    np.testing.assert_array_equal(full[i0:i1], sim[tag + "_segment_countrate_tk95"])
    seg = full[i0:i1]
    if tag != "tk95":
        std = np.std(seg)
        assert std == sim[tag + "_sample_std"][0]
        seg, its = osim.e13_adjust(seg, sim[tag + "_pdf_draws"], max_iter=400)
        np.testing.assert_array_equal(seg, sim[tag + "_segment_countrate"])
    rates = osim.downsample(seg_t, seg, cfg["strategy"])
    np.testing.assert_array_equal(np.array(rates), sim[tag + "_rates"])


def test_e13_pdf_draws_come_from_the_factories(sim):
    """the recorded pdf.rvs draws are reproduced by the oracle's (sigma, scale) / (loc, scale) factories: the
    multiset statistics match the requested mean / std (stats.py:116-146)"""
    for tag, lo in (("e13_lognormal", True), ("e13_uniform", False)):
        draws = sim[tag + "_pdf_draws"]; mean = sim[tag + "_psd_S0_w0_mean"][2]; std = sim[tag + "_sample_std"][0]
        assert abs(draws.mean() - mean) < 5 * std / np.sqrt(len(draws))
        if lo:
            assert draws.min() > 0
        else:
            loc, scale = osim.create_uniform_params(mean, std)
            assert draws.min() >= loc and draws.max() <= loc + scale


@pytest.mark.parametrize("exposures", [0.5, 0.6, 0.1])
def test_reference_downsampling_cases(sim, exposures):
    timestamps = np.append(np.arange(1, 3.1, 2), np.arange(5, 7.1, 2))
    cfg = osim.setup(timestamps, exposures, aliasing_factor=1, extension_factor=1.0)
    times = np.arange(0.5, 10.1, 0.1)
    got = osim.downsample(times, sim["downsample_%g_in" % exposures], cfg["strategy"])
    np.testing.assert_array_equal(np.array(got), sim["downsample_%g_rates" % exposures])


# ---- noise models -----------------------------------------------------------------------------------

def test_poisson_noise_exact(noise):
    n = len(noise["poisson_rates"])
    r, dy = osim.poisson_noise(noise["poisson_rates"], noise["poisson_exposures"], np.zeros(n), np.zeros(n), noise["poisson_counts"])
    np.testing.assert_array_equal(r, noise["poisson_noisy"]); np.testing.assert_array_equal(dy, noise["poisson_dy"])


def test_gaussian_noise(noise):
    r, dy = osim.gaussian_noise(noise["poisson_rates"], noise["gaussian_sigma"][0], noise["gaussian_normals"])
    np.testing.assert_allclose(r, noise["gaussian_noisy"], rtol=1e-15); np.testing.assert_array_equal(dy, noise["gaussian_dy"])


def test_kraft_noise(noise):
    """rates: the reference's scipy-numerical median of ITS kraft_pdf (reference code end to end);
    dy of the low-count bins goes through the poisson_conf_interval stand-in (parity unpinned vs astropy)"""
    r, dy = osim.kraft_noise(noise["kraft_rates"], noise["kraft_exposures"], noise["kraft_bkg_counts"], noise["kraft_bkg_rate_err"],
                             noise["kraft_counts"])
    low = noise["kraft_counts"] < 15
    assert 5 < low.sum() < len(low) - 5
    np.testing.assert_array_equal(r[~low], noise["kraft_noisy"][~low]); np.testing.assert_array_equal(dy[~low], noise["kraft_dy"][~low])
    np.testing.assert_allclose(r[low], noise["kraft_noisy"][low], rtol=1e-8)
    np.testing.assert_allclose(dy[low], noise["kraft_dy"][low], rtol=1e-7)


def test_kraft_median_vs_reference_pdf(noise):
    for (N, B), med in zip(noise["kraft_median_NB"], noise["kraft_median"]):
        assert abs(osim.kraft_median(int(N), B) - med) < 1e-9 * max(1.0, med)


def test_pdf_factories(noise):
    for (m, s), (sg, sc), (loc, scale) in zip(noise["pdf_factory_mean_std"], noise["lognormal_s_scale"], noise["uniform_loc_scale"]):
        assert osim.create_log_normal_params(m, s) == (sg, sc)
        assert osim.create_uniform_params(m, s) == (loc, scale)


# ---- GPModelling glue + Term coefficient maps (reference code) on a dense-Cholesky GP ---------------------

def test_gp_cases_present(gp):
    assert len(gp["gp_cases"]) == 11


@pytest.mark.parametrize("tag", ["gp_drw_none", "gp_drw_constant", "gp_drw_lorentzian_none", "gp_drw_lorentzian_constant",
                                 "gp_lorentzian_drw_none", "gp_cosinus_drw_none", "gp_bendingpl_none", "gp_sho_real_none",
                                 "gp_matern_jitter_none"])
def test_log_probability_and_neg_log_like(gp, tag):
    """GPModelling._log_probability / _neg_log_like (gpmodelling.py:127-169): recursion oracle vs the reference's glue"""
    name, mean = split_case(tag)
    spec = SPECS[name]
    t, y, dy = gp["gp_t"], gp["gp_y"], gp["gp_dy"]
    P = gp[tag + "_params"]; lp_ref = gp[tag + "_logprob"]; nll_ref = gp[tag + "_negloglike"]
    bounds = [tuple(b) for b in gp[tag + "_bounds"]]
    fit_mean = mean == "constant"
    prior = ogp.log_prior(P, bounds)
    if name == "bendingpl":
        prior = np.where(P[:, 0] < P[:, 1], -np.inf, prior)            # celerite_models.py:85-90
    inside = np.isfinite(prior)
    assert np.array_equal(inside, np.isfinite(lp_ref))
    assert (~inside).sum() >= 1
    ll = np.real(ogp.loglike_params(spec, P[inside], t, y, dy, fit_mean=fit_mean))
    np.testing.assert_allclose(ll, lp_ref[inside], rtol=1e-10)
    np.testing.assert_allclose(-ll, nll_ref[inside], rtol=1e-10)
    lld = ogp.loglike_params(spec, P[inside], t, y, dy, fit_mean=fit_mean, dense=True)
    np.testing.assert_allclose(lld, lp_ref[inside], rtol=1e-12)


def test_psd_identities(gp):
    w = gp["psd_omega"]
    np.testing.assert_allclose(oterms.psd([("drw",)], np.array([-3.0, -8.0]), w), gp["psd_term_drw"], rtol=1e-13)
    np.testing.assert_allclose(oterms.psd([("lorentzian",)], np.array([-4.0, 3.0, -8.5]), w), gp["psd_term_lorentzian"], rtol=1e-13)
    np.testing.assert_allclose(oterms.psd_bendingpowerlaw(w, S0=np.exp(-3.0), omega0=np.exp(-8.0)), gp["psd_model_bendingpowerlaw"], rtol=1e-14)
    np.testing.assert_allclose(oterms.psd_lorentzian(w, S0=np.exp(-4.0), Q=np.exp(3.0), omega0=np.exp(-8.5)), gp["psd_model_lorentzian"], rtol=1e-14)
    np.testing.assert_allclose(oterms.psd_sho(w, S0=0.3, Q=4.0, omega0=2e-3), gp["psd_model_sho"], rtol=1e-14)
    np.testing.assert_allclose(oterms.psd_matern32(w, sigma=0.7, rho=900.0), gp["psd_model_matern32"], rtol=1e-14)
    np.testing.assert_allclose(oterms.psd_matern52(w, sigma=0.7, rho=900.0), gp["psd_model_matern52"], rtol=1e-14)
    # the reference's two DRW PSD forms agree (Term.get_psd of DampedRandomWalk vs psd_models.BendingPowerlaw)
    np.testing.assert_allclose(gp["psd_term_drw"], gp["psd_model_bendingpowerlaw"], rtol=1e-13)
