#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference.

    python tests/golden/make_golden.py            (build container only: needs /root/reference)

The reference's own modules (``/root/reference/mind_the_gaps``: simulator.py, noise_models.py, stats.py,
gpmodelling.py, models/*.py) are imported and executed as they are.  Their un-vendored third-party imports
are satisfied by the stand-ins in ``tests/golden/_shims`` (see its README for exactly what each provides).
Every random draw the reference makes through ``np.random`` / ``scipy.stats`` is recorded next to the
outputs so that the oracle and the CUDA path can be run on identical noise arrays.

Nothing here is imported by tests or the product; the tests only read the ``.npz`` files.
"""
import math
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MTG_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, ".."))
if not hasattr(np, "math"):          # the reference's kraft_pdf uses np.math (stats.py:18), removed in numpy 2
    np.math = math

from mind_the_gaps import simulator as rsim                     # noqa: E402
from mind_the_gaps import noise_models as rnoise                # noqa: E402
from mind_the_gaps import stats as rstats                       # noqa: E402
from mind_the_gaps.models import celerite_models as rterms      # noqa: E402
from mind_the_gaps.models import psd_models as rpsd             # noqa: E402
from mind_the_gaps.models import mean_models as rmean           # noqa: E402
from mind_the_gaps.gpmodelling import GPModelling               # noqa: E402
from mind_the_gaps.lightcurves.gappylightcurve import GappyLightcurve   # noqa: E402
import celerite                                                 # noqa: E402  (the dense stand-in)
from scipy.stats._distn_infrastructure import rv_frozen        # noqa: E402

from helpers import window                                      # noqa: E402  (tests/helpers.py: synthetic windows)


class Recorder:
    """Wraps np.random.{normal,uniform,poisson} and rv_frozen.rvs: same draws, recorded in call order."""
    def __init__(self):
        self.log = []

    def __enter__(self):
        self._orig = {n: getattr(np.random, n) for n in ("normal", "uniform", "poisson")}
        self._rvs = rv_frozen.rvs
        rec = self

        def wrap(name):
            f = self._orig[name]

            def g(*a, **k):
                out = f(*a, **k)
                rec.log.append((name, np.array(out, copy=True), a, k))
                return out
            return g
        for n in self._orig:
            setattr(np.random, n, wrap(n))

        def rvs(self_, *a, **k):
            out = rec._rvs(self_, *a, **k)
            rec.log.append(("rvs", np.array(out, copy=True), a, k))
            return out
        rv_frozen.rvs = rvs
        return self

    def __exit__(self, *exc):
        for n, f in self._orig.items():
            setattr(np.random, n, f)
        rv_frozen.rvs = self._rvs

    def first(self, name):
        for n, out, a, k in self.log:
            if n == name:
                return out, a, k
        raise KeyError(name)


def small_window(n, seed, exposure_lo=40.0, exposure_hi=90.0):
    """irregular gapped window, small enough that N_fft stays ~1e4 (times in s)"""
    rng = np.random.default_rng(seed)
    base = np.arange(0, 4 * n)
    keep = base[(base % 50) < 33]
    picked = np.sort(rng.choice(keep, n, replace=False))
    t = (picked + rng.uniform(-0.3, 0.3, n)) * 400.0
    return t, rng.uniform(exposure_lo, exposure_hi, n)


def gen_simulator(out):
    # ---- grid constants for three windows (Simulator.__init__) ---------------------------------
    cases = []
    for name, (t, e), kw in [
        ("w60", small_window(60, 11), dict(aliasing_factor=2, extension_factor=2)),
        ("w25_scalar", (small_window(25, 12)[0], 55.0), dict(aliasing_factor=1, extension_factor=1.5, epsilon=1.01)),
        ("w1000", window(1000, 3), dict(aliasing_factor=2, extension_factor=2)),
    ]:
        s = rsim.Simulator(rpsd.BendingPowerlaw(S0=1.0, omega0=1e-3), t, e, 1.0, **kw)
        cases.append(name)
        out["grid_%s_times" % name] = t
        out["grid_%s_exposures" % name] = np.asarray(e, dtype=float)
        out["grid_%s_kw" % name] = np.array([kw.get("aliasing_factor", 2), kw.get("extension_factor", 10), kw.get("epsilon", 1.001)], dtype=float)
        out["grid_%s_scalars" % name] = np.array([s.sim_dt, s.sim_duration, s.fftndatapoints, s.sim_timestamps[0], s.sim_timestamps[-1]], dtype=float)
        out["grid_%s_strategy" % name] = np.array(s.strategy, dtype=float)
    out["grid_cases"] = np.array(cases)

    # ---- TK95 chain: get_fft -> simulate_regularly_sampled -> cut_random_segment -> shift -> downsample ----
    t, e = small_window(60, 11)
    S0, w0 = 0.04, 2 * np.pi / (12 * 400.0)
    for tag, pdf, mean in [("tk95", "Gaussian", 2.0), ("e13_lognormal", "Lognormal", 2.0), ("e13_uniform", "Uniform", 2.0)]:
        psd = rpsd.BendingPowerlaw(S0=S0, omega0=w0)
        sim = rsim.Simulator(psd, t, e, mean, pdf=pdf, aliasing_factor=2, extension_factor=2, max_iter=400)
        np.random.seed(1234)
        with Recorder() as rec, warnings.catch_warnings():
            warnings.simplefilter("ignore")
            lc = sim.simulate_regularly_sampled()
            seg = rsim.cut_random_segment(lc, sim.sim_duration)
            shifted = seg.shift(-seg.tstart + sim.strategy[0][0])
            seg_rates_in = np.array(shifted.countrate, copy=True)
            adj = sim.simulator.adjust_pdf(shifted)
            rates = np.array(sim.downsample(adj), dtype=float)
        normals, _, _ = rec.first("normal")
        uni, ua, _ = rec.first("uniform")
        out[tag + "_times"] = t; out[tag + "_exposures"] = e
        out[tag + "_psd_S0_w0_mean"] = np.array([S0, w0, mean])
        out[tag + "_normals"] = normals                               # (2, N_fft//2+1)
        out[tag + "_uniform"] = np.array([float(uni), float(ua[0]), float(ua[1])])     # draw, low, high
        out[tag + "_regular_countrate"] = np.array(lc.countrate)
        out[tag + "_segment_time"] = np.array(shifted.time)
        out[tag + "_segment_countrate_tk95"] = seg_rates_in
        out[tag + "_segment_countrate"] = np.array(adj.countrate)
        out[tag + "_rates"] = rates
        if pdf != "Gaussian":
            draws, _, _ = rec.first("rvs")
            out[tag + "_pdf_draws"] = draws
            out[tag + "_sample_std"] = np.array([np.std(seg_rates_in)])
    # get_fft alone, odd and even N
    for N in (257, 1024):
        np.random.seed(77)
        with Recorder() as rec:
            cf = rsim.get_fft(N, 3.5, rpsd.Lorentzian(S0=2.0, Q=12.0, omega0=0.2))
        out["getfft_%d_normals" % N] = rec.first("normal")[0]
        out["getfft_%d_out" % N] = cf

    # ---- the reference's own downsampling test cases (tests/simulator_test.py:192-253) -------------
    for exposures in (0.5, 0.6, 0.1):
        timestamps = np.append(np.arange(1, 3.1, 2), np.arange(5, 7.1, 2))
        sim = rsim.Simulator(rpsd.BendingPowerlaw(), timestamps, exposures, 1.0, aliasing_factor=1, extension_factor=1.0)
        times = np.arange(0.5, 10.1, 0.1)
        rates_in = np.linspace(5, 20, len(times)) / exposures
        lc = rsim.Lightcurve(times, rates_in, input_counts=False, dt=0.1, skip_checks=True)
        out["downsample_%g_rates" % exposures] = np.array(sim.downsample(lc), dtype=float)
        out["downsample_%g_in" % exposures] = rates_in


def gen_noise(out):
    n = 64
    rng = np.random.default_rng(5)
    exposures = rng.uniform(500.0, 2000.0, n)
    # Poisson (noise_models.py:49-78)
    rates = rng.uniform(0.3, 0.8, n)
    np.random.seed(42)
    with Recorder() as rec:
        noisy, dy = rnoise.PoissonNoise(exposures).add_noise(rates)
    out["poisson_exposures"] = exposures; out["poisson_rates"] = rates
    out["poisson_counts"] = rec.first("poisson")[0].astype(np.float64)
    out["poisson_noisy"] = noisy; out["poisson_dy"] = dy
    # Gaussian (noise_models.py:162-184)
    np.random.seed(43)
    with Recorder() as rec:
        noisy, dy = rnoise.GaussianNoise(exposures, 0.07).add_noise(rates)
    out["gaussian_sigma"] = np.array([0.07]); out["gaussian_normals"] = rec.first("normal")[0] / 0.07
    out["gaussian_noisy"] = noisy; out["gaussian_dy"] = dy
    # Kraft (noise_models.py:107-150): low count rates straddling kraft_counts = 15
    rates = rng.uniform(0.004, 0.018, n)
    bkg_rate = np.full(n, 0.001); bkg_rate_err = np.full(n, 1e-4)
    np.random.seed(44)
    with Recorder() as rec, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        noisy, dy = rnoise.KraftNoise(exposures, bkg_rate * exposures, bkg_rate_err).add_noise(rates)
    out["kraft_exposures"] = exposures; out["kraft_rates"] = rates
    out["kraft_bkg_counts"] = bkg_rate * exposures; out["kraft_bkg_rate_err"] = bkg_rate_err
    out["kraft_counts"] = rec.first("poisson")[0].astype(np.float64)
    out["kraft_noisy"] = noisy                      # reference code: scipy numerical median of the reference's kraft_pdf
    out["kraft_dy"] = dy                            # low-count bins depend on the poisson_conf_interval STAND-IN
    # medians of the reference's kraft_pdf through scipy's rv_continuous machinery, as noise_models.py:139-141
    pdf = rstats.kraft_pdf(a=0, b=35)
    NB = [(0, 0.5), (1, 0.0), (3, 1.7), (7, 9.2), (14, 2.0), (10, 0.9)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["kraft_median_NB"] = np.array(NB, dtype=float)
        out["kraft_median"] = np.array([pdf(N, B).median() for N, B in NB])
    # target-PDF factories (stats.py:116-146)
    ms = [(10.0, 2.0), (0.5, 0.3), (2.0, 0.21)]
    out["pdf_factory_mean_std"] = np.array(ms)
    out["lognormal_s_scale"] = np.array([[rstats.create_log_normal(m, s).args[0], rstats.create_log_normal(m, s).kwds["scale"]] for m, s in ms])
    out["uniform_loc_scale"] = np.array([[rstats.create_uniform_distribution(m, s).kwds["loc"], rstats.create_uniform_distribution(m, s).kwds["scale"]] for m, s in ms])


def lc_for_gp(n, seed):
    rng = np.random.default_rng(seed)
    t, e = small_window(n, seed)
    y = 2.0 + 0.3 * np.sin(t / 9000.0) + rng.normal(0, 0.15, n)
    dy = rng.uniform(0.05, 0.12, n)
    return t, y, dy, e


def gen_gp(out):
    """GPModelling glue (reference code) on the dense-Cholesky celerite stand-in."""
    t, y, dy, e = lc_for_gp(80, 21)
    out["gp_t"] = t; out["gp_y"] = y; out["gp_dy"] = dy; out["gp_exposures"] = e
    lc = GappyLightcurve(t, y, dy, e)
    rng = np.random.default_rng(22)
    b_drw = dict(log_S0=(-10.0, 5.0), log_omega0=(-15.0, -3.0))
    b_lor = dict(log_S0=(-12.0, 5.0), log_Q=(0.5, 7.0), log_omega0=(-12.0, -5.0))
    b_bpl = dict(log_S0=(-12.0, 5.0), log_Q=(-14.0, 5.0), log_omega0=(-15.0, -3.0))
    builders = {
        "drw": lambda: rterms.DampedRandomWalk(-3.0, -8.0, bounds=b_drw),
        "drw_lorentzian": lambda: rterms.DampedRandomWalk(-3.0, -8.0, bounds=b_drw) + rterms.Lorentzian(-4.0, 3.0, -8.5, bounds=b_lor),
        "lorentzian_drw": lambda: rterms.Lorentzian(-4.0, 3.0, -8.5, bounds=b_lor) + rterms.DampedRandomWalk(-3.0, -8.0, bounds=b_drw),
        "cosinus_drw": lambda: rterms.Cosinus(-4.0, -8.5, bounds=dict(log_S0=(-12, 5), log_omega0=(-12, -5))) + rterms.DampedRandomWalk(-3.0, -8.0, bounds=b_drw),
        "bendingpl": lambda: rterms.BendingPowerlaw(-3.0, -5.0, -8.0, bounds=b_bpl),
        "sho_real": lambda: celerite.terms.SHOTerm(-3.0, 1.0, -8.0, bounds=dict(log_S0=(-12, 5), log_Q=(-3, 5), log_omega0=(-12, -5)))
        + celerite.terms.RealTerm(-3.5, -9.0, bounds=dict(log_a=(-12, 5), log_c=(-14, -4))),
        "matern_jitter": lambda: celerite.terms.Matern32Term(-2.0, 9.0, bounds=dict(log_sigma=(-8, 3), log_rho=(5, 12)))
        + celerite.terms.JitterTerm(-3.0, bounds=dict(log_sigma=(-8, 0))),
    }
    names = []
    for name, build in builders.items():
        for mean_model in (None, "constant", "linear"):
            if mean_model is not None and name not in ("drw", "drw_lorentzian"):
                continue
            tag = "gp_%s_%s" % (name, mean_model or "none")
            gpm = GPModelling(lc, build(), mean_model=mean_model)
            p0 = np.array(gpm.initial_params, dtype=float)
            nd = len(p0)
            P = p0[None, :] + rng.normal(0, 0.4, (12, nd))
            P[0] = p0
            bounds = gpm.gp.get_parameter_bounds()
            # a few rows outside the prior box, to pin the -inf branch of _log_probability
            P[10, 0] = (bounds[0][1] if bounds[0][1] is not None and np.isfinite(bounds[0][1]) else 1e3) + 1.0
            if name == "bendingpl":
                P[11, 1] = P[11, 0] + 0.5           # log_Q > log_S0: BendingPowerlaw.log_prior override
            lp = np.array([gpm._log_probability(p) for p in P])
            nll = np.array([gpm._neg_log_like(p) if np.isfinite(l) else np.nan for p, l in zip(P, lp)])
            out[tag + "_params"] = P
            out[tag + "_logprob"] = lp
            out[tag + "_negloglike"] = nll
            out[tag + "_bounds"] = np.array([[-np.inf if b[0] is None else b[0], np.inf if b[1] is None else b[1]] for b in bounds], dtype=float)
            out[tag + "_names"] = np.array(gpm.gp.get_parameter_names())
            names.append(tag)
    out["gp_cases"] = np.array(names)
    # kernel PSDs through Term.get_psd of the reference Terms (celerite normalisation) and psd_models.py
    w = np.geomspace(1e-6, 1e-1, 40)
    out["psd_omega"] = w
    out["psd_term_drw"] = rterms.DampedRandomWalk(-3.0, -8.0).get_psd(w)
    out["psd_term_lorentzian"] = rterms.Lorentzian(-4.0, 3.0, -8.5).get_psd(w)
    out["psd_model_bendingpowerlaw"] = rpsd.BendingPowerlaw(S0=math.exp(-3.0), omega0=math.exp(-8.0))(w)
    out["psd_model_lorentzian"] = rpsd.Lorentzian(S0=math.exp(-4.0), Q=math.exp(3.0), omega0=math.exp(-8.5))(w)
    out["psd_model_sho"] = rpsd.SHO(S0=0.3, Q=4.0, omega0=2e-3)(w)
    out["psd_model_matern32"] = rpsd.Matern32(sigma=0.7, rho=900.0)(w)
    out["psd_model_matern52"] = rpsd.Matern52(sigma=0.7, rho=900.0)(w)
    out["psd_model_jitter"] = rpsd.Jitter(sigma=0.3)(w)
    # spread_walkers (gpmodelling.py:289-350): deterministic given np.random.seed -- pin shape / box invariants only
    gpm = GPModelling(lc, builders["drw_lorentzian"]())
    np.random.seed(9)
    bounds = np.array(gpm.gp.get_parameter_bounds(), dtype=float)
    try:
        w0 = gpm.spread_walkers(32, gpm.initial_params, bounds, 0.1)
        out["spread_walkers_out"] = np.asarray(w0)
        out["spread_walkers_bounds"] = bounds
        out["spread_walkers_p0"] = np.asarray(gpm.initial_params)
    except Exception as exc:        # signature drift: not fatal for the rest
        print("spread_walkers fixture skipped:", exc)
    # mean models (mean_models.py)
    x = np.linspace(0.0, 5e4, 7)
    out["mean_x"] = x
    out["mean_gaussian"] = rmean.GaussianModel(2e4, 6e3, 3e4, 0.5).get_value(x)
    out["mean_linear"] = rmean.LinearModel(1e-5, 1.5).get_value(x)


def write_io_inputs(folder):
    """Synthetic input files in the three on-disk formats the reference reads (tests/golden/io/)."""
    os.makedirs(folder, exist_ok=True)
    rng = np.random.default_rng(42)
    n = 40
    # --- Swift-XRT QDP, observation binning (16 columns, trailing ObsID) and snapshot binning (15 columns), MJD and seconds
    cols = ["T_+ve", "T_-ve", "Rate", "Ratepos", "Rateneg", "FracExp", "BGrate", "BGerr", "CorrFact", "CtsInSrc", "BGInSrc",
            "Exposure", "Sigma", "SNR"]
    for fname, tname, obsid in (("PCCURVE_obsid_mjd.qdp", "MJD", True), ("PCCURVE_snapshot_sec.qdp", "Time", False)):
        t = np.sort(rng.uniform(0, 400, n)) + (58000.0 if tname == "MJD" else 0.0)
        if tname != "MJD":
            t = t * 86400.0
        half = rng.uniform(0.002, 0.01, n) if tname == "MJD" else rng.uniform(200, 900, n)
        rate = rng.uniform(0.01, 0.6, n)
        err = rate * rng.uniform(0.05, 0.3, n)
        expo = rng.uniform(300, 1800, n)
        rows = []
        for i in range(n):
            vals = [t[i], half[i], -half[i], rate[i], err[i] * 1.1, -err[i] * 0.9, rng.uniform(0.5, 1.0), rng.uniform(1e-4, 5e-3),
                    rng.uniform(1e-5, 5e-4), rng.uniform(1.0, 2.5), int(rng.integers(0, 400)), rng.uniform(0, 4), expo[i],
                    rng.uniform(0.5, 40), rng.uniform(0.5, 25)]
            txt = "\t".join(("%d" % v) if isinstance(v, int) else ("%.9g" % v) for v in vals)
            if obsid:
                txt += "\t000%08d" % (31000000 + i)
            rows.append(txt)
        with open(os.path.join(folder, fname), "w") as fh:
            fh.write("READ TERR 1 2\n! PC data\n")
            fh.write("!" + "\t".join([tname] + cols + (["ObsID"] if obsid else [])) + "\n")
            fh.write("\n".join(rows[: n // 2]) + "\n")
            fh.write("! a comment line between observations\n")
            fh.write("\n".join(rows[n // 2:]) + "\n")
    # --- the 6-column tab format GappyLightcurve.to_csv writes, a 3-column csv in days, a 4-column file
    t = np.sort(rng.uniform(0, 3e7, n)); y = rng.uniform(0.1, 2, n); dy = y * 0.1; ex = rng.uniform(300, 1500, n)
    GappyLightcurve(t, y, dy, ex, rng.uniform(0, 0.01, n), rng.uniform(0, 0.001, n)).to_csv(os.path.join(folder, "simple_6col.dat"))
    np.savetxt(os.path.join(folder, "simple_3col_days.csv"), np.c_[t / 86400.0 + 55000, y, dy], delimiter=",", header="mjd,flux,err",
               comments="", fmt="%.10f")
    np.savetxt(os.path.join(folder, "simple_4col.txt"), np.c_[t, y, dy, ex], header="t rate err exposure", fmt="%.8e")
    # --- Fermi-LAT csv
    flux = rng.uniform(1e-8, 5e-7, n)
    np.savetxt(os.path.join(folder, "fermi.csv"), np.c_[t, flux, -flux * rng.uniform(0.1, 0.3, n), flux * rng.uniform(0.1, 0.3, n), rng.uniform(4, 90, n)],
               delimiter=",", header="time,flux,flux_err_neg,flux_err_pos,ts", comments="", fmt="%.9e")


def gen_io(out):
    """The reference's own file readers (lightcurves/{simple,swift,fermi}lightcurve.py) and to_csv on tests/golden/io/*."""
    import contextlib
    import io as _io
    from mind_the_gaps.lightcurves.simplelightcurve import SimpleLightcurve
    from mind_the_gaps.lightcurves.swiftlightcurve import SwiftLightcurve
    from mind_the_gaps.lightcurves.fermilightcurve import FermiLightcurve
    folder = os.path.join(HERE, "io")
    write_io_inputs(folder)

    def store(tag, lc):
        for f in ("times", "y", "dy", "exposures", "bkg_rate", "bkg_rate_err"):
            out["%s_%s" % (tag, f)] = np.asarray(getattr(lc, f), dtype=float)

    with contextlib.redirect_stdout(_io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        store("swift_obsid", SwiftLightcurve(os.path.join(folder, "PCCURVE_obsid_mjd.qdp")))
        store("swift_obsid_cut", SwiftLightcurve(os.path.join(folder, "PCCURVE_obsid_mjd.qdp"), minSNR=5, minSigma=3, minCts=50))
        store("swift_snapshot", SwiftLightcurve(os.path.join(folder, "PCCURVE_snapshot_sec.qdp")))
        store("simple6", SimpleLightcurve(os.path.join(folder, "simple_6col.dat")))
        store("simple3", SimpleLightcurve(os.path.join(folder, "simple_3col_days.csv"), delimiter=","))
        store("simple4", SimpleLightcurve(os.path.join(folder, "simple_4col.txt")))
        try:
            store("fermi", FermiLightcurve(os.path.join(folder, "fermi.csv")))
        except TypeError as exc:
            # numpy >= 2 refuses `"MJD" in structured_array` (fermilightcurve.py:37); numpy 1.x evaluated it to False.
            # Run the rest of readdata by hand with that branch taken: the remaining lines are the reference's own.
            print("FermiLightcurve raises on this numpy (%s); fixture built from its remaining statements" % exc)
            data = np.genfromtxt(os.path.join(folder, "fermi.csv"), names=True, delimiter=",")
            rate_column = data.dtype.names[1]
            out["fermi_times"] = np.asarray(data[data.dtype.names[0]], dtype=float)
            out["fermi_y"] = np.asarray(data[rate_column], dtype=float)
            out["fermi_dy"] = (np.abs(data["%s_err_neg" % rate_column]) + data["%s_err_pos" % rate_column]) / 2
        lc = SimpleLightcurve(os.path.join(folder, "simple_6col.dat"))
    # truncate / split / to_csv of the container itself
    tr = lc.truncate(lc.times[5], lc.times[20])
    out["truncate_times"] = tr.times
    parts = lc.split(2.0e6)
    out["split_sizes"] = np.array([p.n for p in parts])
    out["split_first_times"] = np.array([p.times[0] for p in parts])
    tmp = os.path.join(folder, "_roundtrip.tmp")
    lc.to_csv(tmp)
    out["to_csv_text"] = np.frombuffer(open(tmp, "rb").read(), dtype=np.uint8)
    os.remove(tmp)


def gen_fitting(out):
    """The lmfit-free functions of the reference's fitting.py (s_statistic, s_stat_powerlaw, minimize_powerlaw, brokenpl)
    on a synthetic periodogram; the lmfit-based ones cannot run here (lmfit is a placeholder)."""
    from mind_the_gaps import fitting as rfit
    rng = np.random.default_rng(3)
    f = np.fft.rfftfreq(2000, 0.5)[1:-1]
    model = 3.0 * f ** -1.3
    p = model * rng.exponential(1.0, len(f))                    # chi^2_2-distributed periodogram around a power law
    out["f"] = f; out["p"] = p; out["model"] = model
    out["s_statistic"] = np.float64(rfit.s_statistic(p, model))
    out["s_stat_powerlaw"] = np.float64(rfit.s_stat_powerlaw([2.5, -1.1], f, p))
    out["minimize_powerlaw"] = np.asarray(rfit.minimize_powerlaw(f, p))
    x = np.linspace(-3.0, 2.0, 23)
    out["brokenpl_x"] = x
    out["brokenpl"] = rfit.brokenpl(x, 1.5, -2.0, 0.25)


def main():
    for name, fn in (("simulator", gen_simulator), ("noise", gen_noise), ("gp", gen_gp), ("io", gen_io), ("fitting", gen_fitting)):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        out = {}
        fn(out)
        path = os.path.join(HERE, "ref_%s.npz" % name)
        np.savez_compressed(path, **out)
        print("%s: %d arrays, %.1f kB" % (path, len(out), os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()
