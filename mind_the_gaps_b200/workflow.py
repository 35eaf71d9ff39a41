"""The command-line workflow of ``docs/workflow.md`` (reference: five scripts -- ``celerite_script.py``,
``generate_lcs_significance.py``, ``fit_lcs.py``, ``plot_ratio_test.py`` -- that the documentation describes but the
tree does not contain).  Same steps, same on-disk artefacts (``best_fit.dat``, ``samples.dat``,
``lightcurves/*.dat``, ``fits_results.dat``), on top of ``GPModelling`` / ``lrt``:

    python -m mind_the_gaps_b200.workflow fit input.dat -c null.config -o out [--fit] [--tmin T --tmax T]
    python -m mind_the_gaps_b200.workflow generate --input_dir <fit dir> -f input.dat -n 5000 --pdf Gaussian
    python -m mind_the_gaps_b200.workflow fit-lcs <dir>/lightcurves/*.dat -c alt.config -o sims
    python -m mind_the_gaps_b200.workflow ratio-test -n <null sims>/fits_results.dat -a <alt sims>/fits_results.dat \\
                                                     -dn <null data>/fits_results.dat -da <alt data>/fits_results.dat
    python -m mind_the_gaps_b200.workflow ppp input.dat -c null.config -a alt.config -n 10000        (all of it, batched)

Lightcurve files are the tab-separated tables of the workflow (time, rate, error, exposure[, background rate and error]):
``lightcurves.SimpleLightcurve`` reads them, ``GappyLightcurve.to_csv`` writes them.  The celerite config files of the
reference are not in its tree either; the format read here is one kernel component per line,

    # component      parameter = initial : lower : upper   ...
    DampedRandomWalk log_S0=-3:-10:5  log_omega0=-13:-20:-5
    Lorentzian       log_S0=-5:-15:5  log_Q=3:0.5:7  log_omega0=-12:-16:-9

with the component names of ``mind_the_gaps.models`` and ``celerite.terms`` and an optional ``mean = constant|linear|gaussian``
line.  The numerical work is the GPU path; this module is file handling and orchestration."""
import argparse
import glob
import os
import sys

import numpy as np

from .celerite import terms as _cterms
from .models import celerite_models as _mterms

__all__ = ["COMPONENTS", "parse_config", "model_label", "write_table", "read_table", "fit_lightcurve", "generate_lightcurves",
           "fit_lightcurves", "ratio_test", "main"]

COMPONENTS = {
    "DampedRandomWalk": _mterms.DampedRandomWalk, "Lorentzian": _mterms.Lorentzian, "Cosinus": _mterms.Cosinus,
    "BendingPowerlaw": _mterms.BendingPowerlaw, "RealTerm": _cterms.RealTerm, "ComplexTerm": _cterms.ComplexTerm,
    "SHOTerm": _cterms.SHOTerm, "Matern32Term": _cterms.Matern32Term, "JitterTerm": _cterms.JitterTerm,
}


def parse_config(path_or_lines):
    """``(kernel, mean_model, component_names)`` from a model config (see the module docstring)."""
    lines = open(path_or_lines).read().splitlines() if isinstance(path_or_lines, str) else list(path_or_lines)
    kernel, names, mean = None, [], None
    for raw in lines:
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.lower().startswith("mean"):
            mean = line.split("=", 1)[1].strip().lower()
            continue
        parts = line.replace(" = ", "=").replace("= ", "=").replace(" =", "=").split()
        comp = parts[0]
        if comp not in COMPONENTS:
            raise ValueError("unknown model component '%s' (known: %s)" % (comp, ", ".join(sorted(COMPONENTS))))
        cls = COMPONENTS[comp]
        values, bounds = {}, {}
        for item in parts[1:]:
            if "=" not in item:
                raise ValueError("cannot read '%s' in line '%s' (expected parameter=initial:lower:upper)" % (item, raw))
            key, spec = item.split("=", 1)
            fields = spec.split(":")
            values[key] = float(fields[0])
            if len(fields) == 3:
                bounds[key] = (float(fields[1]), float(fields[2]))
            elif len(fields) != 1:
                raise ValueError("parameter '%s': give initial or initial:lower:upper" % key)
        missing = [n for n in cls.parameter_names if n not in values]
        if missing or len(values) != len(cls.parameter_names):
            raise ValueError("%s needs the parameters %s" % (comp, ", ".join(cls.parameter_names)))
        term = cls(*[values[n] for n in cls.parameter_names], bounds=bounds or None)
        kernel = term if kernel is None else kernel + term
        names.append(comp)
    if kernel is None:
        raise ValueError("the config defines no kernel component")
    return kernel, mean, names


def model_label(names):
    return "_".join(names)


def write_table(path, columns, rows, header_extra=None):
    """Tab-separated table: optional ``#`` comment line, a header line of column names, one row per record."""
    rows = np.atleast_2d(np.asarray(rows, dtype=np.float64))
    with open(path, "w") as fh:
        if header_extra:
            fh.write("# %s\n" % header_extra)
        fh.write("\t".join(columns) + "\n")
        for r in rows:
            fh.write("\t".join("%.10e" % v for v in r) + "\n")


def read_table(path):
    """``(column names, rows)`` of a table written by ``write_table`` (``#`` lines are comments)."""
    lines = [ln for ln in open(path).read().splitlines() if ln.strip() and not ln.lstrip().startswith("#")]
    names = tuple(lines[0].split("\t"))
    rows = np.array([[float(v) for v in ln.split("\t")] for ln in lines[1:]], dtype=np.float64).reshape(len(lines) - 1, len(names))
    return names, rows


def _load_lightcurve(path, tmin=None, tmax=None):
    from .lightcurves import SimpleLightcurve
    lc = SimpleLightcurve(path)
    if tmin is not None or tmax is not None:
        lc = lc.truncate(-np.inf if tmin is None else tmin, np.inf if tmax is None else tmax)
    return lc


def _short(name):
    return name.replace("kernel:", "").replace("terms[", "t").replace("]:", "_").replace("mean:", "mean_")


def fit_lightcurve(lc_file, config, outname="output", fit=True, tmin=None, tmax=None, max_steps=50000, walkers=12, outdir=".",
                   modelling=None):
    """Step 1 (``celerite_script.py``): posteriors of one model on the input lightcurve -> ``best_fit.dat``, ``samples.dat``.
    Returns the directory written."""
    kernel, mean, names = parse_config(config)
    lc = _load_lightcurve(lc_file, tmin, tmax)
    if modelling is None:
        from .gpmodelling import GPModelling as modelling
    gpm = modelling(lc, kernel, mean)
    gpm.derive_posteriors(fit=fit, max_steps=max_steps, walkers=walkers, progress=False)
    d = os.path.join(outdir, "celerite_%s_t%d-%d_m_%s%s" % (outname, int(lc.times[0]), int(lc.times[-1]), model_label(names), "_fit" if fit else ""))
    os.makedirs(d, exist_ok=True)
    cols = [_short(n) for n in gpm.parameter_names]
    write_table(os.path.join(d, "best_fit.dat"), cols + ["loglikelihood"], np.r_[gpm.max_parameters, gpm.max_loglikelihood],
                header_extra="maximum of the posterior samples; converged=%s" % getattr(gpm, "converged", "?"))
    write_table(os.path.join(d, "samples.dat"), cols + ["loglikelihood"], np.c_[gpm.mcmc_samples, gpm.loglikelihoods])
    return d


def generate_lightcurves(input_dir, lc_file, config, n=5000, pdf="Gaussian", outdir=".", extension_factor=2, simulate=None):
    """Step 2 (``generate_lcs_significance.py``): ``n`` lightcurves from posterior samples of the null model ->
    ``<dir>/lightcurves/lc<i>.dat`` in the input format.  ``simulate(lc, kernel, samples, pdf, extension_factor)`` returns
    ``(rates (n, N), errors (n, N))``; the default is one batched device call."""
    kernel, mean, names = parse_config(config)
    lc = _load_lightcurve(lc_file)
    cols, samples = read_table(os.path.join(input_dir, "samples.dat"))
    nk = len(kernel.get_parameter_vector())
    draws = samples[np.random.randint(len(samples), size=int(n)), :nk]
    if simulate is None:
        def simulate(lc_, kernel_, draws_, pdf_, ext_):
            sim = lc_.get_simulator(kernel_.get_psd, pdf_, extension_factor=ext_)
            try:
                kernel_.set_parameter_vector(draws_[0])
                rates = sim.generate_lightcurves(len(draws_), parameters=draws_)
                return sim.add_noise(rates)
            finally:
                sim.close()
    rates, errors = simulate(lc, kernel, draws, pdf, extension_factor)
    d = os.path.join(outdir, "lightcurves_significance_t%d-%ds_n%d_pdf%s" % (int(lc.times[0]), int(lc.times[-1]), int(n), pdf))
    os.makedirs(os.path.join(d, "lightcurves"), exist_ok=True)
    from .lightcurves import GappyLightcurve
    for i in range(len(rates)):
        GappyLightcurve(lc.times, rates[i], errors[i], lc.exposures, lc.bkg_rate, lc.bkg_rate_err).to_csv(
            os.path.join(d, "lightcurves", "lc%d.dat" % i))
    write_table(os.path.join(d, "parameters.dat"), [_short(c) for c in cols[:nk]], draws)
    return d


def fit_lightcurves(files, config, outname="sims", outdir=".", fitter=None):
    """Steps 3-4 (``fit_lcs.py``): maximum log-likelihood of one model for every lightcurve file -> ``fits_results.dat``
    (one row per file: log-likelihood, then the parameters).  Files sharing their timestamps are fitted in ONE batched
    device call (``fitting.fit_lightcurves``); ``fitter(lcs, kernel) -> (loglike (L,), params (L, ndim))`` overrides it."""
    kernel, mean, names = parse_config(config)
    files = sorted(files)
    lcs = [_load_lightcurve(f) for f in files]
    if fitter is None:
        def fitter(lcs_, kernel_):
            from . import fitting
            from .celerite import GP
            t0 = lcs_[0].times
            if not all(len(l.times) == len(t0) and np.array_equal(l.times, t0) for l in lcs_):
                raise ValueError("fit-lcs: the lightcurves must share their timestamps (simulations of one window)")
            gp = GP(kernel_, mean=0.0)
            res = fitting.fit_lightcurves(gp, t0, np.array([l.y for l in lcs_]), np.array([l.dy for l in lcs_]), n_starts=4)
            return -np.array([r.fun for r in res]), np.array([r.x for r in res])
    loglike, params = fitter(lcs, kernel)
    d = os.path.join(outdir, "fit_lcs_%s_m_%s" % (outname, model_label(names)))
    os.makedirs(d, exist_ok=True)
    cols = ["loglikelihood"] + [_short(n) for n in kernel.get_parameter_names()]
    write_table(os.path.join(d, "fits_results.dat"), cols, np.c_[loglike, params], header_extra="files: %s" % " ".join(os.path.basename(f) for f in files))
    return d


def ratio_test(null_sims, alt_sims, null_data, alt_data):
    """Step 5 (``plot_ratio_test.py``): ``T = -2 (L_null - L_alt)`` of the simulations against that of the data; returns
    ``dict(T_dist, T_obs, p_value, percentile_95, percentile_997)`` (the notebook's arithmetic, ``tutorial_ppp.ipynb`` cell 15)."""
    from .stats import lrt_pvalue, lrt_statistic
    ln = read_table(null_sims)[1][:, 0]; la = read_table(alt_sims)[1][:, 0]
    if len(ln) != len(la):
        raise ValueError("the two simulation tables hold different numbers of lightcurves (%d vs %d)" % (len(ln), len(la)))
    T_dist = lrt_statistic(ln, la)
    T_obs = float(lrt_statistic(read_table(null_data)[1][0, 0], read_table(alt_data)[1][0, 0]))
    return dict(T_dist=T_dist, T_obs=T_obs, p_value=float(lrt_pvalue(T_dist, T_obs)), percentile_95=float(np.percentile(T_dist, 95)),
                percentile_997=float(np.percentile(T_dist, 99.7)))


def _ppp(args):
    """All five steps for one lightcurve with the batched LRT driver (no per-lightcurve files unless asked for)."""
    from . import lrt
    from .gpmodelling import GPModelling
    null_kernel, mean, null_names = parse_config(args.config)
    alt_kernel, _, alt_names = parse_config(args.alternative)
    lc = _load_lightcurve(args.lightcurve, args.tmin, args.tmax)
    null_model = GPModelling(lc, null_kernel, mean); null_model.derive_posteriors(fit=True, max_steps=args.max_steps, progress=False)
    alt_model = GPModelling(lc, alt_kernel, mean); alt_model.derive_posteriors(fit=True, max_steps=args.max_steps, progress=False)
    nk = len(null_kernel.get_parameter_vector())
    draws = null_model.mcmc_samples[np.random.randint(len(null_model.mcmc_samples), size=args.n), :nk]
    res = lrt.simulate_and_fit(lc, null_kernel, alt_kernel, draws, pdf=args.pdf, refine_walkers=args.refine_walkers, refine_steps=args.refine_steps)
    T_obs = -2 * (null_model.max_loglikelihood - alt_model.max_loglikelihood)
    os.makedirs(args.outdir, exist_ok=True)
    res.save(os.path.join(args.outdir, "lrt_%s_vs_%s.npz" % (model_label(null_names), model_label(alt_names))))
    print("Observed LRT_stat: %.3f" % T_obs)
    print("p-value: %.4f  (%d simulations; 95 %%: %.2f, 99.7 %%: %.2f)" % (res.pvalue(T_obs), args.n, np.percentile(res.T, 95), np.percentile(res.T, 99.7)))
    return res, T_obs


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m mind_the_gaps_b200.workflow", description=__doc__.split("\n\n")[0])
    sub = ap.add_subparsers(dest="command", required=True)
    p = sub.add_parser("fit"); p.add_argument("lightcurve"); p.add_argument("-c", "--config", required=True); p.add_argument("-o", "--outname", default="output")
    p.add_argument("--fit", action="store_true"); p.add_argument("--tmin", type=float); p.add_argument("--tmax", type=float)
    p.add_argument("--max-steps", type=int, default=50000); p.add_argument("--walkers", type=int, default=12); p.add_argument("--outdir", default=".")
    p = sub.add_parser("generate"); p.add_argument("--input_dir", required=True); p.add_argument("-f", "--file", required=True)
    p.add_argument("-c", "--config", required=True); p.add_argument("-n", type=int, default=5000); p.add_argument("--pdf", default="Gaussian"); p.add_argument("--outdir", default=".")
    p = sub.add_parser("fit-lcs"); p.add_argument("files", nargs="+"); p.add_argument("-c", "--config", required=True); p.add_argument("-o", "--outname", default="sims")
    p.add_argument("--outdir", default="."); p.add_argument("--cores", type=int, default=1, help="accepted for compatibility; the fits are one batched device call")
    p = sub.add_parser("ratio-test"); p.add_argument("-n", required=True); p.add_argument("-a", required=True); p.add_argument("-dn", required=True); p.add_argument("-da", required=True)
    p = sub.add_parser("ppp"); p.add_argument("lightcurve"); p.add_argument("-c", "--config", required=True); p.add_argument("-a", "--alternative", required=True)
    p.add_argument("-n", type=int, default=10000); p.add_argument("--pdf", default="Gaussian"); p.add_argument("--max-steps", type=int, default=50000)
    p.add_argument("--refine-walkers", type=int, default=30); p.add_argument("--refine-steps", type=int, default=500)
    p.add_argument("--tmin", type=float); p.add_argument("--tmax", type=float); p.add_argument("--outdir", default=".")
    args = ap.parse_args(argv)
    if args.command == "fit":
        print(fit_lightcurve(args.lightcurve, args.config, args.outname, args.fit, args.tmin, args.tmax, args.max_steps, args.walkers, args.outdir))
    elif args.command == "generate":
        print(generate_lightcurves(args.input_dir, args.file, args.config, args.n, args.pdf, args.outdir))
    elif args.command == "fit-lcs":
        files = [f for pat in args.files for f in (glob.glob(pat) or [pat])]
        print(fit_lightcurves(files, args.config, args.outname, args.outdir))
    elif args.command == "ratio-test":
        out = ratio_test(args.n, args.a, args.dn, args.da)
        print("Observed LRT_stat: %.3f" % out["T_obs"])
        print("p-value: %.4f  (%d simulations; 95 %%: %.2f, 99.7 %%: %.2f)" % (out["p_value"], len(out["T_dist"]), out["percentile_95"], out["percentile_997"]))
    else:
        _ppp(args)
    return 0


if __name__ == "__main__":
    sys.exit(main())
