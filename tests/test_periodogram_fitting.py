"""The reference's periodogram-fitting helpers (``mind_the_gaps/fitting.py``), restated without lmfit: the numpy / scipy
ones are pinned on outputs of the reference itself (``tests/golden/ref_fitting.npz``), the least-squares ones on closed
forms and on recovering a known power law."""
import os

import numpy as np
import pytest

from mind_the_gaps_b200 import fitting

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(HERE, "golden", "ref_fitting.npz"))


def test_whittle_statistic_and_powerlaw_minimisation_match_the_reference(ref):
    f, p, model = ref["f"], ref["p"], ref["model"]
    assert fitting.s_statistic(p, model) == ref["s_statistic"]
    assert fitting.s_stat_powerlaw([2.5, -1.1], f, p) == ref["s_stat_powerlaw"]
    np.testing.assert_allclose(fitting.minimize_powerlaw(f, p), ref["minimize_powerlaw"], rtol=1e-6)
    np.testing.assert_array_equal(fitting.brokenpl(ref["brokenpl_x"], 1.5, -2.0, 0.25), ref["brokenpl"])


def test_fit_psd_powerlaw_vaughan_2005(ref):
    f, p = ref["f"], ref["p"]
    slope, slope_err, norm, norm_err = fitting.fit_psd_powerlaw(f, p)
    # unconstrained optimum inside the bounds = ordinary least squares in log10-log10
    m, c = np.polyfit(np.log10(f), np.log10(p), 1)
    assert abs(slope.value - m) < 1e-8 and abs(float(slope) - m) < 1e-8
    assert abs(norm - 10 ** (c + 0.25068)) < 1e-6 * norm
    # analytic uncertainties of Vaughan (2005), eq. for a least-squares line with variance pi^2 / (6 ln(10)^2)
    lf = np.log10(f); n = len(f); s2 = np.pi ** 2 / (6 * np.log(10) ** 2)
    delta = n * np.sum(lf ** 2) - np.sum(lf) ** 2
    assert abs(slope_err - np.sqrt(n * s2 / delta)) < 1e-14
    assert abs(norm_err - (norm - 10 ** (np.log10(norm) - np.sqrt(s2 * np.sum(lf ** 2) / delta)))) < 1e-12 * norm
    # the generating power law (3 f^-1.3) is recovered within the quoted uncertainties
    assert abs(slope + 1.3) < 3 * slope_err and abs(np.log10(norm / 3.0)) < 0.1


def test_linear_fit_bounds_and_break_detection():
    rng = np.random.default_rng(1)
    f = np.logspace(-3, 0, 300)
    steep = 2.0 * f ** -5.0 * np.exp(rng.normal(0, 0.05, len(f)))          # steeper than the allowed -4
    fit = fitting.linear_fit(f, steep)
    assert abs(fit.params["pow_slope"] + 4.0) < 1e-9 and fit.nvarys == 2
    # a power law that flattens above f = 0.05: the broken model wins the BIC and sets the frequency range
    fb = 0.05
    p = np.where(f < fb, (f / fb) ** -2.0, 1.0) * np.exp(rng.normal(0, 0.05, len(f)))
    best, fmax = fitting.chose_frequency_range(f, p)
    assert best.nvarys == 3 and abs(np.log(fmax / fb)) < 0.15
    assert abs(best.params["slope"] + 2.0) < 0.1
    # a pure power law: no break
    best, fmax = fitting.chose_frequency_range(f, 2.0 * f ** -1.5 * np.exp(rng.normal(0, 0.05, len(f))))
    assert best.nvarys == 2 and fmax == f.max()
    slope, slope_err, norm, norm_err = fitting.fit_lomb_scargle(f, 2.0 * f ** -1.5 * np.exp(rng.normal(0, 0.05, len(f))))
    assert abs(slope + 1.5) < 0.05 and slope_err > 0 and abs(norm / 2.0 - 1) < 0.1 and norm_err > 0
