"""K12 / K13 (csrc/mtg_tk95.cu): Timmer & Koenig synthesis fused with the exposure binning -- the inverse real FFT
split as n = N1 x N2 with shared-memory column FFTs and only the samples the bins need ever formed -- against

* the NumPy oracle (np.fft.irfft + the reference's binning) on IDENTICAL injected normals, for every column length
  N1 = 2^3 .. 2^13 the kernel can be configured with, odd and even N2, table and celerite PSDs;
* the spectrum -> cuFFT -> binning chain of the same library on the same Philox seed, at the production grid size.

Index work (segment, bin membership) is exact; floating point agrees to FFT rounding (north_star: 1e-9)."""
import os

import numpy as np
import pytest

from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from mind_the_gaps_b200.models.psd_models import BendingPowerlaw
from mind_the_gaps_b200.simulator import Simulator, _fast_length
from oracle import simulator as osim

from helpers import window

pytestmark = pytest.mark.gpu
DAY = 86400.0


class _env(object):
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        for k, v in self.kv.items():
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _oracle_rates(simu, psd, mean, normals, u):
    n = simu.fftndatapoints
    grid = simu.sim_timestamps
    out = []
    for l in range(len(u)):
        full = osim.simulate_regular(n, simu.sim_dt, psd, mean, normals[l, 0], normals[l, 1])
        i0, i1, shift = osim.cut_segment(grid, simu.sim_dt, simu.sim_duration, simu.strategy[0][0], u[l])
        out.append((i0, i1, osim.downsample(grid[i0:i1] + shift, full[i0:i1], simu.strategy)))
    return out


@pytest.mark.parametrize("log2n1,odd", [(3, True), (4, False), (5, True), (6, False), (7, True), (8, False), (9, True),
                                        (10, False), (11, True), (12, False), (13, True), (13, False)])
def test_fused_matches_oracle_on_identical_noise(log2n1, odd):
    t, exposures = window(60, 3)
    psd = BendingPowerlaw(S0=0.01, omega0=2 * np.pi / (20 * DAY))
    n_req = Simulator(psd, t, exposures, 0.5, extension_factor=2).fftndatapoints
    n1 = 1 << log2n1
    n2 = -(-n_req // n1)
    if (n2 % 2 == 1) != odd:
        n2 += 1
    with _env(MTG_SIM_FUSED=2, MTG_SIM_FUSED_LOG2N1=log2n1):
        simu = Simulator(psd, t, exposures, 0.5, pdf="Gaussian", extension_factor=2, fft_length=n1 * n2)
        assert simu.fused
        M = simu.fftndatapoints // 2 + 1
        rng = np.random.default_rng(7 + log2n1)
        L = 3
        normals = rng.normal(size=(L, 2, M)); u = rng.random(L)
        rates, info = simu.generate_lightcurves(L, normals=normals, u_segment=u, return_info=True)
    for l, (i0, i1, ref) in enumerate(_oracle_rates(simu, psd, 0.5, normals, u)):
        assert (info["segment"][l, 0], info["segment"][l, 1]) == (i0, i1 - i0)            # index work: exact
        np.testing.assert_allclose(rates[l], ref, rtol=1e-9, atol=1e-12)


def test_fused_celerite_psd_and_zero_mean():
    """PSD evaluated on the device from kernel parameters (one parameter set per lightcurve), mean 0 (no per-sample shift)."""
    t, exposures = window(40, 5)
    kernel = DampedRandomWalk(np.log(0.01), np.log(2 * np.pi / (20 * DAY))) + \
        Lorentzian(np.log(0.005), np.log(30.0), np.log(2 * np.pi / (5 * DAY)))
    with _env(MTG_SIM_FUSED=2):
        simu = Simulator(kernel.get_psd, t, exposures, 0.0, extension_factor=2, fft_length="fast")
        assert simu.fused
        M = simu.fftndatapoints // 2 + 1
        rng = np.random.default_rng(2)
        L = 2
        normals = rng.normal(size=(L, 2, M)); u = rng.random(L)
        rates = simu.generate_lightcurves(L, normals=normals, u_segment=u)
    for l, (_, _, ref) in enumerate(_oracle_rates(simu, kernel.get_psd, 0.0, normals, u)):
        np.testing.assert_allclose(rates[l], ref, rtol=1e-9, atol=1e-9 * np.std(ref))


def test_fused_equals_cufft_chain_on_the_same_seed_small():
    """Philox path: both implementations draw spectrum entry k of lightcurve l from the same counter."""
    t, exposures = window(80, 11)
    psd = BendingPowerlaw(S0=0.01, omega0=2 * np.pi / (20 * DAY))
    out = {}
    for mode in (0, 2):
        with _env(MTG_SIM_FUSED=mode):
            simu = Simulator(psd, t, exposures, 0.5, extension_factor=2, fft_length="fast")
            assert simu.fused == (mode == 2)
            out[mode] = simu.generate_lightcurves(37, seed=1234)
            simu.close()
    # the production kernel stages the generated entries in single precision (a Philox deviate carries no more)
    np.testing.assert_allclose(out[2], out[0], rtol=2e-6, atol=2e-6 * np.std(out[0]))


def test_fused_equals_cufft_chain_at_production_size():
    """Config C3 / C4 shape: 1000 gapped points, n_fft ~ 2.7e6 = 8192 x N2; the default policy picks the fused path."""
    t, exposures = window(1000, 2)
    kernel = DampedRandomWalk(np.log(0.01), np.log(2 * np.pi / (20 * DAY)))
    P = np.tile(kernel.get_parameter_vector(), (6, 1)) + np.random.default_rng(0).normal(0, 0.1, (6, 2))
    out = {}
    for mode in (0, 1):
        with _env(MTG_SIM_FUSED=mode):
            simu = Simulator(kernel.get_psd, t, exposures, 0.5, extension_factor=2, fft_length="fast")
            assert simu.fftndatapoints == _fast_length(simu.fftndatapoints) and simu.fftndatapoints % 8192 == 0
            assert simu.fused == (mode == 1)
            out[mode] = simu.generate_lightcurves(6, parameters=P, seed=99)
            simu.close()
    assert np.all(np.isfinite(out[1]))
    np.testing.assert_allclose(out[1], out[0], rtol=2e-6, atol=2e-6 * np.std(out[0]))


def test_single_role_and_warp_specialised_kernels_agree():
    """MTG_TK95_WS=0 runs generation and transform back to back in the same warps (double-precision staging)."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from helpers import window
        from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
        from mind_the_gaps_b200.simulator import Simulator
        t, ex = window(300, 4)
        k = DampedRandomWalk(np.log(0.01), np.log(2 * np.pi / (20 * 86400.0))) + Lorentzian(np.log(0.005), np.log(30.0), np.log(2 * np.pi / (5 * 86400.0)))
        s = Simulator(k.get_psd, t, ex, 0.5, extension_factor=2, fft_length="fast")
        assert s.fused
        np.save(sys.argv[1], s.generate_lightcurves(9, seed=5))
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    out = {}
    with tempfile.TemporaryDirectory() as d:
        for ws in ("0", "1"):
            f = os.path.join(d, "o%s.npy" % ws)
            subprocess.run([sys.executable, "-c", code, f], check=True, env=dict(os.environ, MTG_TK95_WS=ws, MTG_SIM_FUSED="2"), timeout=300)
            out[ws] = np.load(f)
    np.testing.assert_allclose(out["1"], out["0"], rtol=2e-6, atol=2e-6 * np.std(out["0"]))
