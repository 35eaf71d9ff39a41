"""GPU: K10, the ensemble sampler for few ensembles (one CTA per proposal, parallel-in-time log-probability, barrier
across CTAs per half step) -- same algorithm and random streams as K8, so its chain must coincide with K8's; its
log-probabilities equal the scan kernel's bit for bit and K1's / the oracle's to rounding."""
import ctypes

import numpy as np
import pytest

from mind_the_gaps_b200 import _native
from mind_the_gaps_b200.celerite import GP
from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
from mind_the_gaps_b200.sampler import DeviceEnsembleSampler
from oracle import gp as ogp

from helpers import C2_BOUNDS, c2_truth, fake_lightcurve
from test_ensemble_gpu import B_DRW, _drw_problem, _start

pytestmark = pytest.mark.gpu


def test_chain_equals_k8_chain_and_log_prob_equals_scan_kernel():
    t, y, dy, ex, gp = _drw_problem()
    W = 16
    p0 = _start(gp, W, np.random.default_rng(0))
    out = {}
    for kern in ("scan", "k8"):
        np.random.seed(5)                        # the sampler mixes one np.random draw into its Philox key
        s = DeviceEnsembleSampler(W, gp, y, seed=11, block=40, kernel=kern)
        s.run_mcmc(p0, 80)
        out[kern] = (s.get_chain().copy(), s.get_log_prob().copy(), s.acceptance_fraction.copy())
    chain, lp, acc = out["scan"]
    assert chain.shape == (80, W, 2) and lp.shape == (80, W)
    # same proposals, same uniforms: the two kernels walk the same chain
    np.testing.assert_allclose(chain, out["k8"][0], rtol=1e-9)
    np.testing.assert_allclose(lp, out["k8"][1], rtol=1e-10)
    np.testing.assert_array_equal(acc, out["k8"][2])
    flat = chain.reshape(-1, 2)
    np.testing.assert_array_equal(lp.reshape(-1), gp.log_probability_batch(y, flat, kernel="scan"))
    np.testing.assert_allclose(lp.reshape(-1)[::37], np.real(ogp.loglike_params([("drw",)], flat[::37], t, y, dy)), rtol=1e-10)
    lo = np.array([B_DRW["log_S0"][0], B_DRW["log_omega0"][0]]); hi = np.array([B_DRW["log_S0"][1], B_DRW["log_omega0"][1]])
    assert np.all(flat >= lo) and np.all(flat <= hi)
    assert 0.15 < acc.mean() < 0.9


def test_reproducible_and_independent_of_block_size():
    t, y, dy, ex, gp = _drw_problem()
    W = 13                                       # odd: the halves have 7 and 6 walkers
    p0 = _start(gp, W, np.random.default_rng(1))
    chains = []
    for block in (7, 50, 200):
        np.random.seed(5)
        s = DeviceEnsembleSampler(W, gp, y, seed=3, block=block, kernel="scan")
        s.run_mcmc(p0, 120)
        chains.append((s.get_chain().copy(), s.get_log_prob().copy()))
    for c, l in chains[1:]:
        assert np.array_equal(c, chains[0][0]) and np.array_equal(l, chains[0][1])


def test_c2_model_64_walkers_and_automatic_choice():
    """Config C2: DRW + Lorentzian, N = 1000, 64 walkers -> 32 CTAs; 'auto' routes one ensemble to K10."""
    t, y, dy, _ = fake_lightcurve(1000, 2)
    p = c2_truth()
    kernel = DampedRandomWalk(p[0], p[1], bounds=C2_BOUNDS[:2]) + Lorentzian(p[2], p[3], p[4], bounds=C2_BOUNDS[2:])
    gp = GP(kernel, mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    W = 64
    p0 = p[None, :] + np.random.default_rng(2).normal(0, 0.01, (W, 5))
    res = {}
    for kern in ("auto", "scan", "k8"):
        np.random.seed(5)
        s = DeviceEnsembleSampler(W, gp, y, seed=5, block=25, kernel=kern)
        s.run_mcmc(p0, 50)
        res[kern] = (s.get_chain().copy(), s.get_log_prob().copy())
    assert np.array_equal(res["auto"][0], res["scan"][0]) and np.array_equal(res["auto"][1], res["scan"][1])
    np.testing.assert_allclose(res["scan"][0], res["k8"][0], rtol=1e-9)
    flat = res["scan"][0].reshape(-1, 5)[::53]
    ref = ogp.loglike_params([("drw",), ("lorentzian",)], flat, t, y, dy)
    np.testing.assert_allclose(res["scan"][1].reshape(-1)[::53], ref, rtol=1e-10)


def test_several_ensembles_each_on_its_own_lightcurve_and_running_maximum():
    import torch
    E, W, N, nsteps = 3, 12, 300, 40
    t, _, _, _ = fake_lightcurve(N, 7)
    rng = np.random.default_rng(8)
    Y = 0.5 + 0.05 * rng.normal(size=(E, N)); DY = rng.uniform(0.01, 0.03, (E, N))
    gp = GP(DampedRandomWalk(-6.0, -13.5, bounds=B_DRW), mean=0.0)
    model = gp.device_model()
    dev = torch.device("cuda", 0)
    d_t = torch.from_numpy(np.ascontiguousarray(t)).to(dev)
    d_y = torch.from_numpy(np.ascontiguousarray(Y.T)).to(dev)                       # time-major [N][E]
    d_d = torch.from_numpy(np.ascontiguousarray(((DY + 1e-12) ** 2).T)).to(dev)
    d_m = torch.from_numpy(np.ascontiguousarray(Y.mean(axis=1))).to(dev)
    p0 = np.array([-6.0, -13.5])[None, None, :] + rng.normal(0, 0.05, (E, W, 2))
    outs = {}
    for name, flag in (("scan", _native.FLAG_SCAN), ("k8", _native.FLAG_NO_SCAN)):
        pos = torch.from_numpy(p0.copy()).to(dev); lp = torch.zeros((E, W), dtype=torch.float64, device=dev)
        chain = torch.empty((E, nsteps, W, 2), dtype=torch.float64, device=dev); clp = torch.empty((E, nsteps, W), dtype=torch.float64, device=dev)
        acc = torch.zeros((E, W), dtype=torch.int32, device=dev)
        blp = torch.empty(E, dtype=torch.float64, device=dev); bpos = torch.empty((E, 2), dtype=torch.float64, device=dev)
        _native.check(_native.lib().mtg_ensemble_dev(ctypes.byref(model), N, E, E, W, d_t.data_ptr(), d_y.data_ptr(), d_d.data_ptr(), d_m.data_ptr(),
                                                     pos.data_ptr(), lp.data_ptr(), 1, nsteps, 2.0, 77, 0, chain.data_ptr(), clp.data_ptr(),
                                                     acc.data_ptr(), blp.data_ptr(), bpos.data_ptr(), _native.FLAG_DS_MEAN | flag,
                                                     torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        outs[name] = [x.cpu().numpy() for x in (pos, lp, chain, clp, acc, blp, bpos)]
    for a, b in zip(outs["scan"], outs["k8"]):
        np.testing.assert_allclose(a, b, rtol=1e-9)
    pos, lp, chain, clp, acc, blp, bpos = outs["scan"]
    # final state = last chain entry; running maximum >= every stored log-probability and attained at best_pos
    np.testing.assert_array_equal(chain[:, -1], pos); np.testing.assert_array_equal(clp[:, -1], lp)
    for e in range(E):
        assert blp[e] >= clp[e].max()
        ref = ogp.loglike_params([("drw",)], bpos[e][None, :], t, Y[e], DY[e])[0]
        assert abs(ref - blp[e]) <= 1e-10 * abs(ref)
        refs = ogp.loglike_params([("drw",)], chain[e, ::9].reshape(-1, 2), t, Y[e], DY[e])
        np.testing.assert_allclose(clp[e, ::9].reshape(-1), refs, rtol=1e-10)


def test_four_state_dimensions_two_sho_terms():
    """J = 4 (two SHO terms, the second one crossing Q = 1/2 inside the ensemble): K10 == K8 chain, oracle log-probabilities."""
    from mind_the_gaps_b200.celerite import terms
    t, y, dy, _ = fake_lightcurve(400, 5)
    b = dict(log_S0=(-15, 5), log_Q=(-3, 4), log_omega0=(-16, -8))
    gp = GP(terms.SHOTerm(-6.0, np.log(3.0), -12.0, bounds=b) + terms.SHOTerm(-7.0, np.log(0.5), -11.0, bounds=b), mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    W = 14
    p0 = gp.get_parameter_vector()[None, :] + np.random.default_rng(3).normal(0, 0.05, (W, 6))
    res = {}
    for kern in ("scan", "k8"):
        np.random.seed(5)
        s = DeviceEnsembleSampler(W, gp, y, seed=9, block=30, kernel=kern)
        s.run_mcmc(p0, 60)
        res[kern] = (s.get_chain().copy(), s.get_log_prob().copy())
    np.testing.assert_allclose(res["scan"][0], res["k8"][0], rtol=1e-9)
    flat = res["scan"][0].reshape(-1, 6)[::41]
    ref = np.concatenate([ogp.loglike_params([("sho",), ("sho",)], flat[i:i + 1], t, y, dy) for i in range(len(flat))])
    np.testing.assert_allclose(res["scan"][1].reshape(-1)[::41], ref, rtol=1e-10)
