"""GPU parity: K1 (batched celerite log-likelihood) vs the oracle, through the C ABI.

Tolerance: north_star asks 1e-9 relative in FP64; the tests assert 1e-10."""
import ctypes

import numpy as np
import pytest

from mind_the_gaps_b200 import _native
from mind_the_gaps_b200.celerite import GP, terms
from mind_the_gaps_b200.celerite.modeling import ConstantModel
from mind_the_gaps_b200.models import BendingPowerlaw, Cosinus, DampedRandomWalk, Lorentzian
from oracle import gp as ogp

from helpers import C2_BOUNDS, c2_truth, fake_lightcurve

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def build(kind):
    if kind == "drw":
        return DampedRandomWalk(np.log(0.01), np.log(1e-6)), [("drw",)]
    if kind == "drw+lor":
        p = c2_truth()
        return (DampedRandomWalk(p[0], p[1], bounds=C2_BOUNDS[:2]) + Lorentzian(p[2], p[3], p[4], bounds=C2_BOUNDS[2:]),
                [("drw",), ("lorentzian",)])
    if kind == "complex+real":
        return (terms.ComplexTerm(log_a=-6.0, log_c=-16.0, log_d=-11.8) + terms.RealTerm(log_a=-6.0, log_c=-13.0),
                [("complex3",), ("real",)])
    if kind == "3sho+real":
        return (terms.SHOTerm(-6, 1.0, -12.0) + terms.SHOTerm(-7, 0.5, -11.0) + terms.SHOTerm(-8, 2.0, -10.5)
                + terms.RealTerm(-6.0, -14.0), [("sho",), ("sho",), ("sho",), ("real",)])
    if kind == "sho_overdamped+jitter":
        return terms.SHOTerm(-6.0, np.log(0.2), -12.0) + terms.JitterTerm(-4.0), [("sho",), ("jitter",)]
    if kind == "matern+real":
        return terms.Matern32Term(-2.0, 13.0) + terms.RealTerm(-6.0, -13.0), [("matern32", 0.01), ("real",)]
    if kind == "bpl+cos":
        return BendingPowerlaw(-5.0, -7.0, -12.0) + Cosinus(-8.0, -11.0), [("bendingpl",), ("cosinus",)]
    if kind == "complex4":
        return terms.ComplexTerm(-5.0, -8.0, -12.0, -12.5), [("complex4",)]
    raise KeyError(kind)


KINDS = ["drw", "drw+lor", "complex+real", "3sho+real", "sho_overdamped+jitter", "matern+real", "bpl+cos", "complex4"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("N", [1, 2, 255, 256, 257, 1000])
def test_loglike_matches_oracle(kind, N):
    t, y, dy, _ = fake_lightcurve(max(N, 2), 21)
    t, y, dy = t[:N], y[:N], dy[:N]
    kernel, spec = build(kind)
    gp = GP(kernel, mean=ConstantModel(float(np.mean(y))), fit_mean=False)
    gp.compute(t, dy + 1e-12)
    p0 = gp.get_parameter_vector()
    p = p0[None, :] + np.random.default_rng(N).normal(0, 0.05, (70, len(p0)))
    if kind == "sho_overdamped+jitter":
        p[:, 1] = np.log(0.2) + np.random.default_rng(1).normal(0, 0.05, 70)     # stay below Q = 1/2
    ll, status = gp.log_likelihood_batch(y, p, return_status=True, kernel="k1")
    ref = ogp.loglike_params(spec, p, t, y, dy)
    assert np.all(status == 0)
    np.testing.assert_allclose(ll, ref, rtol=RTOL)


def test_sho_straddling_q_half_in_one_batch():
    t, y, dy, _ = fake_lightcurve(300, 5)
    gp = GP(terms.SHOTerm(-6.0, 0.0, -12.0) + terms.RealTerm(-6.0, -13.0), mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    p = np.tile(gp.get_parameter_vector(), (64, 1))
    p[:, 1] = np.log(np.linspace(0.1, 3.0, 64))        # both regimes inside the same warps
    ll = gp.log_likelihood_batch(y, p, kernel="k1")
    spec = [("sho",), ("real",)]
    ref = np.concatenate([ogp.loglike_params(spec, p[i:i + 1], t, y, dy) for i in range(64)])
    np.testing.assert_allclose(ll, ref, rtol=RTOL)


def test_prior_and_not_positive_definite_status():
    t, y, dy, _ = fake_lightcurve(200, 6)
    kernel, spec = build("drw+lor")
    gp = GP(kernel, mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    p = np.tile(gp.get_parameter_vector(), (4, 1))
    p[1, 0] = 31.0        # above the upper bound of log_S0
    p[2, 3] = -6.0        # below the lower bound of log_Q
    ll, status = gp.log_likelihood_batch(y, p, apply_prior=True, return_status=True, kernel="k1")
    assert list(status) == [0, 3, 3, 0]
    assert np.isfinite(ll[0]) and ll[1] == -np.inf and ll[2] == -np.inf
    # without the prior the same points evaluate
    ll2 = gp.log_likelihood_batch(y, p, kernel="k1")
    np.testing.assert_allclose(ll2, ogp.loglike_params(spec, p, t, y, dy), rtol=RTOL)
    # a complex term violating a c >= b d is not positive definite -> -inf, status 1 (celerite LinAlgError)
    gp4 = GP(terms.ComplexTerm(-5.0, -8.0, -12.0, -12.5), mean=float(np.mean(y)))
    gp4.compute(t, dy + 1e-12)
    bad = np.array([[-5.0, 3.0, -12.0, -11.0]])
    ll4, st4 = gp4.log_likelihood_batch(y, bad, return_status=True, kernel="k1")
    assert ll4[0] == -np.inf and st4[0] == 1
    gp4.set_parameter_vector(bad[0])
    with pytest.raises(np.linalg.LinAlgError):
        gp4.log_likelihood(y)
    assert gp4.log_likelihood(y, quiet=True) == -np.inf


def test_single_point_api_matches_reference_call_sequence():
    """GPModelling._log_probability: set_parameter_vector -> log_prior -> log_likelihood."""
    t, y, dy, _ = fake_lightcurve(200, 7)
    kernel, spec = build("drw+lor")
    gp = GP(kernel, mean=ConstantModel(float(np.mean(y)), bounds=[(y.min(), y.max())]), fit_mean=False)
    gp.compute(t, dy + 1e-12)
    p = gp.get_parameter_vector() + 0.1
    gp.set_parameter_vector(p)
    assert gp.log_prior() == 0.0
    ref = ogp.loglike_params(spec, p[None, :], t, y, dy)[0]
    assert abs(gp.log_likelihood(y) - ref) <= RTOL * abs(ref)


def test_fitted_constant_mean():
    t, y, dy, _ = fake_lightcurve(300, 8)
    gp = GP(DampedRandomWalk(np.log(0.01), np.log(1e-6)), mean=ConstantModel(float(np.mean(y))), fit_mean=True)
    gp.compute(t, dy + 1e-12)
    assert gp.vector_size == 3
    p = np.tile(gp.get_parameter_vector(), (33, 1))
    p[:, 2] += np.linspace(-0.1, 0.1, 33)
    ll = gp.log_likelihood_batch(y, p, kernel="k1")
    ref = ogp.loglike_params([("drw",)], p, t, y, dy, fit_mean=True)
    np.testing.assert_allclose(ll, ref, rtol=RTOL)


@pytest.mark.parametrize("S,group", [(5, 13), (64, 1), (3, 100)])
def test_per_dataset_series_through_c_abi(S, group):
    """S data sets, `group` parameter sets each, frozen per-data-set mean (the LRT layout)."""
    N = 333
    t, _, _, _ = fake_lightcurve(N, 9)
    rng = np.random.default_rng(S)
    Y = 0.5 + 0.05 * rng.normal(size=(S, N))
    DY = rng.uniform(0.01, 0.03, (S, N))
    kernel, spec = build("drw+lor")
    gp = GP(kernel, mean=0.0)
    model = gp.device_model()
    B = S * group
    p = c2_truth()[None, :] + rng.normal(0, 0.1, (B, 5))
    ll = np.empty(B); status = np.empty(B, dtype=np.int32)
    diag = np.ascontiguousarray((DY + 1e-12) ** 2)
    means = np.ascontiguousarray(Y.mean(axis=1))
    rc = _native.lib().mtg_loglike_host(ctypes.byref(model), N, S, group, B, _native.ptr(np.ascontiguousarray(t)),
                                        _native.ptr(Y), _native.ptr(diag), _native.ptr(means), _native.ptr(p),
                                        _native.ptr(ll), _native.ptr(status), _native.FLAG_DS_MEAN)
    _native.check(rc)
    ref = np.concatenate([ogp.loglike_params(spec, p[s * group:(s + 1) * group], t, Y[s], DY[s]) for s in range(S)])
    np.testing.assert_allclose(ll, ref, rtol=RTOL)


def test_empty_batch_and_bad_arguments():
    t, y, dy, _ = fake_lightcurve(50, 10)
    gp = GP(DampedRandomWalk(np.log(0.01), np.log(1e-6)), mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    assert gp.log_likelihood_batch(y, np.zeros((0, 2))).shape == (0,)
    with pytest.raises(ValueError):
        gp.log_likelihood_batch(y, np.zeros((3, 5)))
    with pytest.raises(ValueError):
        gp.compute(t[::-1], dy)
    with pytest.raises(ValueError):
        gp.log_likelihood_batch(y[:-1], np.zeros((1, 2)))


def test_large_batch_linearity_property():
    """Size-independent property at the bench size: scaling y-mean residuals and dy by k and the
    amplitudes by k^2 shifts every log-likelihood by exactly -N log k."""
    N, B = 1000, 64 * 256
    t, y, dy, _ = fake_lightcurve(N, 12)
    kernel, spec = build("drw+lor")
    mu = float(np.mean(y))
    gp = GP(kernel, mean=mu)
    gp.compute(t, dy + 1e-12)
    p = c2_truth()[None, :] + np.random.default_rng(3).normal(0, 0.1, (B, 5))
    ll = gp.log_likelihood_batch(y, p, kernel="k1")
    k = 3.0
    gp2 = GP(build("drw+lor")[0], mean=k * mu)
    gp2.compute(t, k * (dy + 1e-12))
    p2 = p.copy(); p2[:, 0] += 2 * np.log(k); p2[:, 2] += 2 * np.log(k)
    ll2 = gp2.log_likelihood_batch(k * y, p2)
    np.testing.assert_allclose(ll2, ll - N * np.log(k), rtol=1e-9)
    # and a sample of the big batch against the oracle
    idx = np.random.default_rng(4).choice(B, 128, replace=False)
    np.testing.assert_allclose(ll[idx], ogp.loglike_params(spec, p[idx], t, y, dy), rtol=RTOL)


@pytest.mark.parametrize("S,group", [(1, 0), (40417, 9)])
def test_host_call_chunk_pipeline_matches_unpipelined_slices(S, group):
    """Batches of >= 4 K1 waves go through the host entry's H2D / kernel / D2H chunk pipeline (three streams,
    chunks of whole data-set groups): results and gradients must equal those of small unpipelined calls
    and the oracle, with and without per-data-set rates."""
    N = 24
    t, _, _, _ = fake_lightcurve(N, 21)
    rng = np.random.default_rng(S + group)
    if S == 1:
        B = 2 * 2 * 148 * 512 + 12345
        group = B                                    # one shared data set: every series belongs to group 0
        Y = 0.5 + 0.05 * rng.normal(size=(1, N)); DY = rng.uniform(0.01, 0.03, (1, N))
    else:
        B = S * group
        Y = 0.5 + 0.05 * rng.normal(size=(S, N)); DY = rng.uniform(0.01, 0.03, (S, N))
    kernel, spec = build("drw+lor")
    model = GP(kernel, mean=0.0).device_model()
    p = c2_truth()[None, :] + rng.normal(0, 0.1, (B, 5))
    diag = np.ascontiguousarray((DY + 1e-12) ** 2)
    means = np.ascontiguousarray(Y.mean(axis=1))
    tt = np.ascontiguousarray(t)
    lib = _native.lib()

    def call(sl, ds):
        Bc = sl.stop - sl.start
        ll = np.empty(Bc); g = np.empty((Bc, 5)); st = np.empty(Bc, dtype=np.int32)
        Yc = np.ascontiguousarray(Y[ds]); dc = np.ascontiguousarray(diag[ds]); mc = np.ascontiguousarray(means[ds])
        pc = np.ascontiguousarray(p[sl])
        _native.check(lib.mtg_loglike_host(ctypes.byref(model), N, Yc.shape[0], group, Bc, _native.ptr(tt), _native.ptr(Yc), _native.ptr(dc),
                                           _native.ptr(mc), _native.ptr(pc), _native.ptr(ll), _native.ptr(st), _native.FLAG_DS_MEAN))
        ll2 = np.empty(Bc)
        _native.check(lib.mtg_loglike_grad_host(ctypes.byref(model), N, Yc.shape[0], group, Bc, _native.ptr(tt), _native.ptr(Yc), _native.ptr(dc),
                                                _native.ptr(mc), _native.ptr(pc), _native.ptr(ll2), _native.ptr(g), _native.ptr(st),
                                                _native.FLAG_DS_MEAN))
        np.testing.assert_allclose(ll2, ll, rtol=1e-12)
        return ll, g, st

    ll, g, st = call(slice(0, B), slice(0, S))                      # pipelined
    assert np.all(st == 0) and np.all(np.isfinite(ll)) and np.all(np.isfinite(g))
    # the same series through small calls that stay below the pipeline threshold: first, middle (across a chunk seam), last
    g1 = group if S > 1 else 1
    chunk = -(-2 * 148 * 512 // g1) * g1
    for b0 in (0, chunk - 5 * g1, B - 7 * g1):
        sl = slice(b0, b0 + 7 * g1)
        ds = slice(b0 // g1, b0 // g1 + 7) if S > 1 else slice(0, 1)
        ll_s, g_s, _ = call(sl, ds)
        np.testing.assert_array_equal(ll[sl], ll_s)
        np.testing.assert_array_equal(g[sl], g_s)
    idx = rng.choice(B, 64, replace=False)
    ref = np.array([ogp.loglike_params(spec, p[i:i + 1], t, Y[i // group if S > 1 else 0], DY[i // group if S > 1 else 0])[0] for i in idx])
    np.testing.assert_allclose(ll[idx], ref, rtol=RTOL)


def test_c5_real_shape_parity_at_size():
    """config C5 exactly as bench.py runs it: N = 1e5 points, 3 x SHOTerm + RealTerm (celerite J = 7, ndim = 11),
    parameter sets = truth + N(0, 0.1).  K1 against (i) the NumPy recursion of the oracle on 4 parameter sets at full
    length, (ii) the C restatement of celerite's two-pass solver on 32, (iii) the dense Cholesky on an N = 3000 prefix
    (which pins the recursion itself at this width), and the status words."""
    import bench
    from oracle import cref
    t, y, dy = bench.c5_workload(100000)
    kernel = bench.c5_kernel()
    spec = [("sho",), ("sho",), ("sho",), ("real",)]
    gp = GP(kernel, mean=float(np.mean(y)))
    gp.compute(t, dy + 1e-12)
    p0 = gp.get_parameter_vector()
    assert len(p0) == 11
    p = p0[None, :] + np.random.default_rng(7).normal(0, 0.1, (32, 11))
    ll, status = gp.log_likelihood_batch(y, p, return_status=True, apply_prior=True)
    assert np.all(status == 0) and np.all(np.isfinite(ll))
    ref_c = cref.loglike_params(spec, p, t, y, dy)
    np.testing.assert_allclose(ll, ref_c, rtol=RTOL)
    ref_np = ogp.loglike_params(spec, p[:4], t, y, dy)
    np.testing.assert_allclose(ll[:4], ref_np, rtol=RTOL)
    # prefix against the dense Cholesky (the mean stays the full series' mean, as in the device call)
    n = 3000
    gp_s = GP(bench.c5_kernel(), mean=float(np.mean(y)))
    gp_s.compute(t[:n], dy[:n] + 1e-12)
    ll_s = gp_s.log_likelihood_batch(y[:n], p[:6])
    ref_d = ogp.loglike_params(spec, p[:6], t[:n], y[:n], dy[:n], mean=float(np.mean(y)), dense=True)
    np.testing.assert_allclose(ll_s, ref_d, rtol=1e-9)


def test_host_entries_on_two_devices_from_one_process():
    """VERDICT r1 (10) / ADVICE r1: the host entries keep one context (streams, pinned staging, device buffers) PER DEVICE.
    One process drives device 0, then device 1, then device 0 again through mtg_loglike_host / mtg_loglike_grad_host: every
    call must give the same (oracle-checked) numbers, also at a batch large enough for the chunk pipeline.  Runs when at
    least two GPUs are visible (`gpurun --gpus 2`); the single-GPU suite skips it."""
    lib = _native.lib()
    if lib.mtg_device_count() < 2:
        pytest.skip("needs two visible GPUs")
    N = 64
    t, y, dy, _ = fake_lightcurve(N, 5)
    kernel, spec = build("drw+lor")
    model = GP(kernel, mean=0.0).device_model()
    rng = np.random.default_rng(3)
    diag = np.ascontiguousarray((dy + 1e-12) ** 2); yy = np.ascontiguousarray(y - y.mean()); tt = np.ascontiguousarray(t)
    out = {}
    try:
        for B in (300, 2 * 2 * 148 * 512 + 77):
            p = np.ascontiguousarray(c2_truth()[None, :] + rng.normal(0, 0.1, (B, 5)))
            for visit, dev in enumerate((0, 1, 0)):
                _native.check(lib.mtg_set_device(dev))
                ll = np.empty(B); ll2 = np.empty(B); g = np.empty((B, 5)); st = np.empty(B, dtype=np.int32)
                _native.check(lib.mtg_loglike_host(ctypes.byref(model), N, 1, B, B, _native.ptr(tt), _native.ptr(yy), _native.ptr(diag), None,
                                                   _native.ptr(p), _native.ptr(ll), _native.ptr(st), 0))
                _native.check(lib.mtg_loglike_grad_host(ctypes.byref(model), N, 1, B, B, _native.ptr(tt), _native.ptr(yy), _native.ptr(diag), None,
                                                        _native.ptr(p), _native.ptr(ll2), _native.ptr(g), _native.ptr(st), 0))
                assert np.all(st == 0)
                out[(B, visit)] = (ll, g)
            for visit in (1, 2):
                np.testing.assert_array_equal(out[(B, visit)][0], out[(B, 0)][0])
                np.testing.assert_array_equal(out[(B, visit)][1], out[(B, 0)][1])
            idx = rng.choice(B, 32, replace=False)
            ref = ogp.loglike_params(spec, p[idx], t, y - y.mean(), dy)
            np.testing.assert_allclose(out[(B, 1)][0][idx], ref, rtol=RTOL)
    finally:
        lib.mtg_set_device(0)
