"""GPU: the inner loop's own exp / sincos / reciprocal (csrc/mtg_math.cuh) against numpy in extended precision.
variant 0 = polynomial kernels, variant 1 = the 256-entry table kernels K2/K8 run, variant 2 = the 2048-entry table kernels K1 runs."""
import numpy as np
import pytest

from mind_the_gaps_b200 import _native

pytestmark = pytest.mark.gpu


def run(x, variant):
    x = np.ascontiguousarray(x, dtype=np.float64)
    e = np.empty_like(x); s = np.empty_like(x); c = np.empty_like(x); r = np.empty_like(x)
    _native.check(_native.lib().mtg_test_math_host(len(x), variant, _native.ptr(x), _native.ptr(e), _native.ptr(s),
                                                   _native.ptr(c), _native.ptr(r)))
    return e, s, c, r


def ulps(got, ref):
    ref = np.asarray(ref, dtype=np.longdouble)
    return np.abs((got - ref).astype(np.float64)) / np.spacing(np.abs(ref.astype(np.float64)))


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_exp2_of_nonpositive_arguments(variant):
    """2^t, t <= 0: the recursion evaluates exp(-c dt) as 2^(-c log2(e) dt) with an exact range reduction."""
    rng = np.random.default_rng(0)
    x = -np.concatenate([rng.uniform(0, 1, 200000), 10.0 ** rng.uniform(-12, 3.0, 200000),
                         np.arange(0, 2048) / 256.0, np.arange(0, 2048) / 512.0, np.arange(0, 8192) / 4096.0,
                         [0.0, 1e-300, 0.5, 1.5, 1000.0, 1019.5]])
    e = run(x, variant)[0]
    assert ulps(e, np.exp2(x.astype(np.longdouble))).max() <= 1.5
    e2 = run(np.array([-1020.5, -1100.0, -1e6, -1e300, -np.inf]), variant)[0]
    assert np.all(e2 == 0.0)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_sincos_phase_range(variant):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e5, 1e5, 200000), rng.uniform(1e5, 1e9, 100000), rng.uniform(1e9, 1.0e12, 100000), rng.uniform(1.2e12, 1e15, 1000),
                        np.arange(-600, 600) * (np.pi / 256), np.arange(-5000, 5000) * (np.pi / 2048),
                        [0.0, np.pi / 4, np.pi / 2, np.pi, 1e-8, -1e-8, 99999.9]])
    _, s, c, _ = run(x, variant)
    xl = x.astype(np.longdouble)
    # absolute error relative to 1 (values near a zero crossing are limited by the argument's own rounding)
    assert np.max(np.abs((s - np.sin(xl)).astype(float))) < 4e-16
    assert np.max(np.abs((c - np.cos(xl)).astype(float))) < 4e-16


def test_reciprocal_of_positive_normals():
    """1/D of the recursion: MUFU seed + one cubic Newton step."""
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(1, 2, 200000), 10.0 ** rng.uniform(-300, 300, 200000), [1.0, 2.0, 3.0, 1e-300, 1e300]])
    r = run(x, 1)[3]
    assert ulps(r, 1.0 / x.astype(np.longdouble)).max() <= 1.5
