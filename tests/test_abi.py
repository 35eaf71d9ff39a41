"""CPU: the C-ABI library loads and exports every symbol include/mtg_b200.h declares."""
import ctypes
import os
import re

from mind_the_gaps_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mtg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mtg_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    lib = ctypes.CDLL(_native.library_path())
    names = declared_symbols()
    assert "mtg_loglike_dev" in names and "mtg_lbfgs_host" in names
    for name in names:
        assert hasattr(lib, name), "libmtg_b200.so does not export %s" % name


def test_binding_covers_header():
    assert set(declared_symbols()) == set(_native.SIGNATURES), \
        "ctypes SIGNATURES and include/mtg_b200.h disagree"


def test_version_and_model_width_without_gpu():
    lib = _native.lib()
    assert lib.mtg_version() == 100
    from mind_the_gaps_b200.celerite import GP
    from mind_the_gaps_b200.models import DampedRandomWalk, Lorentzian
    gp = GP(DampedRandomWalk(0.0, 0.0) + Lorentzian(0.0, 1.0, 0.5))
    m = gp.device_model()
    jc, jx = ctypes.c_int(), ctypes.c_int()
    assert lib.mtg_model_width(ctypes.byref(m), ctypes.byref(jc), ctypes.byref(jx)) == 0
    assert (jc.value, jx.value) == (4, 3)      # J = 4 as celerite counts it


def test_struct_layout_matches_header():
    # sizeof(mtg_model_t) computed from the header's field list
    n_terms, n_par = _native.MTG_MAX_TERMS, _native.MTG_MAX_PARAMS
    expect = 4 + 4 * n_terms + 16 * n_terms
    expect = (expect + 7) // 8 * 8 + 8 * n_terms          # extra[] is double aligned
    expect += 4 + 16 + 4                                   # mean_kind, mean_pidx, n_full
    expect = (expect + 7) // 8 * 8 + 8 * n_par             # full_default
    expect += 4 + 4 * n_par                                # n_free, free_idx
    expect = (expect + 7) // 8 * 8 + 16 * n_par            # lo, hi
    assert ctypes.sizeof(_native.MtgModel) == expect
