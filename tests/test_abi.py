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


def test_sim_config_layout_and_flag_constants_match_header():
    """mtg_sim_config_t: nine 4-byte / 8-byte scalars in the header's order, then six pointers; the MTG_SIM_* / enum
    constants of the binding are the header's."""
    text = open(os.path.join(ROOT, "include", "mtg_b200.h")).read()
    body = re.search(r"typedef struct mtg_sim_config \{(.*?)\} mtg_sim_config_t;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [(t.strip(), n) for t, n in re.findall(r"([A-Za-z0-9_ ]+?[ \*])(\w+);", body)]
    names = [n for _, n in fields]
    assert names == [f[0] for f in _native.MtgSimConfig._fields_]
    size = 0
    for t, _ in fields:
        w = 4 if t.startswith("int32_t") else 8
        size = (size + w - 1) // w * w + w
    assert ctypes.sizeof(_native.MtgSimConfig) == (size + 7) // 8 * 8
    assert int(re.search(r"#define MTG_SIM_E13_PADDED (\d+)", text).group(1)) == _native.SIM_E13_PADDED
    assert (_native.PDF_GAUSSIAN, _native.PDF_LOGNORMAL, _native.PDF_UNIFORM) == (0, 1, 2)
    assert re.search(r"MTG_PDF_GAUSSIAN = 0, MTG_PDF_LOGNORMAL = 1, MTG_PDF_UNIFORM = 2", text)
    assert re.search(r"MTG_NOISE_POISSON = 0, MTG_NOISE_KRAFT = 1, MTG_NOISE_GAUSSIAN = 2", text)
    assert (_native.NOISE_POISSON, _native.NOISE_KRAFT, _native.NOISE_GAUSSIAN) == (0, 1, 2)
