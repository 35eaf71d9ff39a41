"""Lightcurve file readers and ``to_csv`` against outputs of the unmodified reference readers
(``tests/golden/ref_io.npz`` <- ``tests/golden/make_golden.py io`` on the files in ``tests/golden/io/``)."""
import os
import warnings

import numpy as np
import pytest

from mind_the_gaps_b200.lightcurves import (ExposureTimeError, FermiLightcurve, GappyLightcurve, SimpleLightcurve,
                                            SwiftLightcurve)

HERE = os.path.dirname(os.path.abspath(__file__))
IO = os.path.join(HERE, "golden", "io")
FIELDS = ("times", "y", "dy", "exposures", "bkg_rate", "bkg_rate_err")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(HERE, "golden", "ref_io.npz"))


def check(ref, tag, lc, fields=FIELDS):
    for f in fields:
        np.testing.assert_array_equal(np.asarray(getattr(lc, f), dtype=float), ref["%s_%s" % (tag, f)], err_msg="%s.%s" % (tag, f))


@pytest.mark.parametrize("tag,fname,kw", [
    ("swift_obsid", "PCCURVE_obsid_mjd.qdp", {}),
    ("swift_obsid_cut", "PCCURVE_obsid_mjd.qdp", dict(minSNR=5, minSigma=3, minCts=50)),
    ("swift_snapshot", "PCCURVE_snapshot_sec.qdp", {}),
])
def test_swift_qdp_bit_exact(ref, tag, fname, kw, capsys):
    lc = SwiftLightcurve(os.path.join(IO, fname), **kw)
    check(ref, tag, lc)
    assert lc.n == len(ref[tag + "_times"]) and (tag != "swift_obsid_cut" or lc.n < 40)


@pytest.mark.parametrize("tag,fname,kw", [("simple6", "simple_6col.dat", {}), ("simple3", "simple_3col_days.csv", dict(delimiter=",")),
                                          ("simple4", "simple_4col.txt", {})])
def test_simple_text_formats_bit_exact(ref, tag, fname, kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        lc = SimpleLightcurve(os.path.join(IO, fname), **kw)
    check(ref, tag, lc)


def test_simple_without_exposures_warns():
    with pytest.warns(UserWarning, match="no exposures"):
        lc = SimpleLightcurve(os.path.join(IO, "simple_3col_days.csv"), delimiter=",")
    assert np.all(lc.exposures == 0) and lc.times[0] > 55000 * 86400.0 - 1


def test_fermi_csv(ref):
    lc = FermiLightcurve(os.path.join(IO, "fermi.csv"))
    check(ref, "fermi", lc, fields=("times", "y", "dy"))
    assert np.all(lc.exposures == 0) and np.all(lc.bkg_rate == 0)


def test_swift_proposal_style_file_is_rejected_like_the_reference(tmp_path):
    p = tmp_path / "short.qdp"
    p.write_text("!Time\tT_+ve\tT_-ve\tRate\tRatepos\tSigma\n1.0\t0.5\t-0.5\t0.1\t0.01\t3.0\n2.0\t0.5\t-0.5\t0.2\t0.01\t4.0\n")
    with pytest.raises(IndexError):
        SwiftLightcurve(str(p))


def test_to_csv_bytes_and_round_trip(ref, tmp_path):
    lc = SimpleLightcurve(os.path.join(IO, "simple_6col.dat"))
    out = tmp_path / "lc.dat"
    lc.to_csv(str(out))
    assert out.read_bytes() == ref["to_csv_text"].tobytes()
    again = SimpleLightcurve(str(out))
    for f in FIELDS:
        np.testing.assert_array_equal(getattr(again, f), getattr(lc, f))


def test_truncate_split_rand_remove(ref):
    lc = SimpleLightcurve(os.path.join(IO, "simple_6col.dat"))
    np.testing.assert_array_equal(lc.truncate(lc.times[5], lc.times[20]).times, ref["truncate_times"])
    parts = lc.split(2.0e6)
    np.testing.assert_array_equal([p.n for p in parts], ref["split_sizes"])
    np.testing.assert_array_equal([p.times[0] for p in parts], ref["split_first_times"])
    assert lc.rand_remove(7).n == lc.n - 7
    with pytest.raises(ValueError):
        lc.rand_remove(lc.n + 1)
    with pytest.raises(ValueError):
        lc.truncate(5.0, 1.0)


def test_exposure_spacing_check():
    with pytest.raises(ExposureTimeError):
        GappyLightcurve(np.array([0.0, 1.0, 2.0]), np.ones(3), np.ones(3), exposures=3.0)
