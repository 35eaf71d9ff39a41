"""``SwiftLightcurve``: a ``GappyLightcurve`` from a Swift-XRT light-curve file in QDP format, as produced by
the UKSSDC pipeline (``PCCURVE.qdp`` / ``WTCURVE.qdp``) -- reference
``mind_the_gaps/lightcurves/swiftlightcurve.py:10-108``.

File layout: two QDP command lines, then a ``!``-commented, tab-separated header
``Time T_+ve T_-ve Rate Ratepos Rateneg FracExp BGrate BGerr CorrFact CtsInSrc BGInSrc Exposure Sigma SNR [ObsID]``
(15 columns for snapshot binning, 16 with the ObsID of observation binning).  Columns are taken by
position as the reference does: 0 time, 3 rate (+ ``<rate>pos`` / ``<rate>neg``), 7/8 background rate and
error, 9 correction factor, 12 exposure.  The exposure is divided by the correction factor and the
background rate multiplied by it, so that ``bkg_rate * exposure`` stays the background counts in the
source region (swiftlightcurve.py:53-58).
"""
import numpy as np

from ._table import read_named, to_seconds
from .gappylightcurve import GappyLightcurve

_FLOATS = "f8, f8, f8, f8, f8, f8, f8, f8, f8, f8, i8, f8, f8, f8, f8"


class SwiftLightcurve(GappyLightcurve):
    def __init__(self, swift_xrt_file="PCCURVE.qdp", minSNR=0, minSigma=0, minCts=0):
        try:
            data = self.readPCCURVE(swift_xrt_file, minSNR=minSNR, minSigma=minSigma, minCts=minCts)
        except ValueError:
            data = self.readsimplePCCURVE(swift_xrt_file, minSigma=minSigma)
        names = data.dtype.names
        if len(names) < 13:
            # the 6-column proposal-style file carries no background / exposure columns: the reference
            # indexes past the end of the name tuple here (swiftlightcurve.py:36-41) -> same exception type
            raise IndexError("%s has %d columns; exposure, background and correction-factor columns are needed"
                             % (swift_xrt_file, len(names)))
        rate = names[3]
        in_days = names[0] == "MJD"
        print("Time columnd in MJD" if in_days else "Time columnd in seconds")
        time = to_seconds(data[names[0]], in_days)
        yerr = (data["%spos" % rate] - data["%sneg" % rate]) / 2      # Rateneg is stored negative
        corr = data[names[9]]
        super().__init__(time, data[rate], yerr, data[names[12]] / corr, data[names[7]] * corr, data[names[8]] * corr)

    def readPCCURVE(self, file="PCCURVE.qdp", minExposure=0, minSigma=0, minSNR=0, minCts=0):
        """Rows of the pipeline file passing ``Exposure > minExposure``, ``SNR > minSNR``, ``Sigma > minSigma``,
        ``CtsInSrc >= minCts`` (strict / non-strict exactly as swiftlightcurve.py:88)."""
        print("Trying to read data from %s" % file)
        common = dict(delimiter="\t", skip_header=2, comments="!")
        try:
            data = read_named(file, dtype=_FLOATS + ", U30", **common)        # observation binning: trailing ObsID
            print("File with obsids")
        except ValueError:
            data = read_named(file, dtype=_FLOATS, **common)                  # snapshot binning
            print("File with snapshots")
        keep = (data["Exposure"] > minExposure) & (data["SNR"] > minSNR) & (data["Sigma"] > minSigma) & (data["CtsInSrc"] >= minCts)
        print("Filtered %d datapoints by minSNR = %d and minSigma = %.2f, minCts = %d and minExp=%.1f"
              % (len(data) - np.count_nonzero(keep), minSNR, minSigma, minCts, minExposure))
        return data[keep]

    def readsimplePCCURVE(self, file="PCCURVE.qdp", minExposure=0, minSigma=0, minSNR=0):
        """The 6-column file attached to proposals (time, T_+ve, T_-ve, rate, +err, -err/Sigma)."""
        print("Reading %s data" % file)
        data = read_named(file, delimiter="\t", comments="!", dtype="f8, f8, f8, f8, f8, f8", deletechars="~", usecols=(0, 1, 2, 3, 4, 5))
        keep = ((data["T_+ve"] - data["T_-ve"]) > minExposure) & (data["Sigma"] > minSigma)
        print("Filtered %d datapoints by minSNR = %d and minSigma = %.2f" % (len(data) - np.count_nonzero(keep), minSNR, minSigma))
        return data[keep]
