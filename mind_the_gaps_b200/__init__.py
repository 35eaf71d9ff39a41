"""mind_the_gaps_b200 -- B200-native (sm_100a) implementation of the mind_the_gaps hot path.

Drop-in surface (same names and call signatures as andresgur/mind_the_gaps):
``mind_the_gaps_b200.gpmodelling.GPModelling``, ``.simulator.Simulator``,
``.lightcurves.GappyLightcurve``, ``.models`` (kernel terms, mean models, PSD models), ``.stats``,
``.noise_models``; plus ``.celerite`` -- the slice of the celerite API the reference calls --
evaluated by hand-written CUDA behind the C ABI of ``include/mtg_b200.h``.
"""
__version__ = "0.1.0"
