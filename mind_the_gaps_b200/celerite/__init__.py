"""The slice of the ``celerite`` API that mind_the_gaps uses (``celerite.GP``, ``celerite.terms``,
``celerite.modeling``), evaluated on a B200 by libmtg_b200 instead of celerite's C++ solver."""
from . import modeling, terms
from .gp import GP, LinAlgError

__all__ = ["GP", "LinAlgError", "modeling", "terms"]
