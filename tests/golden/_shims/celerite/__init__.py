"""celerite stand-in for golden generation: the modeling/terms protocol + a GP evaluated by DENSE Cholesky
of the explicitly built covariance matrix (Foreman-Mackey et al. 2017, eq. 8/9).  NOT celerite's solver."""
from . import modeling, terms          # noqa: F401
from .gp import GP                     # noqa: F401

__version__ = "0.4.2-dense-standin"
