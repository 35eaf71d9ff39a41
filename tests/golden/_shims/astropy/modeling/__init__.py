class Model:
    """placeholder base; see models.custom_model"""
    param_names = ()
from . import models  # noqa: E402,F401
