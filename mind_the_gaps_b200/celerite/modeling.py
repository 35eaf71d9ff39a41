"""The ``celerite.modeling`` protocol used by mind_the_gaps, re-implemented without celerite.

The reference subclasses ``celerite.modeling.Model`` for its mean models
(``/root/reference/mind_the_gaps/models/mean_models.py:6-31``) and relies on the parameter
bookkeeping of ``Model`` / ``ModelSet`` through ``celerite.GP`` (``gpmodelling.py:55, 147, 149, 193,
454``): named parameters in a fixed order, optional box bounds given as a list or a dict, frozen /
thawed masks, ``log_prior`` = 0 inside the box else ``-inf``, and ``name:sub`` parameter names for
composite models.
"""
from collections import OrderedDict

import numpy as np

__all__ = ["Model", "ModelSet", "ConstantModel"]


class Model(object):
    """Base class: an ordered set of named scalar parameters with bounds and a frozen mask."""

    parameter_names = tuple()

    def __init__(self, *args, **kwargs):
        names = tuple(self.parameter_names)
        self.unfrozen_mask = np.ones(len(names), dtype=bool)
        self.dirty = True
        bounds = kwargs.pop("bounds", None)
        if len(args):
            if len(args) != len(names):
                raise ValueError("expected {0} arguments but got {1}".format(len(names), len(args)))
            if len(kwargs):
                raise ValueError("parameters must be fully specified by arguments or keyword arguments, not both")
            values = list(args)
        else:
            missing = [k for k in names if k not in kwargs]
            if missing:
                raise ValueError("missing parameter '{0}'".format(missing[0]))
            values = [kwargs.pop(k) for k in names]
            if len(kwargs):
                raise ValueError("unrecognized parameter(s) {0}".format(sorted(kwargs)))
        self.parameter_vector = np.array(values, dtype=np.float64)
        if bounds is None:
            self.parameter_bounds = [(None, None) for _ in names]
        elif isinstance(bounds, dict):
            self.parameter_bounds = [tuple(bounds.get(k, (None, None))) for k in names]
        else:
            bounds = list(bounds)
            if len(bounds) != len(names):
                raise ValueError("the number of bounds must equal the number of parameters")
            self.parameter_bounds = [tuple(b) if b is not None else (None, None) for b in bounds]
        for b in self.parameter_bounds:
            if len(b) != 2:
                raise ValueError("invalid bounds for parameter")
        if not np.isfinite(self.log_prior()):
            raise ValueError("non-finite log prior value")

    # -- attribute access by parameter name ---------------------------------------------------
    def __getattr__(self, name):
        if name.startswith("__") or name in ("parameter_names", "parameter_vector"):
            raise AttributeError(name)
        names = type(self).parameter_names
        if name in names and "parameter_vector" in self.__dict__:
            return self.__dict__["parameter_vector"][names.index(name)]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        names = type(self).parameter_names
        if name in names and "parameter_vector" in self.__dict__:
            self.parameter_vector[names.index(name)] = value
            self.dirty = True
        else:
            object.__setattr__(self, name, value)

    def __len__(self):
        return self.vector_size

    # -- sizes ----------------------------------------------------------------------------------
    @property
    def full_size(self):
        return len(self.parameter_names)

    @property
    def vector_size(self):
        return int(self.unfrozen_mask.sum())

    # -- getters / setters -----------------------------------------------------------------------
    def get_parameter_names(self, include_frozen=False):
        if include_frozen:
            return tuple(self.parameter_names)
        return tuple(p for p, f in zip(self.parameter_names, self.unfrozen_mask) if f)

    def get_parameter_bounds(self, include_frozen=False):
        if include_frozen:
            return list(self.parameter_bounds)
        return [b for b, f in zip(self.parameter_bounds, self.unfrozen_mask) if f]

    def get_parameter_vector(self, include_frozen=False):
        if include_frozen:
            return self.parameter_vector
        return self.parameter_vector[self.unfrozen_mask]

    def set_parameter_vector(self, vector, include_frozen=False):
        v = self.parameter_vector
        if include_frozen:
            v[:] = vector
        else:
            v[self.unfrozen_mask] = vector
        self.parameter_vector = v
        self.dirty = True

    def get_parameter_dict(self, include_frozen=False):
        return OrderedDict(zip(self.get_parameter_names(include_frozen=include_frozen),
                               self.get_parameter_vector(include_frozen=include_frozen)))

    def get_parameter(self, name):
        return self.parameter_vector[self._index(name)]

    def set_parameter(self, name, value):
        self.parameter_vector[self._index(name)] = value
        self.dirty = True

    def _index(self, name):
        try:
            return self.get_parameter_names(include_frozen=True).index(name)
        except ValueError:
            raise ValueError("unknown parameter '{0}'".format(name))

    # -- freezing ---------------------------------------------------------------------------------
    def freeze_parameter(self, name):
        self.unfrozen_mask[self._index(name)] = False

    def thaw_parameter(self, name):
        self.unfrozen_mask[self._index(name)] = True

    def freeze_all_parameters(self):
        self.unfrozen_mask[:] = False

    def thaw_all_parameters(self):
        self.unfrozen_mask[:] = True

    # -- prior / value ------------------------------------------------------------------------------
    def log_prior(self):
        for p, b in zip(self.get_parameter_vector(include_frozen=True), self.get_parameter_bounds(include_frozen=True)):
            if b[0] is not None and p < b[0]:
                return -np.inf
            if b[1] is not None and p > b[1]:
                return -np.inf
        return 0.0

    def get_value(self, *args, **kwargs):
        raise NotImplementedError("overloaded by subclasses")

    def compute_gradient(self, *args, **kwargs):
        raise NotImplementedError("overloaded by subclasses")


class ModelSet(Model):
    """An ordered collection of named sub-models exposed as one flat parameter vector with
    ``name:parameter`` names (celerite's ``GP`` = ``kernel`` + ``mean``; ``TermSum`` = ``terms[i]``)."""

    def __init__(self, models):
        object.__setattr__(self, "models", OrderedDict(models))
        object.__setattr__(self, "dirty", True)

    def __getattr__(self, name):
        models = self.__dict__.get("models")
        if models is not None and name in models:
            return models[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)

    @property
    def dirty(self):
        return self.__dict__.get("_dirty", True) or any(m.dirty for m in self.models.values())

    @dirty.setter
    def dirty(self, value):
        self.__dict__["_dirty"] = value
        for m in self.models.values():
            m.dirty = value

    @property
    def parameter_names(self):
        return self.get_parameter_names(include_frozen=True)

    @property
    def parameter_vector(self):
        return self.get_parameter_vector(include_frozen=True)

    @property
    def parameter_bounds(self):
        return self.get_parameter_bounds(include_frozen=True)

    @property
    def unfrozen_mask(self):
        parts = [m.unfrozen_mask for m in self.models.values()]
        return np.concatenate(parts) if parts else np.zeros(0, dtype=bool)

    @property
    def full_size(self):
        return sum(m.full_size for m in self.models.values())

    @property
    def vector_size(self):
        return sum(m.vector_size for m in self.models.values())

    def get_parameter_names(self, include_frozen=False):
        out = []
        for name, m in self.models.items():
            out.extend("{0}:{1}".format(name, p) for p in m.get_parameter_names(include_frozen=include_frozen))
        return tuple(out)

    def get_parameter_bounds(self, include_frozen=False):
        out = []
        for m in self.models.values():
            out.extend(m.get_parameter_bounds(include_frozen=include_frozen))
        return out

    def get_parameter_vector(self, include_frozen=False):
        parts = [np.atleast_1d(m.get_parameter_vector(include_frozen=include_frozen)) for m in self.models.values()]
        return np.concatenate(parts) if parts else np.zeros(0)

    def set_parameter_vector(self, vector, include_frozen=False):
        vector = np.asarray(vector, dtype=np.float64)
        i = 0
        for m in self.models.values():
            n = m.full_size if include_frozen else m.vector_size
            m.set_parameter_vector(vector[i:i + n], include_frozen=include_frozen)
            i += n
        if i != len(vector):
            raise ValueError("dimension mismatch: got {0} values for {1} parameters".format(len(vector), i))
        self.__dict__["_dirty"] = True

    def _split(self, name):
        head, _, rest = name.partition(":")
        if head not in self.models or not rest:
            raise ValueError("unknown parameter '{0}'".format(name))
        return self.models[head], rest

    def get_parameter(self, name):
        m, rest = self._split(name)
        return m.get_parameter(rest)

    def set_parameter(self, name, value):
        m, rest = self._split(name)
        m.set_parameter(rest, value)
        self.__dict__["_dirty"] = True

    def freeze_parameter(self, name):
        m, rest = self._split(name)
        m.freeze_parameter(rest)

    def thaw_parameter(self, name):
        m, rest = self._split(name)
        m.thaw_parameter(rest)

    def freeze_all_parameters(self):
        for m in self.models.values():
            m.freeze_all_parameters()

    def thaw_all_parameters(self):
        for m in self.models.values():
            m.thaw_all_parameters()

    def log_prior(self):
        lp = 0.0
        for m in self.models.values():
            lp += m.log_prior()
            if not np.isfinite(lp):
                return -np.inf
        return lp


class ConstantModel(Model):
    """A constant mean (``gpmodelling.py:85-86, 94-95``)."""

    parameter_names = ("value",)

    def get_value(self, x):
        return self.value + np.zeros(len(np.atleast_1d(x)))

    def compute_gradient(self, x):
        return np.ones((1, len(np.atleast_1d(x))))
