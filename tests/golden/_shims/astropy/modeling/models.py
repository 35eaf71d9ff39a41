"""astropy.modeling.models.custom_model stand-in: f(x, **defaults) -> callable class, parameters as attributes."""
import inspect


def custom_model(func):
    sig = inspect.signature(func)
    names = [p for p in list(sig.parameters)[1:]]
    defaults = {p: sig.parameters[p].default for p in names}

    class _Custom:
        param_names = tuple(names)

        def __init__(self, *args, **kwargs):
            vals = dict(defaults)
            for n, v in zip(names, args):
                vals[n] = v
            vals.update(kwargs)
            for n in names:
                setattr(self, n, vals[n])

        def __call__(self, x):
            return func(x, **{n: getattr(self, n) for n in names})

    _Custom.__name__ = func.__name__
    _Custom.__doc__ = func.__doc__
    return _Custom
