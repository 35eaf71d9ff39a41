"""emcee stand-in: import-time placeholder (golden generation only; no sampling fixtures)."""
__version__ = "3.1.4-placeholder"


class EnsembleSampler:
    def __init__(self, *a, **k):
        raise NotImplementedError("emcee is not available")


class autocorr:
    class AutocorrError(Exception):
        pass
