class ModelResult:
    pass


class Model:
    def __init__(self, *a, **k):
        raise NotImplementedError("lmfit is not available")
