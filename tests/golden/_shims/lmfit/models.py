from .model import Model


class LognormalModel(Model):
    pass


class LinearModel(Model):
    pass


class SineModel(Model):
    pass


class ConstantModel(Model):
    pass
