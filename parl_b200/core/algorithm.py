"""``parl.Algorithm`` — mirror of parl/core/torch/algorithm.py:24-92 and parl/core/algorithm_base.py:18-111."""
from .model import Model

__all__ = ['Algorithm']


class Algorithm(object):
    def __init__(self, model=None):
        assert isinstance(model, Model)
        self.model = model

    def get_weights(self):
        return self.model.get_weights()

    def set_weights(self, params):
        self.model.set_weights(params)

    def learn(self, *args, **kwargs):
        raise NotImplementedError

    def predict(self, *args, **kwargs):
        raise NotImplementedError

    def sample(self, *args, **kwargs):
        raise NotImplementedError
