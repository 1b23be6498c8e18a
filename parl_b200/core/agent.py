"""``parl.Agent`` — mirror of parl/core/torch/agent.py:29-176 and parl/core/agent_base.py:16-89
(save creates missing directories, restore reproduces outputs, train()/eval() toggle
``agent.training`` and the module mode: parl/core/torch/tests/agent_base_test_torch.py:109-170)."""
import os

import torch

from .algorithm import Algorithm

__all__ = ['Agent']


class Agent(object):
    def __init__(self, algorithm):
        assert isinstance(algorithm, Algorithm)
        self.alg = algorithm
        self.training = True

    def get_weights(self):
        return self.alg.get_weights()

    def set_weights(self, params):
        self.alg.set_weights(params)

    def learn(self, *args, **kwargs):
        raise NotImplementedError

    def predict(self, *args, **kwargs):
        raise NotImplementedError

    def sample(self, *args, **kwargs):
        raise NotImplementedError

    def save(self, save_path, model=None):
        if model is None:
            model = self.alg.model
        dirname = os.path.dirname(save_path)
        if dirname != '' and not os.path.exists(dirname):
            os.makedirs(dirname)
        torch.save(model.state_dict(), save_path)

    def restore(self, save_path, model=None, map_location=None):
        if model is None:
            model = self.alg.model
        model.load_state_dict(torch.load(save_path, map_location=map_location))

    def train(self):
        self.alg.model.train()
        self.training = True

    def eval(self):
        self.alg.model.eval()
        self.training = False
