"""PolicyGradient (REINFORCE) — parl/algorithms/torch/policy_gradient.py:27-75: the model outputs
action probabilities; loss = mean(-log p_a * reward); Adam."""
import torch

from ..core import Algorithm
from ..engine.optim import FlatAdam
from ..utils.misc import check_model_method
from .. import kernels
from ._common import to_device_tensor, ensure_cuda

__all__ = ['PolicyGradient']


class PolicyGradient(Algorithm):
    def __init__(self, model, lr):
        check_model_method(model, 'forward', self.__class__.__name__)
        assert isinstance(lr, float)
        super(PolicyGradient, self).__init__(model)
        self.device = ensure_cuda(model, 'PolicyGradient')
        self.lr = lr
        self.optimizer = FlatAdam(model.parameters(), lr=lr)

    def predict(self, obs):
        with torch.no_grad():
            return self.model(to_device_tensor(obs, self.device))

    def learn(self, obs, action, reward):
        dev = self.device
        prob = self.model(to_device_tensor(obs, dev))
        action = to_device_tensor(action, dev)
        if action.dtype not in (torch.int32, torch.int64):
            action = action.to(torch.int64)
        reward = to_device_tensor(reward, dev, torch.float32).reshape(-1)
        res = kernels.pg_loss_fwd_bwd(prob.detach().float().contiguous(), action.reshape(-1), reward)
        torch.autograd.backward([prob], [res['d_prob'].to(prob.dtype)])
        self.optimizer.step()
        return res['losses'][0]
