"""DQN — signature and semantics of parl/algorithms/torch/dqn.py:29-75 (1-step TD target with the
target network, MSE mean, Adam; ``sync_target`` copies online -> target)."""
import copy

import torch

from ..core import Algorithm
from ..engine.optim import FlatAdam
from ..utils.misc import check_model_method
from .. import kernels
from ._common import to_device_tensor, ensure_cuda

__all__ = ['DQN']


class DQN(Algorithm):
    double_q = False

    def __init__(self, model, gamma=None, lr=None):
        check_model_method(model, 'forward', self.__class__.__name__)
        assert isinstance(gamma, float)
        assert isinstance(lr, float)
        super(DQN, self).__init__(model)
        self.device = ensure_cuda(model, self.__class__.__name__)
        self.target_model = copy.deepcopy(model)
        self.gamma = gamma
        self.lr = lr
        self.optimizer = FlatAdam(model.parameters(), lr=lr)
        self.grad_sync = None

    def predict(self, obs):
        with torch.no_grad():
            return self.model(to_device_tensor(obs, self.device))

    def learn(self, obs, action, reward, next_obs, terminal, sample_weight=None):
        """Returns the loss as a Python float (dqn.py:72).  With ``sample_weight`` (PER importance
        weights, per_alg.py:48-69) returns (loss, td_abs device tensor)."""
        dev = self.device
        obs, next_obs = to_device_tensor(obs, dev), to_device_tensor(next_obs, dev)
        action = to_device_tensor(action, dev)
        if action.dtype not in (torch.int32, torch.int64):
            action = action.to(torch.int64)
        reward = to_device_tensor(reward, dev, torch.float32).reshape(-1)
        terminal = to_device_tensor(terminal, dev, torch.float32).reshape(-1)
        q = self.model(obs)
        with torch.no_grad():
            q_tgt = self.target_model(next_obs).float().contiguous()
            q_onl = self.model(next_obs).float().contiguous() if self.double_q else None
        w = to_device_tensor(sample_weight, dev, torch.float32).reshape(-1) if sample_weight is not None else None
        res = kernels.td_loss_fwd_bwd(q.detach().float().contiguous(), q_tgt, action.reshape(-1), reward, terminal,
                                      self.gamma, q_online_next=q_onl, weights=w, want_td_abs=w is not None)
        torch.autograd.backward([q], [res['d_q'].to(q.dtype)])
        if self.grad_sync is not None:
            self.grad_sync(self.optimizer.grad)
        self.optimizer.step()
        if w is not None:
            return res['losses'][0], res['td_abs']
        return res['losses'].item()

    def sync_target(self):
        self.model.sync_weights_to(self.target_model)
