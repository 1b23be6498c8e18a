"""DDPG — signature and semantics of parl/algorithms/torch/ddpg.py:25-99: critic TD on the target networks
(mse to r + (1 - terminal) * gamma * Q'(s', pi'(s'))), actor loss -Q(s, pi(s)).mean(), Polyak target update.
The networks run through the user's ``parl.Model`` (torch autograd); TD target, loss and the gradient w.r.t. Q come
from rl_twin_q_td_loss_fwd_bwd, both Adam steps from rl_adam_step (one flat buffer per optimizer)."""
import copy

import torch

from ..core import Algorithm
from ..engine.optim import FlatAdam
from ..utils.misc import check_model_method
from .. import kernels
from ._common import to_device_tensor, ensure_cuda

__all__ = ['DDPG']


class DDPG(Algorithm):
    def __init__(self, model, gamma=None, tau=None, actor_lr=None, critic_lr=None):
        for m in ('value', 'policy', 'get_actor_params', 'get_critic_params'):
            check_model_method(model, m, self.__class__.__name__)
        assert isinstance(gamma, float)
        assert isinstance(tau, float)
        assert isinstance(actor_lr, float)
        assert isinstance(critic_lr, float)
        super(DDPG, self).__init__(model)
        self.device = ensure_cuda(model, self.__class__.__name__)
        self.gamma, self.tau, self.actor_lr, self.critic_lr = gamma, tau, actor_lr, critic_lr
        self.target_model = copy.deepcopy(model)
        self.actor_optimizer = FlatAdam(model.get_actor_params(), lr=actor_lr)
        self.critic_optimizer = FlatAdam(model.get_critic_params(), lr=critic_lr)

    def predict(self, obs):
        with torch.no_grad():
            return self.model.policy(to_device_tensor(obs, self.device, torch.float32))

    def _batch(self, obs, action, reward, next_obs, terminal):
        dev, f32 = self.device, torch.float32
        return (to_device_tensor(obs, dev, f32), to_device_tensor(action, dev, f32),
                to_device_tensor(reward, dev, f32).reshape(-1), to_device_tensor(next_obs, dev, f32),
                to_device_tensor(terminal, dev, f32).reshape(-1))

    def learn(self, obs, action, reward, next_obs, terminal):
        obs, action, reward, next_obs, terminal = self._batch(obs, action, reward, next_obs, terminal)
        critic_loss = self._critic_learn(obs, action, reward, next_obs, terminal)
        actor_loss = self._actor_learn(obs)
        self.sync_target()
        return critic_loss, actor_loss

    def _critic_learn(self, obs, action, reward, next_obs, terminal):
        with torch.no_grad():                                                   # ddpg.py:65-67
            tq = self.target_model.value(next_obs, self.target_model.policy(next_obs)).float().reshape(-1).contiguous()
        q = self.model.value(obs, action)
        res = kernels.twin_q_td_loss_fwd_bwd(q.detach().float().reshape(-1).contiguous(), tq, reward, terminal, self.gamma)
        self.critic_optimizer.zero_grad()
        torch.autograd.backward([q], [res['d_q1'].view_as(q).to(q.dtype)])
        self.critic_optimizer.step()
        return res['losses'][0]

    def _actor_learn(self, obs):
        q = self.model.value(obs, self.model.policy(obs))                       # ddpg.py:87
        self.actor_optimizer.zero_grad()
        torch.autograd.backward([q], [torch.full_like(q, -1.0 / q.numel())])    # d(-mean q)/dq
        self.actor_optimizer.step()
        return -q.detach().mean()

    def sync_target(self, decay=None):
        if decay is None:
            decay = 1.0 - self.tau
        self.model.sync_weights_to(self.target_model, decay=decay)
