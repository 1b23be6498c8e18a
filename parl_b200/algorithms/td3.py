"""TD3 — signature and semantics of parl/algorithms/torch/td3.py:25-112: clipped Gaussian noise on the target
policy's action, min of the twin target critics, delayed actor / target updates (every ``policy_freq`` calls).
Networks on torch autograd through the user's ``parl.Model``; TD target, the two mse terms and their gradients from
rl_twin_q_td_loss_fwd_bwd; Adam from rl_adam_step."""
import copy

import torch

from ..core import Algorithm
from ..engine.optim import FlatAdam
from ..utils.misc import check_model_method
from .. import kernels
from ._common import to_device_tensor, ensure_cuda

__all__ = ['TD3']


class TD3(Algorithm):
    def __init__(self, model, gamma=None, tau=None, actor_lr=None, critic_lr=None, policy_noise=0.2, noise_clip=0.5,
                 policy_freq=2):
        for m in ('value', 'policy', 'Q1', 'get_actor_params', 'get_critic_params'):
            check_model_method(model, m, self.__class__.__name__)
        assert isinstance(gamma, float)
        assert isinstance(tau, float)
        assert isinstance(actor_lr, float)
        assert isinstance(critic_lr, float)
        super(TD3, self).__init__(model)
        self.device = ensure_cuda(model, self.__class__.__name__)
        self.gamma, self.tau, self.actor_lr, self.critic_lr = gamma, tau, actor_lr, critic_lr
        self.policy_noise, self.noise_clip, self.policy_freq = policy_noise, noise_clip, policy_freq
        self.target_model = copy.deepcopy(model)
        self.actor_optimizer = FlatAdam(model.get_actor_params(), lr=actor_lr)
        self.critic_optimizer = FlatAdam(model.get_critic_params(), lr=critic_lr)
        self.total_it = 0

    def predict(self, obs):
        with torch.no_grad():
            return self.model.policy(to_device_tensor(obs, self.device, torch.float32))

    def learn(self, obs, action, reward, next_obs, terminal):
        dev, f32 = self.device, torch.float32
        obs, action = to_device_tensor(obs, dev, f32), to_device_tensor(action, dev, f32)
        next_obs = to_device_tensor(next_obs, dev, f32)
        reward = to_device_tensor(reward, dev, f32).reshape(-1)
        terminal = to_device_tensor(terminal, dev, f32).reshape(-1)
        self.total_it += 1
        self._critic_learn(obs, action, reward, next_obs, terminal)
        if self.total_it % self.policy_freq == 0:
            self._actor_learn(obs)

    def _critic_learn(self, obs, action, reward, next_obs, terminal):
        with torch.no_grad():                                                   # td3.py:80-88
            noise = (torch.randn_like(action) * self.policy_noise).clamp(-self.noise_clip, self.noise_clip)
            next_action = (self.target_model.policy(next_obs) + noise).clamp(-1, 1)
            tq1, tq2 = self.target_model.value(next_obs, next_action)
            tq1, tq2 = tq1.float().reshape(-1).contiguous(), tq2.float().reshape(-1).contiguous()
        q1, q2 = self.model.value(obs, action)
        res = kernels.twin_q_td_loss_fwd_bwd(q1.detach().float().reshape(-1).contiguous(), tq1, reward, terminal,
                                             self.gamma, q2=q2.detach().float().reshape(-1).contiguous(),
                                             q2_target_next=tq2)
        self.critic_optimizer.zero_grad()
        torch.autograd.backward([q1, q2], [res['d_q1'].view_as(q1).to(q1.dtype), res['d_q2'].view_as(q2).to(q2.dtype)])
        self.critic_optimizer.step()
        return res['losses'][0]

    def _actor_learn(self, obs):
        q = self.model.Q1(obs, self.model.policy(obs))                          # td3.py:98
        self.actor_optimizer.zero_grad()
        torch.autograd.backward([q], [torch.full_like(q, -1.0 / q.numel())])
        self.actor_optimizer.step()
        self.sync_target()
        return -q.detach().mean()

    def sync_target(self, decay=None):
        if decay is None:
            decay = 1.0 - self.tau
        self.model.sync_weights_to(self.target_model, decay=decay)
