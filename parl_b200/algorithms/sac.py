"""SAC — signature and semantics of parl/algorithms/torch/sac.py:25-126: tanh-squashed Gaussian policy with the
reparameterisation trick, soft twin-critic TD target min(Q1', Q2') - alpha log pi(a'|s'), actor loss
(alpha log pi - min Q).mean(), Polyak target update after every call.  Networks on torch autograd through the user's
``parl.Model``; the critic's TD target / loss / gradient from rl_twin_q_td_loss_fwd_bwd; Adam from rl_adam_step."""
import copy

import torch

from ..core import Algorithm
from ..engine.optim import FlatAdam
from ..utils.misc import check_model_method
from .. import kernels
from ._common import to_device_tensor, ensure_cuda

__all__ = ['SAC']


class SAC(Algorithm):
    def __init__(self, model, gamma=None, tau=None, alpha=None, actor_lr=None, critic_lr=None):
        for m in ('value', 'policy', 'get_actor_params', 'get_critic_params'):
            check_model_method(model, m, self.__class__.__name__)
        assert isinstance(gamma, float)
        assert isinstance(tau, float)
        assert isinstance(alpha, float)
        assert isinstance(actor_lr, float)
        assert isinstance(critic_lr, float)
        super(SAC, self).__init__(model)
        self.device = ensure_cuda(model, self.__class__.__name__)
        self.gamma, self.tau, self.alpha, self.actor_lr, self.critic_lr = gamma, tau, alpha, actor_lr, critic_lr
        self.target_model = copy.deepcopy(model)
        self.actor_optimizer = FlatAdam(model.get_actor_params(), lr=actor_lr)
        self.critic_optimizer = FlatAdam(model.get_critic_params(), lr=critic_lr)
        self.noise_fn = None        # test hook: callable(mean) -> eps tensor instead of torch.randn_like(mean)

    def predict(self, obs):
        with torch.no_grad():
            act_mean, _ = self.model.policy(to_device_tensor(obs, self.device, torch.float32))
            return torch.tanh(act_mean)

    def sample(self, obs):
        """(action, log_prob [N,1]) with the reparameterisation trick and the tanh correction (sac.py:73-85)."""
        obs = to_device_tensor(obs, self.device, torch.float32)
        act_mean, act_log_std = self.model.policy(obs)
        std = act_log_std.exp()
        eps = self.noise_fn(act_mean) if self.noise_fn is not None else torch.randn_like(act_mean)
        x_t = act_mean + std * eps                                              # Normal.rsample
        action = torch.tanh(x_t)
        # Normal(mean, std).log_prob(x_t) in the reference's operation order (torch/distributions/normal.py)
        var = std ** 2
        log_prob = -((x_t - act_mean) ** 2) / (2 * var) - act_log_std.exp().log() - 0.9189385332046727
        log_prob = log_prob - torch.log((1 - action.pow(2)) + 1e-6)
        return action, log_prob.sum(1, keepdim=True)

    def learn(self, obs, action, reward, next_obs, terminal):
        dev, f32 = self.device, torch.float32
        obs, action = to_device_tensor(obs, dev, f32), to_device_tensor(action, dev, f32)
        next_obs = to_device_tensor(next_obs, dev, f32)
        reward = to_device_tensor(reward, dev, f32).reshape(-1)
        terminal = to_device_tensor(terminal, dev, f32).reshape(-1)
        critic_loss = self._critic_learn(obs, action, reward, next_obs, terminal)
        actor_loss = self._actor_learn(obs)
        self.sync_target()
        return critic_loss, actor_loss

    def _critic_learn(self, obs, action, reward, next_obs, terminal):
        with torch.no_grad():                                                   # sac.py:91-95
            next_action, next_logp = self.sample(next_obs)
            tq1, tq2 = self.target_model.value(next_obs, next_action)
            tq1, tq2 = tq1.float().reshape(-1).contiguous(), tq2.float().reshape(-1).contiguous()
            next_logp = next_logp.float().reshape(-1).contiguous()
        q1, q2 = self.model.value(obs, action)
        res = kernels.twin_q_td_loss_fwd_bwd(q1.detach().float().reshape(-1).contiguous(), tq1, reward, terminal,
                                             self.gamma, q2=q2.detach().float().reshape(-1).contiguous(),
                                             q2_target_next=tq2, next_log_prob=next_logp, alpha=self.alpha)
        self.critic_optimizer.zero_grad()
        torch.autograd.backward([q1, q2], [res['d_q1'].view_as(q1).to(q1.dtype), res['d_q2'].view_as(q2).to(q2.dtype)])
        self.critic_optimizer.step()
        return res['losses'][0]

    def _actor_learn(self, obs):
        act, log_pi = self.sample(obs)                                          # sac.py:108-112
        q1_pi, q2_pi = self.model.value(obs, act)
        actor_loss = ((self.alpha * log_pi) - torch.min(q1_pi, q2_pi)).mean()
        self.actor_optimizer.zero_grad()
        actor_loss.backward()
        self.actor_optimizer.step()
        return actor_loss.detach()

    def sync_target(self, decay=None):
        if decay is None:
            decay = 1.0 - self.tau
        self.model.sync_weights_to(self.target_model, decay=decay)
