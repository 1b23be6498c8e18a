"""IMPALA — signature and semantics of parl/algorithms/paddle/impala/impala.py:82-227.

``learn`` = network forward (user Model, autograd) -> ONE fused kernel for log-softmax / entropy /
KL / V-trace scan / losses / d(loss)/d(logits, values) (rl_vtrace_loss_fwd_bwd) -> autograd
backward through the network -> global-norm clip 40 (paddle rule) + Adam in two launches.
"""
import collections

import torch

from ..core import Algorithm
from ..engine.optim import FlatAdam
from ..utils.misc import check_model_method
from .. import kernels
from ._common import to_device_tensor, ensure_cuda

__all__ = ['IMPALA']

VTraceLoss = collections.namedtuple('VTraceLoss', ['total_loss', 'pi_loss', 'vf_loss', 'entropy'])


class IMPALA(Algorithm):
    def __init__(self, model, sample_batch_steps=None, gamma=None, vf_loss_coeff=None, clip_rho_threshold=None,
                 clip_pg_rho_threshold=None):
        assert isinstance(sample_batch_steps, int)
        assert isinstance(gamma, float)
        assert isinstance(vf_loss_coeff, float)
        assert isinstance(clip_rho_threshold, float)
        assert isinstance(clip_pg_rho_threshold, float)
        check_model_method(model, 'policy', self.__class__.__name__)
        check_model_method(model, 'value', self.__class__.__name__)
        super(IMPALA, self).__init__(model)
        self.sample_batch_steps = sample_batch_steps
        self.gamma = gamma
        self.vf_loss_coeff = vf_loss_coeff
        self.clip_rho_threshold = clip_rho_threshold
        self.clip_pg_rho_threshold = clip_pg_rho_threshold
        self.device = ensure_cuda(model, 'IMPALA')
        # impala.py:113-117: Adam(lr=0.001) with ClipGradByGlobalNorm(40)
        self.optimizer = FlatAdam(model.parameters(), lr=0.001, clip='paddle', max_norm=40.0)
        self.grad_sync = None          # multi-GPU: callable(flat_grad) all-reducing in place (SUM)

    def _forward(self, obs):
        if hasattr(self.model, 'policy_and_value'):
            logits, values = self.model.policy_and_value(obs)
        else:
            values = self.model.value(obs)
            logits = self.model.policy(obs)
        return logits, values

    def learn(self, obs, actions, behaviour_logits, rewards, dones, learning_rate, entropy_coeff,
              layout=kernels.ENV_MAJOR):
        """obs [B*T, ...] and friends in the reference's env-major order (impala.py:134-215); pass
        ``layout=kernels.TIME_MAJOR`` for (T,B)-ordered device rollouts."""
        dev = self.device
        obs = to_device_tensor(obs, dev)
        actions = to_device_tensor(actions, dev)
        if actions.dtype not in (torch.int32, torch.int64):
            actions = actions.to(torch.int64)
        behaviour_logits = to_device_tensor(behaviour_logits, dev, torch.float32)
        rewards = to_device_tensor(rewards, dev, torch.float32)
        dones = to_device_tensor(dones, dev)
        if dones.dtype not in (torch.bool, torch.uint8):
            dones = dones != 0
        T = self.sample_batch_steps
        N = actions.numel()
        assert N % T == 0
        B = N // T
        logits, values = self._forward(obs)
        res = kernels.vtrace_loss_fwd_bwd(
            logits.detach().float().contiguous(), behaviour_logits, actions.reshape(-1), rewards.reshape(-1),
            dones.reshape(-1), values.detach().float().contiguous().reshape(-1), T, B, self.gamma, self.vf_loss_coeff,
            entropy_coeff, self.clip_rho_threshold, self.clip_pg_rho_threshold, layout=layout)
        torch.autograd.backward([logits, values], [res['d_logits'].view_as(logits).to(logits.dtype),
                                                   res['d_values'].view_as(values).to(values.dtype)])
        if self.grad_sync is not None:
            self.grad_sync(self.optimizer.grad)
        self.optimizer.step(lr=learning_rate)
        L = res['losses']
        return VTraceLoss(total_loss=L[0], pi_loss=L[1], vf_loss=L[2], entropy=L[3]), L[4]

    def sample(self, obs):
        """impala.py:217-227: returns (probs, logits)."""
        with torch.no_grad():
            logits = self.model.policy(to_device_tensor(obs, self.device))
            return torch.softmax(logits, dim=-1), logits

    def predict(self, obs):
        with torch.no_grad():
            return self.model.policy(to_device_tensor(obs, self.device)).argmax(-1)
