"""PPO — signature and semantics of parl/algorithms/torch/ppo.py:27-196 (type-asserted ctor,
MEAN losses, adv-norm with unbiased std, clipped surrogate + clipped value loss,
clip_grad_norm_(max_grad_norm), Adam(eps))."""
import torch

from ..core import Algorithm
from ..engine.optim import FlatAdam
from ..utils.misc import check_model_method
from .. import kernels
from ._common import to_device_tensor, ensure_cuda

__all__ = ['PPO']


class PPO(Algorithm):
    def __init__(self, model, clip_param=0.1, value_loss_coef=0.5, entropy_coef=0.01, initial_lr=2.5e-4, eps=1e-5,
                 max_grad_norm=0.5, use_clipped_value_loss=True, norm_adv=True, continuous_action=False):
        check_model_method(model, 'value', self.__class__.__name__)
        check_model_method(model, 'policy', self.__class__.__name__)
        assert isinstance(clip_param, float)
        assert isinstance(value_loss_coef, float)
        assert isinstance(entropy_coef, float)
        assert isinstance(initial_lr, float)
        assert isinstance(eps, float)
        assert isinstance(max_grad_norm, float)
        assert isinstance(use_clipped_value_loss, bool)
        assert isinstance(norm_adv, bool)
        assert isinstance(continuous_action, bool)
        super(PPO, self).__init__(model)
        self.clip_param = clip_param
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self.norm_adv = norm_adv
        self.continuous_action = continuous_action
        self.device = ensure_cuda(model, 'PPO')
        self.optimizer = FlatAdam(model.parameters(), lr=initial_lr, eps=eps, clip='torch', max_norm=max_grad_norm)
        self.grad_sync = None
        self.adv_stats_sync = None     # multi-GPU: callable(batch_adv) -> device {mean, 1/(std+1e-8)}
        self.seed = 0
        self._sample_step = 0

    def _policy_outputs(self, obs):
        out = self.model.policy(obs)
        if self.continuous_action:
            mean, std = out
            # the reference models hand back std = exp(logstd) expanded over the batch
            # (benchmark/torch/ppo/mujoco_model.py:46-53); the loss kernel takes the [D] log-std.
            logstd_param = getattr(self.model, 'fc_pi_std', None)
            return mean, std, logstd_param
        return out, None, None

    def learn(self, batch_obs, batch_action, batch_value, batch_return, batch_logprob, batch_adv, lr=None):
        dev = self.device
        obs = to_device_tensor(batch_obs, dev)
        f32 = torch.float32
        bv, br = to_device_tensor(batch_value, dev, f32), to_device_tensor(batch_return, dev, f32)
        blp, badv = to_device_tensor(batch_logprob, dev, f32), to_device_tensor(batch_adv, dev, f32)
        values = self.model.value(obs)
        vflat = values.detach().float().contiguous().reshape(-1)
        stats = None
        if self.norm_adv:
            stats = self.adv_stats_sync(badv) if self.adv_stats_sync is not None else kernels.adv_stats(badv)
        kw = dict(clip_param=self.clip_param, value_loss_coef=self.value_loss_coef, entropy_coef=self.entropy_coef,
                  use_clipped_value_loss=self.use_clipped_value_loss, norm_adv=self.norm_adv, stats=stats)
        if self.continuous_action:
            mean, std = self.model.policy(obs)
            self._check_state_independent_std(std)
            act = to_device_tensor(batch_action, dev, f32)
            std_row = std.detach().float().reshape(-1, std.shape[-1])[0].contiguous()
            logstd = torch.log(std_row)
            res = kernels.ppo_loss_fwd_bwd(vflat, act.reshape(mean.shape), bv, br, blp, badv,
                                           mean=mean.detach().float().contiguous(), logstd=logstd, **kw)
            # d loss / d std (per row) so that autograd reaches however the model parameterises std:
            # d/dlogstd summed over the batch is delivered on row 0 of the expanded std.
            d_std = torch.zeros_like(std)
            d_std.reshape(-1, std.shape[-1])[0] = (res['d_logstd'] / std_row).to(std.dtype)
            torch.autograd.backward([values, mean, std],
                                    [res['d_values'].view_as(values).to(values.dtype), res['d_mean'].to(mean.dtype), d_std])
        else:
            logits = self.model.policy(obs)
            act = to_device_tensor(batch_action, dev)
            if act.dtype not in (torch.int32, torch.int64):
                act = act.to(torch.int64)
            res = kernels.ppo_loss_fwd_bwd(vflat, act.reshape(-1), bv, br, blp, badv,
                                           logits=logits.detach().float().contiguous(), **kw)
            torch.autograd.backward([values, logits],
                                    [res['d_values'].view_as(values).to(values.dtype), res['d_logits'].to(logits.dtype)])
        if self.grad_sync is not None:
            self.grad_sync(self.optimizer.grad)
        self.optimizer.step(lr=float(lr) if lr else None)
        L = res['losses'].tolist()          # the reference returns Python floats (.item(), ppo.py:149)
        return L[0], L[1], L[2]

    def _check_state_independent_std(self, std):
        """The fused Gaussian loss / sampler take ONE log-std vector [D] (the reference models' ``fc_pi_std``
        parameter expanded over the batch, benchmark/torch/ppo/mujoco_model.py:46-53).  A model whose std
        depends on the state would silently get wrong log-probs: checked once, loudly."""
        if getattr(self, '_std_checked', False):
            return
        s2 = std.detach().reshape(-1, std.shape[-1])
        if s2.shape[0] > 1 and not bool((s2 == s2[0:1]).all()):
            raise ValueError('PPO(continuous_action=True): model.policy() returned a state-dependent std; '
                             'parl_b200 supports the reference form std = exp(logstd parameter) only')
        self._std_checked = True

    def sample(self, obs):
        """ppo.py:151-179 -> (value, action, action_log_probs, action_entropy)."""
        with torch.no_grad():
            obs = to_device_tensor(obs, self.device)
            value = self.model.value(obs)
            step = self._sample_step
            self._sample_step += 1
            if self.continuous_action:
                mean, std = self.model.policy(obs)
                self._check_state_independent_std(std)
                logstd = torch.log(std.float().reshape(-1, std.shape[-1])[0]).contiguous()
                action, logp = kernels.sample_gaussian(mean.float().contiguous(), logstd, self.seed, step)
                entropy = (0.5 + 0.5 * 1.8378770664093453 + logstd).sum().expand(mean.shape[0])
            else:
                logits = self.model.policy(obs).float().contiguous()
                action, logp = kernels.sample_categorical(logits, self.seed, step, want_logp=True)
                action = action.long()
                lsm = torch.log_softmax(logits, -1)
                entropy = -(lsm.exp() * lsm).sum(-1)
            return value, action, logp, entropy

    def predict(self, obs):
        with torch.no_grad():
            obs = to_device_tensor(obs, self.device)
            if self.continuous_action:
                action, _ = self.model.policy(obs)
                return action
            return self.model.policy(obs).argmax(dim=-1, keepdim=True)

    def value(self, obs):
        with torch.no_grad():
            return self.model.value(to_device_tensor(obs, self.device))
