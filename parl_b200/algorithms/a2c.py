"""A2C — signature and semantics of parl/algorithms/torch/a2c.py:26-91 (config dict with
'vf_loss_coeff' and 'learning_rate'; SUM losses; clip_grad_norm_(40); Adam, lr per call)."""
import torch

from ..core import Algorithm
from ..engine.optim import FlatAdam
from ..utils.misc import check_model_method
from .. import kernels
from ._common import to_device_tensor, ensure_cuda

__all__ = ['A2C']


class A2C(Algorithm):
    def __init__(self, model, config=None, vf_loss_coeff=None):
        if config is None:            # paddle-style constructor: A2C(model, vf_loss_coeff=...)  (paddle/a2c.py:26)
            config = {'vf_loss_coeff': vf_loss_coeff, 'learning_rate': 0.001}
        assert isinstance(config['vf_loss_coeff'], (int, float))
        check_model_method(model, 'value', self.__class__.__name__)
        check_model_method(model, 'policy', self.__class__.__name__)
        check_model_method(model, 'policy_and_value', self.__class__.__name__)
        super(A2C, self).__init__(model)
        self.vf_loss_coeff = config['vf_loss_coeff']
        self.config = config
        self.device = ensure_cuda(model, 'A2C')
        self.optimizer = FlatAdam(model.parameters(), lr=config['learning_rate'], clip='torch', max_norm=40.0)
        self.grad_sync = None
        self._sample_step = 0
        self.seed = int(config.get('seed', 0)) if hasattr(config, 'get') else 0

    def learn(self, obs, actions, advantages, target_values, lr, entropy_coeff):
        dev = self.device
        obs = to_device_tensor(obs, dev)
        actions = to_device_tensor(actions, dev)
        if actions.dtype not in (torch.int32, torch.int64):
            actions = actions.to(torch.int64)
        advantages = to_device_tensor(advantages, dev, torch.float32)
        target_values = to_device_tensor(target_values, dev, torch.float32)
        logits, values = self.model.policy_and_value(obs)
        res = kernels.a2c_loss_fwd_bwd(logits.detach().float().contiguous(), values.detach().float().contiguous(),
                                       actions.reshape(-1), advantages.reshape(-1), target_values.reshape(-1),
                                       self.vf_loss_coeff, entropy_coeff)
        torch.autograd.backward([logits, values], [res['d_logits'].to(logits.dtype), res['d_values'].to(values.dtype)])
        if self.grad_sync is not None:
            self.grad_sync(self.optimizer.grad)
        self.optimizer.step(lr=lr)
        L = res['losses']
        return L[0], L[1], L[2], L[3]

    def sample(self, obs):
        """a2c.py:73-76 -> (sample_actions int64, values); Categorical sampling by the exact
        inverse-CDF kernel with the Philox action stream."""
        with torch.no_grad():
            logits, values = self.model.policy_and_value(to_device_tensor(obs, self.device))
            acts = kernels.sample_categorical(logits.float().contiguous(), self.seed, self._sample_step)
            self._sample_step += 1
            return acts.long(), values

    def prob_and_value(self, obs):
        with torch.no_grad():
            logits, values = self.model.policy_and_value(to_device_tensor(obs, self.device))
            return torch.softmax(logits, dim=1), values

    def predict(self, obs):
        with torch.no_grad():
            return self.model.policy(to_device_tensor(obs, self.device)).max(-1)[1]

    def value(self, obs):
        with torch.no_grad():
            return self.model.value(to_device_tensor(obs, self.device))
