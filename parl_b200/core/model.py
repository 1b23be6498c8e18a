"""``parl.Model`` on PyTorch — host-side mirror of parl/core/torch/model.py:24-134.

Parameters stay torch tensors (on the B200); ``get_weights`` / ``set_weights`` keep the
reference's numpy-dict contract (one entry per ``state_dict`` key, in order) and
``sync_weights_to`` keeps ``target = decay*target + (1-decay)*self``
(behaviours pinned by parl/core/torch/tests/model_base_test_torch.py:53-335).
"""
import numpy as np
import torch
import torch.nn as nn

__all__ = ['Model']


class Model(nn.Module):
    def sync_weights_to(self, target_model, decay=0.0):
        assert target_model is not self, "cannot copy between identical model"
        assert isinstance(target_model, Model)
        assert self.__class__.__name__ == target_model.__class__.__name__, \
            "must be the same class for params syncing!"
        assert 0 <= decay <= 1
        targets = dict(target_model.named_parameters())
        with torch.no_grad():
            for name, src in self.named_parameters():
                dst = targets[name]
                if decay == 0.0:
                    dst.copy_(src)
                else:
                    dst.mul_(decay).add_(src.detach().to(dst.device), alpha=1.0 - decay)

    def get_weights(self):
        return {k: v.detach().cpu().numpy() for k, v in self.state_dict().items()}

    def set_weights(self, weights):
        if not isinstance(weights, dict):
            raise TypeError('set_weights expects the dict returned by get_weights(), got %s' % type(weights).__name__)
        self.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in weights.items()})
