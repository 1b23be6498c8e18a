"""Atari frame-ring ``ReplayMemory`` in HBM — interface of benchmark/torch/dqn/replay_memory.py:22-113
(``ReplayMemory(max_size, obs_shape, context_len)``, ``append(Experience)``, ``recent_obs()``, ``sample_batch(n)``,
``size()``): single uint8 frames in a ring, the (context_len+1)-frame window and the episode-boundary zeroing are
assembled at sample time by rl_replay_gather_frames.  One transition stream (the reference's layout); the
lock-stepped multi-env engine uses the lane-interleaved ``parl_b200.engine.dqn.DeviceAtariReplay``."""
from collections import deque, namedtuple

import numpy as np
import torch

from .. import kernels

__all__ = ['ReplayMemory', 'Experience']

Experience = namedtuple('Experience', ['obs', 'action', 'reward', 'isOver'])


class ReplayMemory(object):
    def __init__(self, max_size, obs_shape, context_len, device=None):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError('parl_b200 Atari ReplayMemory lives in HBM: no CUDA device visible (no CPU fallback)')
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = dev = torch.device(device)
        self.max_size, self.obs_shape, self.context_len = int(max_size), tuple(obs_shape), int(context_len)
        self.hw = int(np.prod(self.obs_shape))
        assert self.hw % 16 == 0, 'frame bytes must be a multiple of 16 (84x84 = 7056 is)'
        self.obs = torch.zeros((self.max_size, self.hw), dtype=torch.uint8, device=dev)
        self.action = torch.zeros(self.max_size, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(self.max_size, dtype=torch.float32, device=dev)
        self.isOver = torch.zeros(self.max_size, dtype=torch.uint8, device=dev)
        self._curr_size = 0
        self._curr_pos = 0
        self._context = deque(maxlen=self.context_len - 1)

    def append(self, exp):
        p = self._curr_pos
        self.obs[p].copy_(torch.as_tensor(np.ascontiguousarray(exp.obs, dtype=np.uint8)).reshape(-1), non_blocking=True)
        self.action[p], self.reward[p], self.isOver[p] = int(exp.action), float(exp.reward), int(bool(exp.isOver))
        self._curr_size = min(self._curr_size + 1, self.max_size)
        self._curr_pos = (p + 1) % self.max_size
        if exp.isOver:
            self._context.clear()
        else:
            self._context.append(exp)

    def recent_obs(self):
        lst = list(self._context)
        pad = [np.zeros(self.obs_shape, dtype='uint8')] * (self._context.maxlen - len(lst))
        return pad + [k.obs for k in lst]

    def size(self):
        return self._curr_size

    __len__ = size

    def sample_batch_by_index(self, batch_idx, as_numpy=True):
        """batch of START indices (as replay_memory.py:103 builds them) -> [obs, action, reward, isOver]."""
        idx = torch.as_tensor(np.asarray(batch_idx)).to(self.device, torch.int32).contiguous()
        obs = kernels.replay_gather_frames(self.obs, self.isOver, idx, self._curr_size, self.context_len)
        obs = obs.view((idx.numel(), self.context_len + 1) + self.obs_shape)
        real = ((idx + (self.context_len - 1)) % self._curr_size).long()
        act, rew, over = self.action[real], self.reward[real], self.isOver[real].bool()
        if as_numpy:
            return [obs.cpu().numpy(), act.cpu().numpy().astype('int8'), rew.cpu().numpy(), over.cpu().numpy()]
        return [obs, act, rew, over]

    def sample_batch(self, batch_size, as_numpy=True):
        raw = np.random.randint(self._curr_size - self.context_len - 1, size=batch_size)
        return self.sample_batch_by_index((self._curr_pos + raw) % self._curr_size, as_numpy=as_numpy)
