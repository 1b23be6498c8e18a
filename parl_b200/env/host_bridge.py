"""Real-env bridge: HOST gym-API envs feeding the DEVICE learner through pinned staging.

``HostEnvBridge(envs)`` steps a list of host envs with VectorEnv's auto-reset contract (parl/env/vector_env.py:41-63,
benchmark/torch/ppo/env_utils.py:68-117), writes observations / rewards / dones into pinned host arrays and hands the
device copies (async H2D on the current stream) to the algorithms' ``sample`` / the rollout storages, and brings the
sampled actions back through a pinned D2H buffer — one copy each way per step instead of per-env tensor
construction.  The synthetic device envs never need it; this is the path for ALE / MuJoCo processes."""
import numpy as np
import torch

__all__ = ['HostEnvBridge']


class HostEnvBridge(object):
    def __init__(self, envs, obs_shape, obs_dtype=np.float32, device=None):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError('HostEnvBridge stages into HBM: no CUDA device visible')
            device = torch.device('cuda', torch.cuda.current_device())
        self.envs, self.env_num, self.device = list(envs), len(envs), torch.device(device)
        tdt = torch.from_numpy(np.zeros(1, obs_dtype)).dtype
        n = self.env_num
        self.h_obs = torch.empty((n, ) + tuple(obs_shape), dtype=tdt, pin_memory=True)
        self.h_rew = torch.empty(n, dtype=torch.float32, pin_memory=True)
        self.h_done = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        self.d_obs = torch.empty_like(self.h_obs, device=self.device)
        self.d_rew = torch.empty(n, dtype=torch.float32, device=self.device)
        self.d_done = torch.empty(n, dtype=torch.uint8, device=self.device)
        self._np_obs, self._np_rew, self._np_done = self.h_obs.numpy(), self.h_rew.numpy(), self.h_done.numpy()
        self._h_act = None
        self.infos = [None] * n

    def _upload(self):
        self.d_obs.copy_(self.h_obs, non_blocking=True)
        self.d_rew.copy_(self.h_rew, non_blocking=True)
        self.d_done.copy_(self.h_done, non_blocking=True)
        return self.d_obs, self.d_rew, self.d_done

    def reset(self):
        for i, env in enumerate(self.envs):
            self._np_obs[i] = env.reset()
        self._np_rew[:] = 0
        self._np_done[:] = 0
        return self._upload()[0]

    def actions_to_host(self, actions):
        """Device action tensor -> numpy (one D2H through a pinned buffer, synchronised)."""
        if self._h_act is None or self._h_act.shape != actions.shape or self._h_act.dtype != actions.dtype:
            self._h_act = torch.empty(actions.shape, dtype=actions.dtype, pin_memory=True)
        self._h_act.copy_(actions, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._h_act.numpy()

    def step(self, actions):
        """actions: device tensor or numpy [env_num, ...] -> (obs, reward, done) DEVICE tensors; a finished env is
        reset and its reset observation handed on with done = 1 (vector_env.py:54-62)."""
        acts = self.actions_to_host(actions) if torch.is_tensor(actions) else np.asarray(actions)
        for i, env in enumerate(self.envs):
            ob, r, d, info = env.step(acts[i])
            if d:
                ob = env.reset()
            self._np_obs[i], self._np_rew[i], self._np_done[i], self.infos[i] = ob, r, 1 if d else 0, info
        return self._upload()
