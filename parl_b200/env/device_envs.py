"""Device-resident vector envs: thousands of env instances stepped in lock-step by one kernel
(rl_env_*_step), with VectorEnv's auto-reset contract (parl/env/vector_env.py:53-63) and the mock-gym
distributions (parl/tests/gym.py:117-207).  ``reset()`` / ``step(actions)`` return CUDA tensors."""
import torch

from .. import kernels


class _DeviceVecBase(object):
    def __init__(self, num_envs, seed, device, env_offset, ring_cap):
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError('device vector envs need a CUDA device (no CPU fallback)')
            device = torch.device('cuda', torch.cuda.current_device())
        self.device = torch.device(device)
        self.envs_num = self.num_envs = int(num_envs)
        self.seed, self.env_offset = int(seed), int(env_offset)
        self.stats = kernels.EpisodeStats(self.num_envs, self.device, ring_cap)
        self.step_count = 0
        self.reward = torch.zeros(self.num_envs, dtype=torch.float32, device=self.device)
        self.done = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        self._metrics_read = 0

    def next_episode_results(self):
        """(return, length) of episodes completed since the last call — the device analogue of
        MonitorEnv.next_episode_results() used by Actor.get_metrics (examples/IMPALA/actor.py:93-102)."""
        head = int(self.stats.ring_head.item())
        cap = self.stats.ring_cap
        lo = max(self._metrics_read, head - cap)
        idx = [i % cap for i in range(lo, head)]
        self._metrics_read = head
        if not idx:
            return []
        rets = self.stats.ring_ret.cpu()[idx].tolist()
        lens = self.stats.ring_len.cpu()[idx].tolist()
        return list(zip(rets, lens))


class AtariSynthVectorEnv(_DeviceVecBase):
    """Atari-shaped synthetic env: 84x84 uint8 frames ~ U{0..254}, 4-frame stack (NCHW), reward in
    {0,1}, done p=0.1 — the Pong mock (parl/tests/gym.py:137-180) behind FrameStack(4)."""

    def __init__(self, num_envs, seed=0, frame_hw=(84, 84), p_done=0.1, n_actions=18, device=None, env_offset=0,
                 ring_cap=4096):
        super(AtariSynthVectorEnv, self).__init__(num_envs, seed, device, env_offset, ring_cap)
        self.h, self.w = frame_hw
        self.hw = self.h * self.w
        self.p_done, self.n_actions = p_done, n_actions
        self.observation_shape = (4, self.h, self.w)
        # 8-plane ring: obs at step t uses planes t..t+3 (mod 8 handled by copying the 3 carried planes)
        self.planes = torch.zeros((8, self.num_envs, self.hw), dtype=torch.uint8, device=self.device)
        self.ages = torch.zeros((5, self.num_envs), dtype=torch.uint8, device=self.device)
        self._t = 0

    def _obs(self):
        out = torch.empty((self.num_envs, 4, self.h, self.w), dtype=torch.uint8, device=self.device)
        kernels.obs_stack_gather(self.planes, self.ages, self._t, 1, out)
        return out

    def reset(self):
        self._t = 0
        self.step_count = 0
        kernels.env_atari_synth_step(self.planes[3], None, None, None, self.ages[0], self.stats, self.seed, 0,
                                     env_offset=self.env_offset, reset=True)
        return self._obs()

    def step(self, actions=None):
        if self._t == 4:                          # roll the window back: planes 4..7 -> 0..3
            self.planes[0:4].copy_(self.planes[4:8].clone())
            self.ages[0].copy_(self.ages[4])
            self._t = 0
        t = self._t
        kernels.env_atari_synth_step(self.planes[t + 4], self.reward, self.done, self.ages[t], self.ages[t + 1],
                                     self.stats, self.seed, self.step_count, p_done=self.p_done,
                                     env_offset=self.env_offset)
        self._t += 1
        self.step_count += 1
        return self._obs(), self.reward.clone(), self.done.bool(), {}


class MujocoSynthVectorEnv(_DeviceVecBase):
    """MuJoCo-shaped synthetic env: obs ~ N(0,1)^17, act dim 6, reward in {0,1}, done p=0.01
    (HalfCheetah mock, parl/tests/gym.py:183-207)."""

    def __init__(self, num_envs, seed=0, obs_dim=17, act_dim=6, p_done=0.01, max_episode_steps=0, device=None,
                 env_offset=0, ring_cap=4096):
        super(MujocoSynthVectorEnv, self).__init__(num_envs, seed, device, env_offset, ring_cap)
        self.obs_dim, self.act_dim, self.p_done, self.max_episode_steps = obs_dim, act_dim, p_done, max_episode_steps
        self.observation_shape = (obs_dim, )
        self.obs = torch.zeros((self.num_envs, obs_dim), dtype=torch.float32, device=self.device)

    def reset(self):
        self.step_count = 0
        kernels.env_mujoco_synth_step(self.obs, None, None, self.stats, self.seed, 0, env_offset=self.env_offset,
                                      reset=True)
        return self.obs.clone()

    def step(self, actions=None):
        kernels.env_mujoco_synth_step(self.obs, self.reward, self.done, self.stats, self.seed, self.step_count,
                                      p_done=self.p_done, max_episode_steps=self.max_episode_steps,
                                      env_offset=self.env_offset)
        self.step_count += 1
        return self.obs.clone(), self.reward.clone(), self.done.bool(), {}


class CartPoleVectorEnv(_DeviceVecBase):
    """CartPole-v0/v1 physics on the device (gym classic-control constants)."""

    def __init__(self, num_envs, seed=0, max_episode_steps=200, device=None, env_offset=0, ring_cap=4096):
        super(CartPoleVectorEnv, self).__init__(num_envs, seed, device, env_offset, ring_cap)
        self.max_episode_steps = max_episode_steps
        self.observation_shape = (4, )
        self.n_actions = 2
        self.state = torch.zeros((self.num_envs, 4), dtype=torch.float32, device=self.device)
        self.obs = torch.zeros((self.num_envs, 4), dtype=torch.float32, device=self.device)

    def reset(self):
        self.step_count = 0
        kernels.env_cartpole_step(self.state, self.obs, None, None, None, self.stats, self.seed, 0,
                                  max_episode_steps=self.max_episode_steps, env_offset=self.env_offset, reset=True)
        return self.obs.clone()

    def step(self, actions):
        actions = torch.as_tensor(actions).to(self.device, torch.int32).contiguous()
        kernels.env_cartpole_step(self.state, self.obs, self.reward, self.done, actions, self.stats, self.seed,
                                  self.step_count, max_episode_steps=self.max_episode_steps,
                                  env_offset=self.env_offset)
        self.step_count += 1
        return self.obs.clone(), self.reward.clone(), self.done.bool(), {}
