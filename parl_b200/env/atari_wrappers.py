"""``parl.env.atari_wrappers`` — DeepMind-style Atari preprocessing for HOST gym-API envs with the semantics of
parl/env/atari_wrappers.py:30-385 (MonitorEnv, NoopResetEnv, ClipRewardEnv, FireResetEnv, EpisodicLifeEnv,
MaxAndSkipEnv, WarpFrame, FrameStack, TestEnv, wrap_deepmind, get_wrapper_by_cls).  This is the real-env bridge's
host half and what unmodified examples import; the synthetic device envs implement the frame stack /
episode statistics inside their step kernel instead.  Pinned by tests/golden/wrap_deepmind.npz (recorded from the
reference chain on a scripted env)."""
import time
from collections import deque

import numpy as np

from .compat_wrappers import CompatWrapper, Wrapper, is_gym_version_ge, V_NPRANDOM_CHANGED

__all__ = ['wrap_deepmind', 'MonitorEnv', 'get_wrapper_by_cls', 'NoopResetEnv', 'ClipRewardEnv', 'FireResetEnv',
           'EpisodicLifeEnv', 'MaxAndSkipEnv', 'WarpFrame', 'FrameStack', 'TestEnv']


def get_wrapper_by_cls(env, cls):
    """The wrapper of class ``cls`` in the chain around ``env``, or None."""
    while True:
        if isinstance(env, cls):
            return env
        if not isinstance(env, Wrapper):
            return None
        env = env.env


class MonitorEnv(Wrapper):
    """Raw-episode statistics recorded below EpisodicLifeEnv & co (what Actor.get_metrics reads)."""

    def __init__(self, env=None):
        Wrapper.__init__(self, env)
        self._current_reward = None
        self._num_steps = None
        self._total_steps = None
        self._episode_rewards, self._episode_lengths = [], []
        self._num_episodes = 0
        self._num_returned = 0
        self.tstart = time.time()

    def reset(self, **kwargs):
        obs = self.env.reset(**kwargs)
        if self._total_steps is None:
            self._total_steps = sum(self._episode_lengths)
        if self._current_reward is not None:          # close the episode that just ended
            self._episode_rewards.append(self._current_reward)
            self._episode_lengths.append(self._num_steps)
            self._num_episodes += 1
        self._current_reward, self._num_steps = 0, 0
        return obs

    def step(self, action):
        frame, rew, over, info = self.env.step(action)
        self._num_steps, self._total_steps = self._num_steps + 1, self._total_steps + 1
        self._current_reward += rew
        if over:
            assert isinstance(info, dict)
            info['episode'] = dict(r=self._current_reward, l=self._num_steps, t=round(time.time() - self.tstart, 6))
        return frame, rew, over, info

    def get_episode_rewards(self):
        return self._episode_rewards

    def get_episode_lengths(self):
        return self._episode_lengths

    def get_total_steps(self):
        return self._total_steps

    def next_episode_results(self):
        for i in range(self._num_returned, len(self._episode_rewards)):
            yield (self._episode_rewards[i], self._episode_lengths[i])
        self._num_returned = len(self._episode_rewards)


class NoopResetEnv(Wrapper):
    """1..noop_max no-op steps (action 0) after every reset; a done inside them resets again."""

    def __init__(self, env, noop_max=30):
        Wrapper.__init__(self, env)
        self.noop_max = noop_max
        self.override_num_noops = None
        self.noop_action = 0
        assert env.unwrapped.get_action_meanings()[0] == 'NOOP'

    def reset(self, **kwargs):
        self.env.reset(**kwargs)
        if self.override_num_noops is not None:
            noops = self.override_num_noops
        else:
            rng = self.unwrapped.np_random
            noops = rng.integers(1, self.noop_max + 1) if is_gym_version_ge(V_NPRANDOM_CHANGED) \
                else rng.randint(1, self.noop_max + 1)
        assert noops > 0
        obs = None
        for _ in range(noops):
            obs, _, done, _ = self.env.step(self.noop_action)
            if done:
                obs = self.env.reset(**kwargs)
        return obs

    def step(self, ac):
        return self.env.step(ac)


class ClipRewardEnv(Wrapper):
    """reward -> sign(reward)."""

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        frame, r, over, info = self.env.step(action)
        return frame, self.reward(r), over, info

    def reward(self, reward):
        return np.sign(reward)


class FireResetEnv(Wrapper):
    """Press FIRE (then action 2) after reset for games that wait for it."""

    def __init__(self, env):
        Wrapper.__init__(self, env)
        meanings = env.unwrapped.get_action_meanings()
        assert meanings[1] == 'FIRE' and len(meanings) >= 3

    def reset(self, **kwargs):
        self.env.reset(**kwargs)
        obs, _, done, _ = self.env.step(1)
        if done:
            self.env.reset(**kwargs)
        obs, _, done, _ = self.env.step(2)
        if done:
            self.env.reset(**kwargs)
        return obs

    def step(self, ac):
        return self.env.step(ac)


class EpisodicLifeEnv(Wrapper):
    """A lost life ends the (learner-visible) episode; the game is only reset on a real game over."""

    def __init__(self, env):
        Wrapper.__init__(self, env)
        self.lives = 0
        self.was_real_done = True

    def step(self, action):
        frame, r, over, info = self.env.step(action)
        self.was_real_done = over
        remaining = self.env.unwrapped.ale.lives()
        life_lost = 0 < remaining < self.lives
        self.lives = remaining
        return frame, r, (True if life_lost else over), info

    def reset(self, **kwargs):
        if self.was_real_done:
            obs = self.env.reset(**kwargs)
        else:
            obs, _, _, _ = self.env.step(0)           # no-op step out of the lost-life state
        self.lives = self.env.unwrapped.ale.lives()
        return obs


class MaxAndSkipEnv(Wrapper):
    """Repeat the action ``skip`` times, sum the rewards, return the pixel-wise max of the last two frames."""

    def __init__(self, env, skip=4):
        Wrapper.__init__(self, env)
        self._obs_buffer = np.zeros((2, ) + tuple(env.observation_space.shape), dtype=np.uint8)
        self._skip = skip

    def step(self, action):
        acc, over, info = 0.0, None, None
        for i in range(self._skip):
            frame, r, over, info = self.env.step(action)
            slot = i - (self._skip - 2)               # the last two repeats land in slots 0 and 1
            if slot >= 0:
                self._obs_buffer[slot] = frame
            acc += r
            if over:
                break
        return self._obs_buffer.max(axis=0), acc, over, info

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)


class WarpFrame(Wrapper):
    """RGB -> gray -> dim x dim (cv2 INTER_AREA)."""

    def __init__(self, env, dim):
        Wrapper.__init__(self, env)
        self.width = self.height = dim
        self.observation_space = _Box(0, 255, (dim, dim), np.uint8)

    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        frame, r, over, info = self.env.step(action)
        return self.observation(frame), r, over, info

    def observation(self, frame):
        import cv2
        frame = cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY)
        return cv2.resize(frame, (self.width, self.height), interpolation=cv2.INTER_AREA)


class _Box(object):
    """Shape/dtype carrier for observation_space when gym is not installed."""

    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype


class FrameStack(Wrapper):
    """Stack of the k most recent frames ('NHWC' or 'NCHW'); reset fills all k slots with the first frame."""

    def __init__(self, env, k, obs_format='NHWC'):
        Wrapper.__init__(self, env)
        assert obs_format in ('NHWC', 'NCHW')
        self.k, self.obs_format = k, obs_format
        self.frames = deque([], maxlen=k)
        shp = tuple(env.observation_space.shape)
        shape = (shp[0], shp[1], k) if obs_format == 'NHWC' else (k, shp[0], shp[1])
        self.observation_space = _Box(0, 255, shape, env.observation_space.dtype)

    def reset(self):
        ob = self.env.reset()
        for _ in range(self.k):
            self.frames.append(ob)
        return self._get_ob()

    def step(self, action):
        frame, r, over, info = self.env.step(action)
        self.frames.append(frame)
        return self._get_ob(), r, over, info

    def _get_ob(self):
        assert len(self.frames) == self.k
        return np.stack(self.frames, axis=2) if self.obs_format == 'NHWC' else np.array(self.frames)


class TestEnv(Wrapper):
    """Evaluation helper: after ``test_episodes`` raw episodes exposes their rewards and a real-done flag."""
    __test__ = False                                  # not a pytest class

    def __init__(self, env, test_episodes=3):
        Wrapper.__init__(self, env)
        self._env = env
        self._monitor = get_wrapper_by_cls(env, MonitorEnv)
        self._test_episodes = test_episodes
        self._was_real_done = False
        self._eval_rewards = None
        self._end_episode = len(self._monitor.get_episode_rewards()) + test_episodes

    def step(self, action):
        return self._env.step(action)

    def reset(self, **kwargs):
        obs = self._env.reset(**kwargs)
        if len(self._monitor.get_episode_rewards()) >= self._end_episode:
            self._was_real_done = True
            self._eval_rewards = self._monitor.get_episode_rewards()[-self._test_episodes:]
            self._end_episode += self._test_episodes
        else:
            self._was_real_done = False
            self._eval_rewards = None
        return obs

    def get_eval_rewards(self):
        return self._eval_rewards

    def get_real_done(self):
        return self._was_real_done


def wrap_deepmind(env, dim=84, framestack=True, obs_format='NHWC', test=False, test_episodes=3):
    """The reference's wrapper chain, in its order (atari_wrappers.py:356-385)."""
    env = CompatWrapper(env)
    env = MonitorEnv(env)
    env = NoopResetEnv(env, noop_max=30)
    if 'NoFrameskip' in env.spec.id:
        env = MaxAndSkipEnv(env, skip=4)
    env = EpisodicLifeEnv(env)
    if 'FIRE' in env.unwrapped.get_action_meanings():
        env = FireResetEnv(env)
    env = WarpFrame(env, dim)
    env = ClipRewardEnv(env)
    if framestack:
        env = FrameStack(env, 4, obs_format)
    if test:
        env = TestEnv(env, test_episodes)
    return env
