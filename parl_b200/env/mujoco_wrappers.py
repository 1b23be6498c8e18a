"""``parl.env.mujoco_wrappers`` — continuous-control env wrappers for HOST gym-API envs with the semantics of
parl/env/mujoco_wrappers.py:26-240: TimeLimitMaskEnv, MonitorEnv (info['episode']), RunningMeanStd (parallel-variance
merge), VecNormalizeEnv (observation / return normalisation, clip 10), wrap_rms, get_ob_rms.  The device twin of
VecNormalizeEnv for the on-device envs is rl_vecnormalize_step (parl_b200.env.DeviceVecNormalize).  Pinned by
tests/golden/vecnormalize.npz (recorded from the reference class)."""
import time

import numpy as np

from .compat_wrappers import CompatWrapper, Wrapper

__all__ = ['wrap_rms', 'get_ob_rms', 'VecNormalizeEnv', 'RunningMeanStd', 'MonitorEnv', 'TimeLimitMaskEnv',
           'get_wrapper_by_cls', 'update_mean_var_count_from_moments']


class TimeLimitMaskEnv(Wrapper):
    """Marks a done caused by the time limit with info['bad_transition']."""

    def step(self, action):
        obs, rew, done, info = self.env.step(action)
        if done and self.env._max_episode_steps == self.env._elapsed_steps:
            info['bad_transition'] = True
        return obs, rew, done, info

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)


class MonitorEnv(Wrapper):
    """Raw episode reward / length reported through info['episode'] when the episode ends."""

    def __init__(self, env):
        Wrapper.__init__(self, env)
        self.tstart = time.time()
        self.rewards = None

    def step(self, action):
        ob, rew, done, info = self.env.step(action)
        self.update(ob, rew, done, info)
        return ob, rew, done, info

    def update(self, ob, rew, done, info):
        self.rewards.append(rew)
        if done:
            assert isinstance(info, dict)
            info['episode'] = {'r': round(sum(self.rewards), 6), 'l': len(self.rewards),
                               't': round(time.time() - self.tstart, 6)}
            self.reset()

    def reset(self, **kwargs):
        self.rewards = []
        return self.env.reset(**kwargs)


def update_mean_var_count_from_moments(mean, var, count, batch_mean, batch_var, batch_count):
    """Chan et al. parallel merge of (mean, var, count) with a batch's moments."""
    delta = batch_mean - mean
    tot_count = count + batch_count
    new_mean = mean + delta * batch_count / tot_count
    m2 = var * count + batch_var * batch_count + np.square(delta) * count * batch_count / tot_count
    return new_mean, m2 / tot_count, tot_count


class RunningMeanStd(object):
    def __init__(self, epsilon=1e-4, shape=()):
        self.mean = np.zeros(shape, 'float64')
        self.var = np.ones(shape, 'float64')
        self.count = epsilon

    def update(self, x):
        self.update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        self.mean, self.var, self.count = update_mean_var_count_from_moments(self.mean, self.var, self.count, batch_mean,
                                                                             batch_var, batch_count)


class VecNormalizeEnv(Wrapper):
    """Normalises observations with running mean/variance and rewards with the running variance of the discounted
    return; both clipped (defaults +-10)."""

    def __init__(self, env, ob=True, ret=True, clipob=10., cliprew=10., gamma=0.99, epsilon=1e-8):
        Wrapper.__init__(self, env)
        self.ob_rms = RunningMeanStd(shape=env.observation_space.shape[0]) if ob else None
        self.ret_rms = RunningMeanStd(shape=()) if ret else None
        self.clipob, self.cliprew, self.gamma, self.epsilon = clipob, cliprew, gamma, epsilon
        self.ret = np.zeros(1)
        self.training = True

    def step(self, action):
        ob, rew, new, info = self.env.step(action)
        self.ret = self.ret * self.gamma + rew
        ob = self._obfilt(ob)
        if self.ret_rms:
            self.ret_rms.update(self.ret)
            rew = np.clip(rew / np.sqrt(self.ret_rms.var + self.epsilon), -self.cliprew, self.cliprew)
        if new:
            self.ret = np.zeros(1)
        return ob, rew, new, info

    def reset(self):
        self.ret = np.zeros(1)
        return self._obfilt(self.env.reset())

    def _obfilt(self, ob, update=True):
        if not self.ob_rms:
            return ob
        if ob.ndim == 1:
            ob = np.expand_dims(ob, 0)
        if self.training and update:
            self.ob_rms.update(ob)
        ob = np.clip((ob - self.ob_rms.mean) / np.sqrt(self.ob_rms.var + self.epsilon), -self.clipob, self.clipob)
        return np.squeeze(ob, axis=0) if ob.shape[0] == 1 else ob

    def get_ob_rms(self):
        return self.ob_rms

    def set_ob_rms(self, ob_rms):
        self.ob_rms = ob_rms

    def train(self):
        self.training = True

    def eval(self):
        self.training = False


def get_wrapper_by_cls(venv, cls):
    if isinstance(venv, cls):
        return venv
    if hasattr(venv, 'env'):
        return get_wrapper_by_cls(venv.env, cls)
    return None


def get_ob_rms(env):
    vec = get_wrapper_by_cls(env, VecNormalizeEnv)
    return vec.get_ob_rms() if vec else None


def wrap_rms(env, gamma, test=False, ob_rms=None):
    """CompatWrapper -> TimeLimitMaskEnv -> MonitorEnv -> VecNormalizeEnv (evaluation: frozen statistics, no return
    normalisation), mujoco_wrappers.py:222-240."""
    env = MonitorEnv(TimeLimitMaskEnv(CompatWrapper(env)))
    if test:
        env = VecNormalizeEnv(env, ret=False)
        env.eval()
        env.set_ob_rms(ob_rms)
    else:
        env = VecNormalizeEnv(env, gamma=gamma)
    return env
