"""VecNormalizeEnv arithmetic — CPU restatement (TEST INFRASTRUCTURE).

Follows parl/env/mujoco_wrappers.py:95-168 (VecNormalizeEnv.step / reset / _obfilt), :74-92 (RunningMeanStd) and
:191-217 (update_mean_var_count_from_moments) for a batch of B independent env instances, each with its OWN
running statistics fed one sample at a time (benchmark/torch/ppo/env_utils.py wraps every env separately), float64.
Pinned by tests/golden/vecnormalize.npz, recorded from the reference class itself (tests/golden/make_golden_env.py).
"""
import numpy as np


class VecNormState(object):
    def __init__(self, B, D, clipob=10.0, cliprew=10.0, gamma=0.99, epsilon=1e-8):
        self.ob_mean = np.zeros((B, D), np.float64)
        self.ob_var = np.ones((B, D), np.float64)
        self.ob_count = np.full(B, 1e-4, np.float64)
        self.ret_mean = np.zeros(B, np.float64)
        self.ret_var = np.ones(B, np.float64)
        self.ret_count = np.full(B, 1e-4, np.float64)
        self.ret = np.zeros(B, np.float64)
        self.clipob, self.cliprew, self.gamma, self.epsilon = clipob, cliprew, gamma, epsilon


def _update(mean, var, count, x):
    """update_from_moments with batch_mean = x, batch_var = 0, batch_count = 1 (mujoco_wrappers.py:191-217)."""
    delta = x - mean
    tot = count + 1.0
    new_mean = mean + delta * 1.0 / tot
    m2 = var * count + 0.0 + np.square(delta) * count * 1.0 / tot
    return new_mean, m2 / tot, tot


def obfilt(st, ob, update=True):
    """_obfilt on a [B, D] batch of per-env observations."""
    if update:
        st.ob_mean, st.ob_var, cnt = _update(st.ob_mean, st.ob_var, st.ob_count[:, None], ob.astype(np.float64))
        st.ob_count = cnt[:, 0]
    return np.clip((ob - st.ob_mean) / np.sqrt(st.ob_var + st.epsilon), -st.clipob, st.clipob)


def step(st, ob, rew, done, terminal_ob=None, update=True):
    """One VecNormalizeEnv.step per env followed, for finished envs, by the reset the vector env performs
    (benchmark/torch/ppo/env_utils.py:96-104): `ob` is the observation handed on (the reset observation where done),
    `terminal_ob` (optional) the observation the finished episode's last step returned — the reference filters it
    too (and so updates the statistics with it) before the reset observation."""
    rew = np.asarray(rew, np.float64)
    done = np.asarray(done).astype(bool)
    st.ret = st.ret * st.gamma + rew
    if terminal_ob is not None and update:
        t = np.where(done[:, None], terminal_ob, st.ob_mean)           # a sample equal to the mean ...
        m, v, c = _update(st.ob_mean, st.ob_var, st.ob_count[:, None], t.astype(np.float64))
        st.ob_mean = np.where(done[:, None], m, st.ob_mean)            # ... is only applied where done
        st.ob_var = np.where(done[:, None], v, st.ob_var)
        st.ob_count = np.where(done, c[:, 0], st.ob_count)
    st.ret_mean, st.ret_var, st.ret_count = _update(st.ret_mean, st.ret_var, st.ret_count, st.ret)
    rew_n = np.clip(rew / np.sqrt(st.ret_var + st.epsilon), -st.cliprew, st.cliprew)
    st.ret = np.where(done, 0.0, st.ret)
    return obfilt(st, ob, update), rew_n
