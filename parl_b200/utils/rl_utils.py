"""``parl.utils.calc_gae`` / ``calc_discount_sum_rewards`` (parl/utils/rl_utils.py:21-51) for host
callers that hold a Python list of floats (the QuickStart / A2C actor pattern).  The recurrence is
evaluated in float64 exactly like scipy.signal.lfilter([1],[1,-g]) on the reversed sequence; the
(T,B) device path is kernels.gae_scan_segments."""
import numpy as np

__all__ = ['calc_discount_sum_rewards', 'calc_gae']


def calc_discount_sum_rewards(rewards, gamma):
    x = np.asarray(rewards, dtype=np.float64)
    out = np.empty_like(x)
    acc = 0.0
    for i in range(len(x) - 1, -1, -1):
        acc = x[i] + gamma * acc
        out[i] = acc
    return out


def calc_gae(rewards, values, next_value, gamma, lam):
    rewards = np.asarray(rewards, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    tds = rewards + gamma * np.append(values[1:], next_value) - values
    return calc_discount_sum_rewards(tds, gamma * lam)
