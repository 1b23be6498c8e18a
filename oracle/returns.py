"""GAE / discounted-return scans — CPU restatement (TEST INFRASTRUCTURE).

calc_gae / calc_discount_sum_rewards follow parl/utils/rl_utils.py:21-51
(scipy.signal.lfilter on the reversed sequence, float64); segmentation and
bootstrap follow benchmark/torch/a2c/actor.py:82-102.  compute_returns follows
benchmark/torch/ppo/storage.py:45-64 (float32, masked with dones[t+1]).
Pinned by running the reference functions themselves (tests/golden/make_golden.py).
"""
import numpy as np
import scipy.signal


def calc_discount_sum_rewards(rewards, gamma):
    return scipy.signal.lfilter([1.0], [1.0, -gamma], np.asarray(rewards)[::-1])[::-1]   # rl_utils.py:31


def calc_gae(rewards, values, next_value, gamma, lam):
    rewards = np.asarray(rewards)
    values = np.asarray(values)
    tds = rewards + gamma * np.append(values[1:], next_value) - values                 # rl_utils.py:49
    return calc_discount_sum_rewards(tds, gamma * lam)


def a2c_segment_gae_time_major(rewards, values, dones, bootstrap_value, gamma, lam):
    """(T,B) restatement of the per-segment GAE the A2C actor performs
    (benchmark/torch/a2c/actor.py:82-102): a segment ends at done (next_value=0)
    or at the rollout end (next_value = V(next_obs) = bootstrap_value[b]).
    Returns advantages, target_values as float64 [T,B] (the reference casts to
    float32 only in the agent, benchmark/torch/a2c/atari_agent.py:59-60)."""
    rewards = np.asarray(rewards, np.float64)
    values = np.asarray(values, np.float64)
    dones = np.asarray(dones).astype(bool)
    T, B = rewards.shape
    adv = np.zeros((T, B), np.float64)
    for b in range(B):
        start = 0
        for t in range(T):
            if dones[t, b] or t == T - 1:
                nv = 0.0 if dones[t, b] else float(bootstrap_value[b])
                adv[start:t + 1, b] = calc_gae(rewards[start:t + 1, b], values[start:t + 1, b], nv, gamma, lam)
                start = t + 1
    return adv, adv + values


def compute_returns(rewards, values, dones, value, done, gamma=0.99, gae_lambda=0.95):
    """RolloutStorage.compute_returns (benchmark/torch/ppo/storage.py:45-64), float32.
    dones[t] is the done flag observed BEFORE step t (pre-step convention)."""
    rewards = np.asarray(rewards, np.float32)
    values = np.asarray(values, np.float32)
    dones = np.asarray(dones, np.float32)
    value = np.asarray(value, np.float32)
    done = np.asarray(done, np.float32)
    T = rewards.shape[0]
    advantages = np.zeros_like(rewards)
    lastgaelam = 0
    for t in reversed(range(T)):
        if t == T - 1:
            nextnonterminal = 1.0 - done
            nextvalues = value.reshape(1, -1)
        else:
            nextnonterminal = 1.0 - dones[t + 1]
            nextvalues = values[t + 1]
        delta = rewards[t] + gamma * nextvalues * nextnonterminal - values[t]
        advantages[t] = lastgaelam = delta + gamma * gae_lambda * nextnonterminal * lastgaelam
    returns = advantages + values
    return advantages, returns
