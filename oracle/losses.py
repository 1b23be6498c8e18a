"""A2C / PPO / DQN / DDQN / PolicyGradient / PER losses as functions of the
network OUTPUTS — CPU restatement in torch float32 (TEST INFRASTRUCTURE).

Each function restates the post-network part of the reference ``learn`` and
returns the scalar losses plus autograd gradients w.r.t. the network outputs,
which is what the fused CUDA kernels emit.
  A2C   parl/algorithms/torch/a2c.py:40-60
  PPO   parl/algorithms/torch/ppo.py:102-138
  DQN   parl/algorithms/torch/dqn.py:61-69 ; DDQN parl/algorithms/torch/ddqn.py:61-72
  PG    parl/algorithms/torch/policy_gradient.py:54-75
  PER   benchmark/fluid/Prioritized_DQN/per_alg.py:48-69
Pinned by tests/golden/make_golden.py, which runs the reference ``learn``
methods on identity-style models and records losses / parameter gradients.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch.distributions import Categorical, Normal


def _t(x, dtype=torch.float32):
    return torch.as_tensor(np.asarray(x)).to(dtype)


def a2c_loss(logits, values, actions, advantages, target_values, vf_loss_coeff, entropy_coeff):
    lg = _t(logits).clone().requires_grad_(True)
    v = _t(values).clone().requires_grad_(True)
    a = _t(actions, torch.int64)
    adv, tv = _t(advantages), _t(target_values)
    onehot = F.one_hot(a, lg.shape[-1])
    logp = torch.sum(F.log_softmax(lg, dim=1) * onehot, dim=-1)       # a2c.py:45-46
    pi_loss = -1.0 * torch.sum(logp * adv)                            # :48
    vf_loss = 0.5 * torch.sum(torch.square(v - tv))                   # :52-53
    entropy = torch.sum(Categorical(logits=lg).entropy())             # :55-58
    total = pi_loss + vf_loss * vf_loss_coeff + entropy * entropy_coeff
    total.backward()
    return dict(total_loss=total.item(), pi_loss=pi_loss.item(), vf_loss=vf_loss.item(),
                entropy=entropy.item(), d_logits=lg.grad.numpy(), d_values=v.grad.numpy())


def ppo_loss(values, batch_action, batch_value, batch_return, batch_logprob, batch_adv,
             logits=None, mean=None, logstd=None, clip_param=0.1, value_loss_coef=0.5,
             entropy_coef=0.01, use_clipped_value_loss=True, norm_adv=True):
    """ppo.py:102-138. Discrete: pass logits. Continuous: pass mean [M,D] and
    logstd [D] (std = exp(logstd) expanded, benchmark/torch/ppo/mujoco_model.py:46-53)."""
    v = _t(values).clone().requires_grad_(True)
    out = {}
    if logits is None:
        mu = _t(mean).clone().requires_grad_(True)
        ls = _t(logstd).clone().requires_grad_(True)
        dist = Normal(mu, ls.exp().expand_as(mu))
        act = _t(batch_action)
        logp = dist.log_prob(act).sum(1)
        ent = dist.entropy().sum(1)
    else:
        lg = _t(logits).clone().requires_grad_(True)
        dist = Categorical(logits=lg)
        logp = dist.log_prob(_t(batch_action, torch.int64))
        ent = dist.entropy()
    entropy_loss = ent.mean()
    adv = _t(batch_adv)
    if norm_adv:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)                 # ppo.py:115-117
    old_lp, ret, old_v = _t(batch_logprob), _t(batch_return), _t(batch_value)
    ratio = torch.exp(logp - old_lp)
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, 1.0 - clip_param, 1.0 + clip_param) * adv
    action_loss = -torch.min(surr1, surr2).mean()
    vv = v.view(-1)
    if use_clipped_value_loss:
        vclip = old_v + torch.clamp(vv - old_v, -clip_param, clip_param)
        value_loss = 0.5 * torch.max((vv - ret).pow(2), (vclip - ret).pow(2)).mean()
    else:
        value_loss = 0.5 * (ret - vv).pow(2).mean()
    loss = value_loss * value_loss_coef + action_loss - entropy_loss * entropy_coef
    loss.backward()
    out.update(value_loss=value_loss.item(), action_loss=action_loss.item(),
               entropy_loss=entropy_loss.item(), loss=loss.item(), d_values=v.grad.numpy())
    if logits is None:
        out.update(d_mean=mu.grad.numpy(), d_logstd=ls.grad.numpy())
    else:
        out.update(d_logits=lg.grad.numpy())
    return out


def td_loss(q, q_target_next, action, reward, terminal, gamma, q_online_next=None, weights=None):
    """DQN (dqn.py:64-69), DDQN when q_online_next is given (ddqn.py:64-72),
    PER-weighted when weights is given (per_alg.py:56-66)."""
    qq = _t(q).clone().requires_grad_(True)
    qt = _t(q_target_next)
    a = _t(action, torch.int64).view(-1, 1)
    r = _t(reward).view(-1, 1)
    term = _t(terminal).view(-1, 1)
    pred = qq.gather(1, a)
    with torch.no_grad():
        if q_online_next is None:
            max_v = qt.max(1, keepdim=True)[0]
        else:
            greedy = _t(q_online_next).max(dim=1, keepdim=True)[1]
            max_v = qt.gather(1, greedy)
        target = r + (1 - term) * gamma * max_v
    if weights is None:
        loss = F.mse_loss(pred, target)
    else:
        loss = (_t(weights).view(-1, 1) * (pred - target) ** 2).mean()
    loss.backward()
    return dict(loss=loss.item(), d_q=qq.grad.numpy(), target=target.numpy().reshape(-1),
                td_abs=(target - pred).abs().detach().numpy().reshape(-1))


def pg_loss(prob, action, reward):
    """policy_gradient.py:54-75 (model outputs probabilities)."""
    p = _t(prob).clone().requires_grad_(True)
    logp = Categorical(p).log_prob(_t(action, torch.int64))
    loss = torch.mean(-1 * logp * _t(reward))
    loss.backward()
    return dict(loss=loss.item(), d_prob=p.grad.numpy())


def twin_q_td(q1, q1_target_next, reward, terminal, gamma, q2=None, q2_target_next=None, next_log_prob=None, alpha=0.0):
    """Continuous-control critic TD in float32 numpy, operation order of the reference:
    DDPG ddpg.py:63-73 (single critic), TD3 td3.py:86-94 (min of the twin target critics), SAC sac.py:92-99
    (min - alpha * log pi, then reward + gamma * (1 - terminal) * target).  Gradients of the mean-squared errors
    are written out by hand: d mse / d q = 2 (q - target) / N."""
    f = np.float32
    q1 = np.asarray(q1, f).reshape(-1)
    tq = np.asarray(q1_target_next, f).reshape(-1)
    if q2_target_next is not None:
        tq = np.minimum(tq, np.asarray(q2_target_next, f).reshape(-1))
    if next_log_prob is not None:
        tq = (tq - f(alpha) * np.asarray(next_log_prob, f).reshape(-1)).astype(f)
    r, term = np.asarray(reward, f).reshape(-1), np.asarray(terminal, f).reshape(-1)
    target = (r + ((f(1.0) - term) * f(gamma)).astype(f) * tq).astype(f)
    n = q1.size
    d1 = q1 - target
    out = dict(target=target, mse1=float(np.mean(d1.astype(np.float64) ** 2)), d_q1=(d1 * f(2.0 / n)).astype(f))
    out['loss'] = out['mse1']
    if q2 is not None:
        d2 = np.asarray(q2, f).reshape(-1) - target
        out.update(mse2=float(np.mean(d2.astype(np.float64) ** 2)), d_q2=(d2 * f(2.0 / n)).astype(f))
        out['loss'] = out['mse1'] + out['mse2']
    return out
