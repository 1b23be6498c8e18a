"""V-trace returns and the IMPALA loss — CPU restatement (TEST INFRASTRUCTURE).

Follows parl/algorithms/paddle/impala/vtrace.py:99-139 (returns) and
parl/algorithms/paddle/impala/impala.py:25-79 (VTraceLoss), :119-132 (_log_prob),
:148-208 (IMPALA.learn post-network part).  paddle is absent from the build
image, so this restatement is pinned by the reference's own known-answer test
(vtrace_test_paddle.py:33-144, committed as tests/golden/vtrace_kat.npz).
Categorical entropy / KL follow the in-tree spec
parl/core/torch/policy_distribution.py:139-152,180-205 (softmax entropy
-sum p*logp and sum p*(logp-logq)), which is what paddle.distribution.Categorical
computes.
"""
import numpy as np
import torch
import torch.nn.functional as F


def from_importance_weights(behaviour_actions_log_probs, target_actions_log_probs,
                            discounts, rewards, values, bootstrap_value,
                            clip_rho_threshold=1.0, clip_pg_rho_threshold=1.0):
    """numpy float32, same operation order as vtrace.py:99-139."""
    f32 = np.float32
    blp = np.asarray(behaviour_actions_log_probs, f32)
    tlp = np.asarray(target_actions_log_probs, f32)
    discounts = np.asarray(discounts, f32)
    rewards = np.asarray(rewards, f32)
    values = np.asarray(values, f32)
    bootstrap_value = np.asarray(bootstrap_value, f32)
    log_rhos = tlp - blp                                           # vtrace.py:101
    rhos = np.exp(log_rhos)                                        # :103
    clipped_rhos = np.minimum(rhos, f32(clip_rho_threshold)) if clip_rho_threshold is not None else rhos
    cs = np.minimum(rhos, f32(1.0))                                # :109
    values_t_plus_1 = np.concatenate([values[1:], bootstrap_value[None]], axis=0)  # :112
    deltas = clipped_rhos * (rewards + discounts * values_t_plus_1 - values)       # :115
    acc = np.zeros_like(bootstrap_value)
    result = []
    for t in range(discounts.shape[0] - 1, -1, -1):                # :118-122
        acc = deltas[t] + discounts[t] * cs[t] * acc
        result.append(acc)
    result.reverse()
    vs_minus_v_xs = np.stack(result)
    vs = vs_minus_v_xs + values                                    # :125
    vs_t_plus_1 = np.concatenate([vs[1:], bootstrap_value[None]], axis=0)
    clipped_pg_rhos = np.minimum(rhos, f32(clip_pg_rho_threshold)) if clip_pg_rho_threshold is not None else rhos
    pg_advantages = clipped_pg_rhos * (rewards + discounts * vs_t_plus_1 - values)  # :136-137
    return vs.astype(f32), pg_advantages.astype(f32)


def ground_truth_o_t2(blp, tlp, discounts, rewards, values, bootstrap_value,
                      clip_rho_threshold, clip_pg_rho_threshold):
    """The O(T^2) definition used by the reference test (vtrace_test_paddle.py:34-76)."""
    log_rhos = tlp - blp
    seq_len = len(discounts)
    rhos = np.exp(log_rhos)
    cs = np.minimum(rhos, 1.0)
    clipped_rhos = np.minimum(rhos, clip_rho_threshold) if clip_rho_threshold else rhos
    clipped_pg_rhos = np.minimum(rhos, clip_pg_rho_threshold) if clip_pg_rho_threshold else rhos
    v_tp1 = np.concatenate([values, bootstrap_value[None, :]], axis=0)
    vs = []
    for s in range(seq_len):
        v_s = np.copy(values[s])
        for t in range(s, seq_len):
            v_s += (np.prod(discounts[s:t], axis=0) * np.prod(cs[s:t], axis=0) *
                    clipped_rhos[t] * (rewards[t] + discounts[t] * v_tp1[t + 1] - values[t]))
        vs.append(v_s)
    vs = np.stack(vs, axis=0)
    pg = clipped_pg_rhos * (rewards + discounts * np.concatenate([vs[1:], bootstrap_value[None, :]], 0) - values)
    return vs, pg


def kat_inputs(batch_size, seq_len=5):
    """Inputs of the reference known-answer test (vtrace_test_paddle.py:78-112)."""
    def ar(*shape):
        return np.arange(np.prod(shape), dtype=np.float32).reshape(*shape)
    log_rhos = ar(seq_len, batch_size) / (batch_size * seq_len)
    log_rhos = 5 * (log_rhos - 0.5)
    return dict(
        blp=np.ones_like(log_rhos, dtype=np.float32),
        tlp=(log_rhos + 1.0).astype(np.float32),
        discounts=np.array([[0.9 / (b + 1) for b in range(batch_size)] for _ in range(seq_len)], np.float32),
        rewards=ar(seq_len, batch_size),
        values=ar(seq_len, batch_size) / batch_size,
        bootstrap_value=ar(batch_size) + 1.0,
        clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2)


def impala_loss_time_major(target_logits, behaviour_logits, actions, rewards, dones, values,
                           gamma, vf_loss_coeff, entropy_coeff,
                           clip_rho_threshold=1.0, clip_pg_rho_threshold=1.0):
    """IMPALA loss on TIME-MAJOR tensors [T, B, ...] (T includes the bootstrap row).

    Mirrors impala.py:148-208 after ``split_batches``: the last row is dropped
    and values[-1] is the bootstrap.  Returns a dict with the 4 losses, kl,
    vs, pg_adv and the autograd gradients w.r.t. target_logits and values
    (zero in the dropped row).  torch CPU float32.
    """
    tl = torch.as_tensor(np.asarray(target_logits, np.float32)).clone().requires_grad_(True)
    bl = torch.as_tensor(np.asarray(behaviour_logits, np.float32))
    v = torch.as_tensor(np.asarray(values, np.float32)).clone().requires_grad_(True)
    a = torch.as_tensor(np.asarray(actions).astype(np.int64))
    r = torch.as_tensor(np.asarray(rewards, np.float32))
    d = torch.as_tensor(np.asarray(dones).astype(bool))
    A = tl.shape[-1]
    t_logp_all = F.log_softmax(tl, dim=-1)
    b_logp_all = F.log_softmax(bl, dim=-1)
    onehot = F.one_hot(a, A).to(tl.dtype)
    tlp = (t_logp_all * onehot).sum(-1)                 # impala.py:129-131
    blp = (b_logp_all * onehot).sum(-1)
    p = t_logp_all.exp()
    entropy = -(p * t_logp_all).sum(-1)                 # Categorical.entropy
    kl = (p * (t_logp_all - b_logp_all)).sum(-1).mean()  # impala.py:160-162 (all rows)
    tlp_, blp_, ent_ = tlp[:-1], blp[:-1], entropy[:-1]  # impala.py:186-194
    d_, r_ = d[:-1], r[:-1]
    boot = v[-1]
    v_ = v[:-1]
    discounts = (~d_).float() * gamma                   # impala.py:59
    vs, pg = from_importance_weights(blp_.detach().numpy(), tlp_.detach().numpy(),
                                     discounts.numpy(), r_.numpy(), v_.detach().numpy(),
                                     boot.detach().numpy(), clip_rho_threshold, clip_pg_rho_threshold)
    vs_t, pg_t = torch.as_tensor(vs), torch.as_tensor(pg)
    pi_loss = -1.0 * torch.sum(tlp_ * pg_t)             # impala.py:67-68
    vf_loss = 0.5 * torch.sum(torch.square(v_ - vs_t))  # :71-72
    ent = torch.sum(ent_)                               # :75
    total = pi_loss + vf_loss * vf_loss_coeff + ent * entropy_coeff  # :78-79
    total.backward()
    return dict(total_loss=total.item(), pi_loss=pi_loss.item(), vf_loss=vf_loss.item(),
                entropy=ent.item(), kl=kl.item(), vs=vs, pg_advantages=pg,
                d_logits=tl.grad.numpy(), d_values=v.grad.numpy())


def env_major_to_time_major(x, T):
    """[B*T, ...] with flat index b*T+t (examples/IMPALA/actor.py:79-89) -> [T, B, ...]
    exactly as ``split_batches`` (impala.py:170-175)."""
    x = np.asarray(x)
    B = x.shape[0] // T
    return np.swapaxes(x.reshape((B, T) + x.shape[1:]), 0, 1)
