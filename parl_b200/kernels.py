"""Functional torch-tensor front end of the C ABI (include/parl_b200.h).

Every function takes/returns CUDA tensors, enqueues exactly the named kernel on
the current stream and never synchronises.  No CPU fallback: CPU tensors raise.
"""
import torch

from . import _lib
from ._lib import ptr, stream, check, require_cuda

TIME_MAJOR = 0
ENV_MAJOR = 1

_workspaces = {}


def loss_workspace(device, n_cols):
    """Zero-initialised scratch for the loss kernels (partials + ticket), cached per device."""
    need = _lib.load().rl_loss_workspace_bytes(int(n_cols))
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def vtrace_from_importance_weights(behaviour_actions_log_probs, target_actions_log_probs, discounts, rewards, values,
                                   bootstrap_value, clip_rho_threshold=1.0, clip_pg_rho_threshold=1.0):
    """parl/algorithms/paddle/impala/vtrace.py:36-139 on [T,B] float32 CUDA tensors -> (vs, pg_advantages)."""
    args = [behaviour_actions_log_probs, target_actions_log_probs, discounts, rewards, values, bootstrap_value]
    require_cuda(*args)
    T, B = rewards.shape
    vs = torch.empty_like(values)
    pg = torch.empty_like(values)
    cr = -1.0 if clip_rho_threshold is None else float(clip_rho_threshold)
    cp = -1.0 if clip_pg_rho_threshold is None else float(clip_pg_rho_threshold)
    check(_lib.load().rl_vtrace_from_importance_weights(*[ptr(a) for a in args], T, B, cr, cp, ptr(vs), ptr(pg),
                                                        stream()), 'vtrace_from_importance_weights')
    return vs, pg


def vtrace_loss_fwd_bwd(target_logits, behaviour_logits, actions, rewards, dones, values, T, B, gamma,
                        vf_loss_coeff, entropy_coeff, clip_rho_threshold=1.0, clip_pg_rho_threshold=1.0,
                        layout=TIME_MAJOR, want_returns=False, out=None):
    """Fused IMPALA loss + gradient (rl_vtrace_loss_fwd_bwd).

    target_logits / behaviour_logits: [T*B, A] (or [T,B,A]) float32 in `layout` order,
    actions int32/int64, rewards f32, dones bool/uint8, values f32 — all T*B long.
    Returns dict(losses[5] device tensor = total, pi, vf, entropy, kl; d_logits; d_values; vs; pg_advantages).
    """
    require_cuda(target_logits, behaviour_logits, actions, rewards, dones, values)
    A = target_logits.shape[-1]
    assert target_logits.numel() == T * B * A and behaviour_logits.numel() == T * B * A
    assert target_logits.dtype == torch.float32 and behaviour_logits.dtype == torch.float32
    assert actions.dtype in (torch.int32, torch.int64) and actions.numel() == T * B
    if dones.dtype == torch.bool:
        dones = dones.view(torch.uint8)
    assert dones.dtype == torch.uint8 and rewards.dtype == torch.float32 and values.dtype == torch.float32
    dev = target_logits.device
    if out is None:
        out = {}
    losses = out.get('losses')
    if losses is None:
        losses = torch.empty(8, dtype=torch.float32, device=dev)
    d_logits = out.get('d_logits')
    if d_logits is None:
        d_logits = torch.empty_like(target_logits)
    d_values = out.get('d_values')
    if d_values is None:
        d_values = torch.empty_like(values)
    vs = pg = None
    if want_returns:
        vs = torch.empty((T - 1, B), dtype=torch.float32, device=dev)
        pg = torch.empty((T - 1, B), dtype=torch.float32, device=dev)
    ws = loss_workspace(dev, B)
    cr = -1.0 if clip_rho_threshold is None else float(clip_rho_threshold)
    cp = -1.0 if clip_pg_rho_threshold is None else float(clip_pg_rho_threshold)
    check(_lib.load().rl_vtrace_loss_fwd_bwd(
        ptr(target_logits), ptr(behaviour_logits), ptr(actions), 1 if actions.dtype == torch.int64 else 0,
        ptr(rewards), ptr(dones), ptr(values), T, B, A, layout, float(gamma), cr, cp, float(vf_loss_coeff),
        float(entropy_coeff), ptr(losses), ptr(d_logits), ptr(d_values), ptr(vs), ptr(pg), ptr(ws), ws.numel(),
        stream()), 'vtrace_loss_fwd_bwd')
    return dict(losses=losses, d_logits=d_logits, d_values=d_values, vs=vs, pg_advantages=pg)


# --------------------------------------------------------------------------- envs / sampling
class EpisodeStats(object):
    """Device-side episode bookkeeping shared by the env steppers."""

    def __init__(self, B, device, ring_cap=4096):
        self.ep_ret = torch.zeros(B, dtype=torch.float32, device=device)
        self.ep_len = torch.zeros(B, dtype=torch.int32, device=device)
        self.totals = torch.zeros(4, dtype=torch.float32, device=device)
        self.ring_cap = ring_cap
        self.ring_ret = torch.zeros(max(ring_cap, 1), dtype=torch.float32, device=device)
        self.ring_len = torch.zeros(max(ring_cap, 1), dtype=torch.int32, device=device)
        self.ring_head = torch.zeros(1, dtype=torch.int32, device=device)

    def args(self):
        return [ptr(self.ep_ret), ptr(self.ep_len), ptr(self.totals), ptr(self.ring_ret), ptr(self.ring_len),
                ptr(self.ring_head), self.ring_cap]


def env_atari_synth_step(frame_out, reward_out, done_out, age_in, age_out, stats, seed, step, p_done=0.1,
                         env_offset=0, logits=None, actions_out=None, reset=False):
    B, HW = frame_out.shape[0], frame_out[0].numel()
    A = logits.shape[-1] if logits is not None else 0
    check(_lib.load().rl_env_atari_synth_step(
        ptr(frame_out), ptr(reward_out), ptr(done_out), ptr(age_in), ptr(age_out), ptr(logits), A, ptr(actions_out),
        *stats.args(), B, HW, int(seed), int(step), int(env_offset), float(p_done), 1 if reset else 0, stream()),
        'env_atari_synth_step')


def obs_stack_gather(planes, ages, t_begin, t_count, out, layout=TIME_MAJOR, scale=1.0):
    """planes [P,B,HW] u8, ages [T+1,B] u8 -> out [t_count*B, 4, HW] (uint8 or float32)."""
    P, B, HW = planes.shape[0], planes.shape[1], planes[0, 0].numel()
    dt = {torch.uint8: 0, torch.float32: 1}[out.dtype]
    assert out.numel() == t_count * B * 4 * HW
    check(_lib.load().rl_obs_stack_gather(ptr(planes), ptr(ages), B, HW, int(t_begin), int(t_count), layout, dt,
                                          float(scale), ptr(out), stream()), 'obs_stack_gather')
    return out


def env_mujoco_synth_step(obs_out, reward_out, done_out, stats, seed, step, p_done=0.01, max_episode_steps=0,
                          env_offset=0, reset=False):
    B, D = obs_out.shape
    check(_lib.load().rl_env_mujoco_synth_step(
        ptr(obs_out), ptr(reward_out), ptr(done_out), *stats.args(), B, D, int(max_episode_steps), int(seed),
        int(step), int(env_offset), float(p_done), 1 if reset else 0, stream()), 'env_mujoco_synth_step')


def env_cartpole_step(state, obs_out, reward_out, done_out, actions, stats, seed, step, max_episode_steps=200,
                      env_offset=0, reset=False):
    B = state.shape[0]
    check(_lib.load().rl_env_cartpole_step(
        ptr(state), ptr(obs_out), ptr(reward_out), ptr(done_out), ptr(actions), *stats.args(), B,
        int(max_episode_steps), int(seed), int(step), int(env_offset), 1 if reset else 0, stream()),
        'env_cartpole_step')


def sample_categorical(logits, seed, step, env_offset=0, want_logp=False):
    require_cuda(logits)
    N, A = logits.shape
    actions = torch.empty(N, dtype=torch.int32, device=logits.device)
    logp = torch.empty(N, dtype=torch.float32, device=logits.device) if want_logp else None
    check(_lib.load().rl_sample_categorical(ptr(logits), N, A, int(seed), int(step), int(env_offset), ptr(actions),
                                            ptr(logp), stream()), 'sample_categorical')
    return (actions, logp) if want_logp else actions


def sample_gaussian(mean, logstd, seed, step, env_offset=0):
    require_cuda(mean, logstd)
    N, D = mean.shape
    action = torch.empty_like(mean)
    logp = torch.empty(N, dtype=torch.float32, device=mean.device)
    check(_lib.load().rl_sample_gaussian(ptr(mean), ptr(logstd), N, D, int(seed), int(step), int(env_offset),
                                         ptr(action), ptr(logp), stream()), 'sample_gaussian')
    return action, logp
