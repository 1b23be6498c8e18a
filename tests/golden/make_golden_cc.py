"""Continuous-control fixtures recorded from the REFERENCE (build container only):

    python tests/golden/make_golden_cc.py     ->  tests/golden/cc.npz

For DDPG / TD3 / SAC (parl/algorithms/torch/{ddpg,td3,sac}.py) on the small MLPs of cc_models.py: the initial weights,
three batches, the losses ``learn`` returns and the weights of model and target model after three ``learn`` calls.
TD3 runs with policy_noise = 0 (its target noise is a torch.randn_like draw); SAC's reparameterisation noise is
injected: Normal.rsample is patched to ``loc + scale * eps`` with recorded eps (critic draw, then actor draw, per
step).  Also one "table" case per algorithm: the critic TD target / loss / gradient on given Q arrays."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _setup_reference_import  # noqa: E402


def main():
    _setup_reference_import()
    import warnings
    warnings.filterwarnings('ignore')
    import numpy as np
    import torch
    import parl
    from parl.algorithms import DDPG, TD3, SAC
    from torch.distributions import Normal
    from cc_models import make_models

    ACModel = make_models(parl.Model)
    rng = np.random.RandomState(4321)
    OBS, ACT, N, STEPS = 11, 3, 64, 3
    out = dict(obs_dim=OBS, act_dim=ACT, n=N, steps=STEPS)
    eps_queue = []

    def patched_rsample(self, sample_shape=torch.Size()):
        return self.loc + self.scale * eps_queue.pop(0)
    orig_rsample = Normal.rsample
    Normal.rsample = patched_rsample
    try:
        for kind, cls, kw in (('ddpg', DDPG, dict(gamma=0.99, tau=0.005, actor_lr=3e-4, critic_lr=1e-3)),
                              ('td3', TD3, dict(gamma=0.99, tau=0.005, actor_lr=3e-4, critic_lr=3e-4, policy_noise=0.0,
                                                noise_clip=0.5, policy_freq=2)),
                              ('sac', SAC, dict(gamma=0.99, tau=0.005, alpha=0.2, actor_lr=3e-4, critic_lr=3e-4))):
            torch.manual_seed(7)
            model = ACModel(OBS, ACT, kind)
            alg = cls(model, **kw)
            alg.model.cpu(), alg.target_model.cpu()
            p = kind + '_'
            for k, v in alg.model.state_dict().items():
                out[p + 'w0_' + k] = v.detach().numpy().copy()
            for s in range(STEPS):
                obs = rng.randn(N, OBS).astype(np.float32)
                act = np.clip(rng.randn(N, ACT), -1, 1).astype(np.float32)
                rew = rng.randn(N, 1).astype(np.float32)
                nobs = rng.randn(N, OBS).astype(np.float32)
                term = (rng.rand(N, 1) < 0.2).astype(np.float32)
                q = p + 's%d_' % s
                out.update({q + 'obs': obs, q + 'action': act, q + 'reward': rew, q + 'next_obs': nobs, q + 'terminal': term})
                if kind == 'sac':
                    e1, e2 = rng.randn(N, ACT).astype(np.float32), rng.randn(N, ACT).astype(np.float32)
                    out.update({q + 'eps_critic': e1, q + 'eps_actor': e2})
                    eps_queue[:] = [torch.tensor(e1), torch.tensor(e2)]
                r = alg.learn(torch.tensor(obs), torch.tensor(act), torch.tensor(rew), torch.tensor(nobs), torch.tensor(term))
                if r is not None:
                    out[q + 'critic_loss'] = np.float32(r[0].item())
                    out[q + 'actor_loss'] = np.float32(r[1].item())
            for k, v in alg.model.state_dict().items():
                out[p + 'w1_' + k] = v.detach().numpy().copy()
            for k, v in alg.target_model.state_dict().items():
                out[p + 't1_' + k] = v.detach().numpy().copy()
    finally:
        Normal.rsample = orig_rsample

    # ---- table case: the critic TD arithmetic alone (what rl_twin_q_td_loss_fwd_bwd replaces) ------------------
    import torch.nn.functional as F
    M = 257
    q1, q2 = rng.randn(M, 1).astype(np.float32), rng.randn(M, 1).astype(np.float32)
    tq1, tq2 = rng.randn(M, 1).astype(np.float32), rng.randn(M, 1).astype(np.float32)
    logp = rng.randn(M, 1).astype(np.float32)
    rew = rng.randn(M, 1).astype(np.float32)
    term = (rng.rand(M, 1) < 0.3).astype(np.float32)
    out.update(dict(tab_q1=q1, tab_q2=q2, tab_tq1=tq1, tab_tq2=tq2, tab_logp=logp, tab_reward=rew, tab_terminal=term))
    T = torch.tensor
    for name, gamma, alpha in (('ddpg', 0.99, 0.0), ('td3', 0.98, 0.0), ('sac', 0.97, 0.2)):
        a, b = T(q1, requires_grad=True), T(q2, requires_grad=True)
        if name == 'ddpg':      # ddpg.py:63-73
            target = T(rew) + ((1. - T(term)) * gamma * T(tq1)).detach()
            loss = F.mse_loss(a, target)
        elif name == 'td3':     # td3.py:86-94
            target = T(rew) + (1 - T(term)) * gamma * torch.min(T(tq1), T(tq2))
            loss = F.mse_loss(a, target) + F.mse_loss(b, target)
        else:                   # sac.py:92-99
            target = torch.min(T(tq1), T(tq2)) - alpha * T(logp)
            target = T(rew) + gamma * (1. - T(term)) * target
            loss = F.mse_loss(a, target) + F.mse_loss(b, target)
        loss.backward()
        out.update({'tab_%s_target' % name: target.numpy(), 'tab_%s_loss' % name: np.float32(loss.item()),
                    'tab_%s_d_q1' % name: a.grad.numpy(), 'tab_%s_gamma' % name: np.float32(gamma),
                    'tab_%s_alpha' % name: np.float32(alpha)})
        if name != 'ddpg':
            out['tab_%s_d_q2' % name] = b.grad.numpy()
    np.savez(os.path.join(HERE, 'cc.npz'), **out)
    print('wrote cc.npz with', len(out), 'arrays')


if __name__ == '__main__':
    main()
