"""GPU parity of the continuous-control family: rl_twin_q_td_loss_fwd_bwd against the oracle and
against the critic TD arithmetic recorded from the reference's DDPG / TD3 / SAC expressions, and the algorithm classes
``parl_b200.algorithms.{DDPG,TD3,SAC}`` against three ``learn`` calls of the reference's classes on the same small MLPs,
batches and noise (tests/golden/make_golden_cc.py -> cc.npz)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import losses as olo

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden'))

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(x):
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(DEV).contiguous()


@pytest.mark.parametrize('name', ['ddpg', 'td3', 'sac'])
def test_twin_q_td_golden(golden, name):
    from parl_b200 import kernels as K
    g = golden('cc')
    kw = {}
    if name != 'ddpg':
        kw.update(q2=cu(g['tab_q2']).view(-1), q2_target_next=cu(g['tab_tq2']).view(-1))
    if name == 'sac':
        kw.update(next_log_prob=cu(g['tab_logp']).view(-1), alpha=float(g['tab_sac_alpha']))
    r = K.twin_q_td_loss_fwd_bwd(cu(g['tab_q1']).view(-1), cu(g['tab_tq1']).view(-1), cu(g['tab_reward']).view(-1),
                                 cu(g['tab_terminal']).view(-1), float(g['tab_%s_gamma' % name]), want_target=True, **kw)
    np.testing.assert_allclose(r['target'].cpu().numpy(), g['tab_%s_target' % name].reshape(-1), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(r['losses'][0].item(), g['tab_%s_loss' % name], rtol=1e-5)
    np.testing.assert_allclose(r['d_q1'].cpu().numpy(), g['tab_%s_d_q1' % name].reshape(-1), rtol=1e-5, atol=1e-8)
    if name != 'ddpg':
        np.testing.assert_allclose(r['d_q2'].cpu().numpy(), g['tab_%s_d_q2' % name].reshape(-1), rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize('N', [1, 127, 4096, 100003])
def test_twin_q_td_vs_oracle(N):
    from parl_b200 import kernels as K
    rng = np.random.RandomState(N)
    q1, q2, tq1, tq2, lp, rew = [rng.randn(N).astype(np.float32) for _ in range(6)]
    term = (rng.rand(N) < 0.25).astype(np.float32)
    o = olo.twin_q_td(q1, tq1, rew, term, 0.99, q2=q2, q2_target_next=tq2, next_log_prob=lp, alpha=0.2)
    r = K.twin_q_td_loss_fwd_bwd(cu(q1), cu(tq1), cu(rew), cu(term), 0.99, q2=cu(q2), q2_target_next=cu(tq2),
                                 next_log_prob=cu(lp), alpha=0.2, want_target=True)
    assert np.array_equal(r['target'].cpu().numpy(), o['target'])            # same float32 operation order: bit-exact
    L = r['losses'].cpu().numpy()
    np.testing.assert_allclose(L, [o['loss'], o['mse1'], o['mse2']], rtol=1e-5)
    np.testing.assert_allclose(r['d_q1'].cpu().numpy(), o['d_q1'], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(r['d_q2'].cpu().numpy(), o['d_q2'], rtol=1e-6, atol=1e-12)
    # determinism of the grid reduction
    r2 = K.twin_q_td_loss_fwd_bwd(cu(q1), cu(tq1), cu(rew), cu(term), 0.99, q2=cu(q2), q2_target_next=cu(tq2),
                                  next_log_prob=cu(lp), alpha=0.2)
    assert torch.equal(r['losses'], r2['losses'])


@pytest.mark.parametrize('kind', ['ddpg', 'td3', 'sac'])
def test_algorithm_matches_reference_after_three_learn_calls(golden, kind):
    import parl_b200 as parl
    from parl_b200.algorithms import DDPG, TD3, SAC
    from cc_models import make_models
    g = golden('cc')
    ACModel = make_models(parl.Model)
    obs_dim, act_dim, steps = int(g['obs_dim']), int(g['act_dim']), int(g['steps'])
    model = ACModel(obs_dim, act_dim, kind)
    p = kind + '_'
    model.load_state_dict({k: torch.from_numpy(g[p + 'w0_' + k]) for k in model.state_dict()})
    if kind == 'ddpg':
        alg = DDPG(model, gamma=0.99, tau=0.005, actor_lr=3e-4, critic_lr=1e-3)
    elif kind == 'td3':
        alg = TD3(model, gamma=0.99, tau=0.005, actor_lr=3e-4, critic_lr=3e-4, policy_noise=0.0, noise_clip=0.5,
                  policy_freq=2)
    else:
        alg = SAC(model, gamma=0.99, tau=0.005, alpha=0.2, actor_lr=3e-4, critic_lr=3e-4)
    for s in range(steps):
        q = p + 's%d_' % s
        if kind == 'sac':
            eps = [cu(g[q + 'eps_critic']), cu(g[q + 'eps_actor'])]
            alg.noise_fn = lambda mean, _e=eps: _e.pop(0)
        r = alg.learn(g[q + 'obs'], g[q + 'action'], g[q + 'reward'], g[q + 'next_obs'], g[q + 'terminal'])
        if kind != 'td3':
            np.testing.assert_allclose(float(r[0]), g[q + 'critic_loss'], rtol=2e-4, atol=1e-6)
            np.testing.assert_allclose(float(r[1]), g[q + 'actor_loss'], rtol=2e-4, atol=1e-5)
    # three Adam steps move a weight by <= 3 lr ~ 1e-3..3e-3: agreement to 2e-5 pins optimizer, actor update, Polyak
    for k, v in alg.model.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), g[p + 'w1_' + k], rtol=1e-3, atol=2e-5, err_msg=k)
    for k, v in alg.target_model.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), g[p + 't1_' + k], rtol=1e-3, atol=2e-5, err_msg='target ' + k)
