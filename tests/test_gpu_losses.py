"""GPU parity: A2C / PPO / DQN / DDQN / PER / PG loss kernels and the GAE scans through the C ABI,
against the fixtures recorded from the reference itself (tests/golden/*.npz) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import losses as olo
from oracle import returns as oret

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def cu(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV).contiguous()


def test_a2c_loss_golden(golden):
    from parl_b200 import kernels as K
    g = golden('a2c')
    for c in range(3):
        p = 'c%d_' % c
        r = K.a2c_loss_fwd_bwd(cu(g[p + 'logits']), cu(g[p + 'values']), cu(g[p + 'actions']), cu(g[p + 'advantages']),
                               cu(g[p + 'target_values']), 0.5, -0.01)
        L = r['losses'].cpu().numpy()
        for i, k in enumerate(('total_loss', 'pi_loss', 'vf_loss', 'entropy')):
            np.testing.assert_allclose(L[i], g[p + k], rtol=1e-4, err_msg=k)
        np.testing.assert_allclose(r['d_logits'].cpu().numpy(), g[p + 'd_logits'], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(r['d_values'].cpu().numpy(), g[p + 'd_values'], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize('N,A', [(5120, 2), (5000, 18), (129, 7), (1, 3)])
def test_a2c_loss_vs_oracle(N, A):
    from parl_b200 import kernels as K
    rng = np.random.RandomState(N + A)
    lg = (2 * rng.randn(N, A)).astype(np.float32)
    v, adv, tv = [rng.randn(N).astype(np.float32) for _ in range(3)]
    a = rng.randint(0, A, N).astype(np.int32)
    o = olo.a2c_loss(lg, v, a, adv, tv, 0.5, -0.01)
    r = K.a2c_loss_fwd_bwd(cu(lg), cu(v), cu(a), cu(adv), cu(tv), 0.5, -0.01)
    L = r['losses'].cpu().numpy()
    for i, k in enumerate(('total_loss', 'pi_loss', 'vf_loss', 'entropy')):
        np.testing.assert_allclose(L[i], o[k], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(r['d_logits'].cpu().numpy(), o['d_logits'], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(r['d_values'].cpu().numpy(), o['d_values'], rtol=1e-4, atol=1e-6)


def test_ppo_loss_golden(golden):
    from parl_b200 import kernels as K
    g = golden('ppo')
    for c in range(int(g['n_cases'])):
        p = 'c%d_' % c
        kw = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01,
                  use_clipped_value_loss=bool(g[p + 'clipv']), norm_adv=bool(g[p + 'norm_adv']))
        if bool(g[p + 'continuous']):
            kw.update(mean=cu(g[p + 'mean']), logstd=cu(g[p + 'logstd']))
            act = cu(g[p + 'batch_action'], torch.float32)
        else:
            kw.update(logits=cu(g[p + 'logits']))
            act = cu(g[p + 'batch_action'])
        r = K.ppo_loss_fwd_bwd(cu(g[p + 'values']), act, cu(g[p + 'batch_value']), cu(g[p + 'batch_return']),
                               cu(g[p + 'batch_logprob']), cu(g[p + 'batch_adv']), **kw)
        L = r['losses'].cpu().numpy()
        for i, k in enumerate(('value_loss', 'action_loss', 'entropy_loss')):
            np.testing.assert_allclose(L[i], g[p + k], rtol=1e-4, atol=1e-6, err_msg='%s case %d' % (k, c))
        for k in ('d_values', 'd_logits', 'd_mean', 'd_logstd'):
            if p + k in g.files:
                np.testing.assert_allclose(r[k].cpu().numpy(), g[p + k], rtol=1e-4, atol=1e-7, err_msg='%s case %d' % (k, c))


def test_ppo_loss_c4_minibatch_vs_oracle():
    """C4 minibatch (M=131072, obs 17 / act 6 Gaussian)."""
    from parl_b200 import kernels as K
    rng = np.random.RandomState(4)
    M, D = 131072, 6
    mean = rng.randn(M, D).astype(np.float32)
    logstd = (0.2 * rng.randn(D)).astype(np.float32)
    act = (mean + np.exp(logstd) * rng.randn(M, D)).astype(np.float32)
    v = rng.randn(M).astype(np.float32)
    oldv = (v + 0.2 * rng.randn(M)).astype(np.float32)
    ret, adv = rng.randn(M).astype(np.float32), rng.randn(M).astype(np.float32)
    with torch.no_grad():
        lp = torch.distributions.Normal(torch.tensor(mean), torch.tensor(np.exp(logstd))).log_prob(torch.tensor(act)).sum(1)
    oldlp = (lp.numpy() + 0.2 * rng.randn(M)).astype(np.float32)
    o = olo.ppo_loss(v, act, oldv, ret, oldlp, adv, mean=mean, logstd=logstd, clip_param=0.2, entropy_coef=0.0)
    r = K.ppo_loss_fwd_bwd(cu(v), cu(act), cu(oldv), cu(ret), cu(oldlp), cu(adv), mean=cu(mean), logstd=cu(logstd),
                           clip_param=0.2, entropy_coef=0.0)
    L = r['losses'].cpu().numpy()
    for i, k in enumerate(('value_loss', 'action_loss', 'entropy_loss', 'loss')):
        np.testing.assert_allclose(L[i], o[k], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(r['d_mean'].cpu().numpy(), o['d_mean'], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(r['d_logstd'].cpu().numpy(), o['d_logstd'], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(r['d_values'].cpu().numpy(), o['d_values'], rtol=1e-4, atol=1e-9)


def test_td_loss_golden(golden):
    from parl_b200 import kernels as K
    g = golden('dqn')
    for c in range(int(g['n_cases'])):
        p = 'c%d_' % c
        args = (cu(g[p + 'q']), cu(g[p + 'q_next_target']), cu(g[p + 'action']).reshape(-1),
                cu(g[p + 'reward']).reshape(-1), cu(g[p + 'terminal']).reshape(-1), 0.99)
        r = K.td_loss_fwd_bwd(*args)
        np.testing.assert_allclose(r['losses'].item(), g[p + 'dqn_loss'], rtol=1e-5)
        np.testing.assert_allclose(r['d_q'].cpu().numpy(), g[p + 'dqn_d_q'], rtol=1e-5, atol=1e-8)
        r = K.td_loss_fwd_bwd(*args, q_online_next=cu(g[p + 'q_next_online']))
        np.testing.assert_allclose(r['losses'].item(), g[p + 'ddqn_loss'], rtol=1e-5)
        np.testing.assert_allclose(r['d_q'].cpu().numpy(), g[p + 'ddqn_d_q'], rtol=1e-5, atol=1e-8)
        w = np.random.RandomState(c).rand(g[p + 'q'].shape[0]).astype(np.float32)
        o = olo.td_loss(g[p + 'q'], g[p + 'q_next_target'], g[p + 'action'], g[p + 'reward'], g[p + 'terminal'], 0.99,
                        weights=w)
        r = K.td_loss_fwd_bwd(*args, weights=cu(w), want_td_abs=True)
        np.testing.assert_allclose(r['losses'].item(), o['loss'], rtol=1e-5)
        np.testing.assert_allclose(r['d_q'].cpu().numpy(), o['d_q'], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(r['td_abs'].cpu().numpy(), o['td_abs'], rtol=1e-5, atol=1e-6)


def test_pg_loss_golden(golden):
    from parl_b200 import kernels as K
    g = golden('pg')
    r = K.pg_loss_fwd_bwd(cu(g['prob']), cu(g['action']), cu(g['reward']))
    np.testing.assert_allclose(r['losses'].item(), g['loss'], rtol=1e-5)
    np.testing.assert_allclose(r['d_prob'].cpu().numpy(), g['d_prob'], rtol=1e-4, atol=1e-7)


def test_gae_scan_ppo_bit_exact_golden(golden):
    from parl_b200 import kernels as K
    g = golden('ppo_returns')
    for c in range(int(g['n_cases'])):
        p = 'c%d_' % c
        adv, ret = K.gae_scan(cu(g[p + 'rewards']), cu(g[p + 'values']), cu(g[p + 'dones']), cu(g[p + 'value']),
                              cu(g[p + 'done']), 0.99, 0.95)
        np.testing.assert_array_equal(adv.cpu().numpy(), g[p + 'adv'])     # bit-exact vs RolloutStorage
        np.testing.assert_array_equal(ret.cpu().numpy(), g[p + 'ret'])


def test_gae_scan_ppo_c4_size():
    from parl_b200 import kernels as K
    rng = np.random.RandomState(9)
    T, B = 2048, 2048
    r, v = rng.rand(T, B).astype(np.float32), rng.randn(T, B).astype(np.float32)
    d = (rng.rand(T, B) < 0.01).astype(np.float32)
    lv, ld = rng.randn(B).astype(np.float32), (rng.rand(B) < 0.01).astype(np.float32)
    adv, ret = K.gae_scan(cu(r), cu(v), cu(d), cu(lv), cu(ld))
    sub = slice(100, 164)
    oa, orr = oret.compute_returns(r[:, sub], v[:, sub], d[:, sub], lv[sub], ld[sub])
    np.testing.assert_array_equal(adv.cpu().numpy()[:, sub], oa)
    np.testing.assert_array_equal(ret.cpu().numpy()[:, sub], orr)


def test_gae_segments_vs_calc_gae(golden):
    from parl_b200 import kernels as K
    rng = np.random.RandomState(2)
    T, B = 20, 256
    r, v = rng.rand(T, B).astype(np.float32), rng.randn(T, B).astype(np.float32)
    d = rng.rand(T, B) < 0.1
    boot = rng.randn(B).astype(np.float32)
    for lam in (1.0, 0.95):
        adv, tv = K.gae_scan_segments(cu(r), cu(v), cu(d), cu(boot), 0.99, lam)
        oa, ot = oret.a2c_segment_gae_time_major(r, v, d, boot, 0.99, lam)
        np.testing.assert_allclose(adv.cpu().numpy(), oa.astype(np.float32), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(tv.cpu().numpy(), ot.astype(np.float32), rtol=1e-6, atol=1e-6)
    # single-segment fixtures recorded from parl.utils.calc_gae itself
    g = golden('gae')
    for c in range(int(g['n_cases'])):
        p = 'c%d_' % c
        L = len(g[p + 'rewards'])
        rr = g[p + 'rewards'].astype(np.float32).reshape(L, 1)
        vv = g[p + 'values'].astype(np.float32).reshape(L, 1)
        nv = np.float32(g[p + 'next_value']).reshape(1)
        adv, _ = K.gae_scan_segments(cu(rr), cu(vv), cu(np.zeros((L, 1), bool)), cu(nv), float(g[p + 'gamma']),
                                     float(g[p + 'lam']))
        want = oret.calc_gae(rr[:, 0].astype(np.float64), vv[:, 0].astype(np.float64), float(nv[0]), float(g[p + 'gamma']),
                             float(g[p + 'lam']))
        np.testing.assert_allclose(adv.cpu().numpy()[:, 0], want, rtol=1e-6, atol=1e-6)
