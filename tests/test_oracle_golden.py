"""The oracle (oracle/*.py) against the fixtures produced by the reference itself
(tests/golden/make_golden.py) and the reference's V-trace known-answer test."""
import numpy as np
import pytest

from oracle import vtrace as ovt
from oracle import returns as oret
from oracle import losses as olo
from oracle import replay as orp
from oracle import philox as oph


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert [int(v) for v in oph.philox4x32(0, 0, 0, 0, 0, 0)] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert [int(v) for v in oph.philox4x32(f, f, f, f, f, f)] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(v) for v in oph.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_exp_exact_accuracy():
    x = np.linspace(-86, 0, 20001).astype(np.float32)
    rel = np.abs(oph.exp_exact(x) - np.exp(x.astype(np.float64))) / np.exp(x.astype(np.float64))
    assert rel.max() < 5e-6
    assert oph.exp_exact(np.float32(0.0)) == np.float32(1.0)


def test_sample_categorical_exact_distribution():
    rng = np.random.RandomState(0)
    logits = np.tile(np.array([[0.0, 1.0, 2.0, -1.0]], np.float32), (200000, 1))
    u = rng.rand(200000).astype(np.float32)
    a = oph.sample_categorical_exact(logits, u)
    p = np.exp(logits[0]) / np.exp(logits[0]).sum()
    freq = np.bincount(a, minlength=4) / len(a)
    assert np.abs(freq - p).max() < 5e-3


@pytest.mark.parametrize('B', [1, 4])
def test_vtrace_kat(golden, B):
    g = golden('vtrace_kat')
    k = {n: g['B%d_%s' % (B, n)] for n in ('blp', 'tlp', 'discounts', 'rewards', 'values', 'bootstrap_value')}
    vs, pg = ovt.from_importance_weights(k['blp'], k['tlp'], k['discounts'], k['rewards'], k['values'],
                                         k['bootstrap_value'], 3.7, 2.2)
    np.testing.assert_almost_equal(g['B%d_vs' % B], vs, 5)          # same tolerance as vtrace_test_paddle.py:140
    np.testing.assert_almost_equal(g['B%d_pg_advantages' % B], pg, 5)


def test_vtrace_kat_survey_values():
    """SURVEY.md §8(c) lists the B=1 vectors explicitly."""
    k = ovt.kat_inputs(1)
    vs, pg = ovt.from_importance_weights(k['blp'], k['tlp'], k['discounts'], k['rewards'], k['values'],
                                         k['bootstrap_value'], 3.7, 2.2)
    np.testing.assert_almost_equal(vs[:, 0], [0.2001819, 2.7096822, 8.513628, 11.932398, 7.3300004], 5)
    np.testing.assert_almost_equal(pg[:, 0], [0.2001819, 1.7096823, 6.5136285, 10.876616, 1.9800003], 5)


def test_a2c_loss_golden(golden):
    g = golden('a2c')
    for c in range(3):
        p = 'c%d_' % c
        o = olo.a2c_loss(g[p + 'logits'], g[p + 'values'], g[p + 'actions'], g[p + 'advantages'],
                         g[p + 'target_values'], 0.5, -0.01)
        for k in ('total_loss', 'pi_loss', 'vf_loss', 'entropy'):
            np.testing.assert_allclose(o[k], g[p + k], rtol=1e-5)
        np.testing.assert_allclose(o['d_logits'], g[p + 'd_logits'], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(o['d_values'], g[p + 'd_values'], rtol=1e-5, atol=1e-6)


def test_ppo_loss_golden(golden):
    g = golden('ppo')
    for c in range(int(g['n_cases'])):
        p = 'c%d_' % c
        kw = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01,
                  use_clipped_value_loss=bool(g[p + 'clipv']), norm_adv=bool(g[p + 'norm_adv']))
        if bool(g[p + 'continuous']):
            kw.update(mean=g[p + 'mean'], logstd=g[p + 'logstd'])
        else:
            kw.update(logits=g[p + 'logits'])
        o = olo.ppo_loss(g[p + 'values'], g[p + 'batch_action'], g[p + 'batch_value'], g[p + 'batch_return'],
                         g[p + 'batch_logprob'], g[p + 'batch_adv'], **kw)
        for k in ('value_loss', 'action_loss', 'entropy_loss'):
            np.testing.assert_allclose(o[k], g[p + k], rtol=1e-5, atol=1e-7)
        for k in ('d_values', 'd_logits', 'd_mean', 'd_logstd'):
            if p + k in g.files:
                np.testing.assert_allclose(o[k], g[p + k], rtol=1e-5, atol=1e-7)


def test_td_loss_golden(golden):
    g = golden('dqn')
    for c in range(int(g['n_cases'])):
        p = 'c%d_' % c
        o = olo.td_loss(g[p + 'q'], g[p + 'q_next_target'], g[p + 'action'], g[p + 'reward'], g[p + 'terminal'], 0.99)
        np.testing.assert_allclose(o['loss'], g[p + 'dqn_loss'], rtol=1e-5)
        np.testing.assert_allclose(o['d_q'], g[p + 'dqn_d_q'], rtol=1e-5, atol=1e-8)
        o = olo.td_loss(g[p + 'q'], g[p + 'q_next_target'], g[p + 'action'], g[p + 'reward'], g[p + 'terminal'], 0.99,
                        q_online_next=g[p + 'q_next_online'])
        np.testing.assert_allclose(o['loss'], g[p + 'ddqn_loss'], rtol=1e-5)
        np.testing.assert_allclose(o['d_q'], g[p + 'ddqn_d_q'], rtol=1e-5, atol=1e-8)


def test_pg_loss_golden(golden):
    g = golden('pg')
    o = olo.pg_loss(g['prob'], g['action'], g['reward'])
    np.testing.assert_allclose(o['loss'], g['loss'], rtol=1e-5)
    np.testing.assert_allclose(o['d_prob'], g['d_prob'], rtol=1e-5, atol=1e-8)


def test_calc_gae_golden(golden):
    g = golden('gae')
    for c in range(int(g['n_cases'])):
        p = 'c%d_' % c
        adv = oret.calc_gae(g[p + 'rewards'], g[p + 'values'], float(g[p + 'next_value']), float(g[p + 'gamma']),
                            float(g[p + 'lam']))
        np.testing.assert_allclose(adv, g[p + 'adv'], rtol=1e-12)
    # the value quoted in SURVEY.md §8(c)
    np.testing.assert_allclose(oret.calc_gae([1, 1, 1], [.5, .4, .3], .2, .99, .95), [2.53394564, 1.741569, 0.898],
                               rtol=1e-7)


def test_a2c_segment_gae_equals_masked_recurrence():
    rng = np.random.RandomState(3)
    T, B = 20, 7
    r, v = rng.rand(T, B), rng.randn(T, B)
    d = rng.rand(T, B) < 0.15
    boot = rng.randn(B)
    adv, tv = oret.a2c_segment_gae_time_major(r, v, d, boot, 0.99, 0.95)
    # masked single-pass recurrence the device kernel uses
    acc = np.zeros(B)
    ref = np.zeros((T, B))
    for t in range(T - 1, -1, -1):
        nv = boot if t == T - 1 else v[t + 1]
        nt = 1.0 - d[t]
        delta = r[t] + 0.99 * nv * nt - v[t]
        acc = delta + 0.99 * 0.95 * nt * acc
        ref[t] = acc
    np.testing.assert_allclose(adv, ref, rtol=1e-10, atol=1e-12)


def test_compute_returns_golden(golden):
    g = golden('ppo_returns')
    for c in range(int(g['n_cases'])):
        p = 'c%d_' % c
        adv, ret = oret.compute_returns(g[p + 'rewards'], g[p + 'values'], g[p + 'dones'], g[p + 'value'], g[p + 'done'])
        np.testing.assert_array_equal(adv, g[p + 'adv'])
        np.testing.assert_array_equal(ret, g[p + 'ret'])


def test_per_golden(golden):
    g = golden('per')
    per = orp.ProportionalPER(alpha=float(g['alpha']), seg_num=int(g['seg_num']), size=int(g['capacity']),
                              eps=float(g['eps']))
    for d in g['store_delta']:
        per.store(None if d < 0 else float(d))
    np.testing.assert_allclose(per.elements.tree, g['tree_after_store'], rtol=1e-13)
    for rnd in range(g['u'].shape[0]):
        idx, w = per.sample(g['u'][rnd], beta=0.5 + 0.1 * rnd)
        np.testing.assert_array_equal(idx, g['indices'][rnd])
        np.testing.assert_allclose(w, g['weights'][rnd], rtol=1e-12)
        per.update(idx, g['new_priorities'][rnd])
    np.testing.assert_allclose(per.elements.tree, g['tree_final'], rtol=1e-12)
    assert per.elements._min == float(g['min_final'])
    assert per._max_priority == float(g['max_priority_final'])


def test_atari_replay_golden(golden):
    g = golden('atari_replay')
    rpm = orp.AtariReplay(int(g['size']), g['frames'].shape[1:], int(g['ctx']))
    for f, a, r, o in zip(g['frames'], g['actions'], g['rewards'], g['overs']):
        rpm.append(f, a, r, o)
    obs, act, rew, over = rpm.sample_batch_by_raw(g['raw'])
    np.testing.assert_array_equal(obs, g['obs'])
    np.testing.assert_array_equal(act, g['action'])
    np.testing.assert_array_equal(rew, g['reward'])
    np.testing.assert_array_equal(over, g['isOver'])


def test_twin_q_td_golden(golden):
    """oracle.losses.twin_q_td against the critic TD arithmetic recorded from the reference's DDPG / TD3 / SAC
    expressions (tests/golden/make_golden_cc.py)."""
    g = golden('cc')
    for name in ('ddpg', 'td3', 'sac'):
        kw = {}
        if name != 'ddpg':
            kw.update(q2=g['tab_q2'], q2_target_next=g['tab_tq2'])
        if name == 'sac':
            kw.update(next_log_prob=g['tab_logp'], alpha=float(g['tab_sac_alpha']))
        o = olo.twin_q_td(g['tab_q1'], g['tab_tq1'], g['tab_reward'], g['tab_terminal'], float(g['tab_%s_gamma' % name]), **kw)
        np.testing.assert_allclose(o['target'], g['tab_%s_target' % name].reshape(-1), rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(o['loss'], g['tab_%s_loss' % name], rtol=1e-5)
        np.testing.assert_allclose(o['d_q1'], g['tab_%s_d_q1' % name].reshape(-1), rtol=1e-5, atol=1e-8)
        if name != 'ddpg':
            np.testing.assert_allclose(o['d_q2'], g['tab_%s_d_q2' % name].reshape(-1), rtol=1e-5, atol=1e-8)
