"""Size-independent properties of the oracle restatements (CPU): they pin the oracle from a second side next to the
golden fixtures, and they are the same properties the GPU parity tests use at full size."""
import numpy as np
import pytest

from oracle import vtrace as ovt
from oracle import returns as oret
from oracle import philox as oph


def _vt(x, **kw):
    return ovt.from_importance_weights(x['blp'], x['tlp'], x['discounts'], x['rewards'], x['values'],
                                       x['bootstrap_value'], **kw)


def _rollout(T, B, seed):
    rng = np.random.RandomState(seed)
    return dict(blp=rng.randn(T, B).astype(np.float32) * 0.3 - 1.0, tlp=rng.randn(T, B).astype(np.float32) * 0.3 - 1.0,
                discounts=(0.99 * (rng.rand(T, B) > 0.1)).astype(np.float32), rewards=rng.randn(T, B).astype(np.float32),
                values=rng.randn(T, B).astype(np.float32), bootstrap_value=rng.randn(B).astype(np.float32))


@pytest.mark.parametrize('T,B', [(1, 1), (5, 3), (50, 16)])
def test_vtrace_recursive_form_equals_definition(T, B):
    """vtrace.py:118-122 (backward recursion) == the O(T^2) definition of the reference's own test."""
    x = _rollout(T, B, T * 100 + B)
    vs, pg = _vt(x, clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2)
    gvs, gpg = ovt.ground_truth_o_t2(x['blp'].astype(np.float64), x['tlp'].astype(np.float64),
                                     x['discounts'].astype(np.float64), x['rewards'].astype(np.float64),
                                     x['values'].astype(np.float64), x['bootstrap_value'].astype(np.float64), 3.7, 2.2)
    np.testing.assert_allclose(vs, gvs, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(pg, gpg, rtol=2e-5, atol=2e-5)


def test_vtrace_on_policy_is_n_step_return():
    """target == behaviour (rho = c = 1): vs_t is the discounted n-step bootstrapped return, independent of V_t."""
    x = _rollout(12, 4, 7)
    x['tlp'] = x['blp'].copy()
    vs, pg = _vt(x)
    ret = x['bootstrap_value'].astype(np.float64)
    for t in range(11, -1, -1):
        ret = x['rewards'][t] + x['discounts'][t].astype(np.float64) * ret
        np.testing.assert_allclose(vs[t], ret, rtol=1e-4, atol=1e-4)
    nxt = np.concatenate([vs[1:], x['bootstrap_value'][None]], 0)
    np.testing.assert_allclose(pg, x['rewards'] + x['discounts'] * nxt - x['values'], rtol=1e-5, atol=1e-5)


def test_vtrace_episode_cut_blocks_the_scan():
    """discount 0 at step t (done): nothing after t can influence vs at or before t."""
    x = _rollout(10, 2, 3)
    x['discounts'][4] = 0.0
    vs_a, _ = _vt(x)
    y = {k: v.copy() for k, v in x.items()}
    y['rewards'][5:] += 100.0
    y['values'][5:] -= 50.0
    y['bootstrap_value'] += 7.0
    vs_b, _ = _vt(y)
    assert np.array_equal(vs_a[:5], vs_b[:5])


def test_env_major_to_time_major_is_split_batches():
    T, B = 5, 3
    flat = np.arange(B * T)                      # flat index b*T+t (examples/IMPALA/actor.py:79-89)
    tm = ovt.env_major_to_time_major(flat, T)
    assert tm.shape == (T, B)
    for t in range(T):
        for b in range(B):
            assert tm[t, b] == b * T + t


def test_gae_lambda_one_is_discounted_return_minus_value():
    """calc_gae with lam = 1: A_t = sum_k gamma^k r_{t+k} + gamma^{n-t} V_next - V_t (rl_utils.py:21-51)."""
    rng = np.random.RandomState(1)
    r, v, nv, g = rng.randn(20), rng.randn(20), 0.37, 0.97
    adv = oret.calc_gae(r, v, nv, g, 1.0)
    ret, want = nv, np.zeros(20)
    for t in range(19, -1, -1):
        ret = r[t] + g * ret
        want[t] = ret - v[t]
    np.testing.assert_allclose(adv, want, rtol=1e-9, atol=1e-9)
    # lam = 0: one-step TD errors
    td = oret.calc_gae(r, v, nv, g, 0.0)
    np.testing.assert_allclose(td, r + g * np.append(v[1:], nv) - v, rtol=1e-12, atol=1e-12)


def test_segment_gae_equals_per_episode_calc_gae():
    """The time-major segmented scan == calc_gae run per episode segment (benchmark/torch/a2c/actor.py:82-102)."""
    rng = np.random.RandomState(5)
    T, B, g, lam = 17, 3, 0.99, 0.95
    r, v = rng.randn(T, B), rng.randn(T, B)
    d = rng.rand(T, B) < 0.2
    boot = rng.randn(B)
    adv, tgt = oret.a2c_segment_gae_time_major(r, v, d, boot, g, lam)
    for b in range(B):
        start = 0
        for t in range(T):
            if d[t, b] or t == T - 1:
                nv = 0.0 if d[t, b] else boot[b]
                seg = oret.calc_gae(r[start:t + 1, b], v[start:t + 1, b], nv, g, lam)
                np.testing.assert_allclose(adv[start:t + 1, b], seg, rtol=1e-9, atol=1e-9)
                np.testing.assert_allclose(tgt[start:t + 1, b], seg + v[start:t + 1, b], rtol=1e-9, atol=1e-9)
                start = t + 1


def test_philox_streams_are_disjoint_and_counter_based():
    """Same (key, counter) -> same block; any counter word change -> different block (the RNG contract of DESIGN.md 2)."""
    a = oph.philox4x32(5, 9, 0, 0, 123, 456)
    assert [int(v) for v in a] == [int(v) for v in oph.philox4x32(5, 9, 0, 0, 123, 456)]
    for ctr in ((6, 9, 0, 0), (5, 10, 0, 0), (5, 9, 1, 0), (5, 9, 0, 1)):
        assert [int(v) for v in oph.philox4x32(*ctr, 123, 456)] != [int(v) for v in a]
    assert [int(v) for v in oph.philox4x32(5, 9, 0, 0, 124, 456)] != [int(v) for v in a]


def test_sample_categorical_exact_edge_cases():
    """Inverse-CDF sampling (action = #{j : cumsum_j <= u * total}): u ~ 0 -> first action with non-negligible mass,
    u -> 1 -> last action, near-one-hot logits -> that action."""
    logits = np.array([[0.0, 0.0, 0.0, 0.0], [-80.0, 0.0, -80.0, -80.0], [3.0, -2.0, 0.5, 1.0]], np.float32)
    lo = oph.sample_categorical_exact(logits, np.full(3, np.float32(1e-6)))
    hi = oph.sample_categorical_exact(logits, np.full(3, np.float32(1.0 - 2 ** -24)))
    assert lo[0] == 0 and hi[0] == 3
    assert lo[1] == 1 and hi[1] == 1
    assert 0 <= lo[2] <= hi[2] <= 3


def test_twin_q_td_properties():
    """Continuous-control critic TD oracle: the hand-written gradients are the derivative of the loss; terminal
    transitions bootstrap nothing; the twin target is the smaller of the two; the entropy term shifts the target by
    -gamma (1 - terminal) alpha log pi."""
    from oracle import losses as olo
    rng = np.random.RandomState(5)
    N = 64
    q1, q2, tq1, tq2, lp, rew = [rng.randn(N).astype(np.float32) for _ in range(6)]
    term = (rng.rand(N) < 0.3).astype(np.float32)
    o = olo.twin_q_td(q1, tq1, rew, term, 0.9, q2=q2, q2_target_next=tq2, next_log_prob=lp, alpha=0.25)
    # finite differences of loss = mse1 + mse2 in float64
    def loss(a, b):
        t = o['target'].astype(np.float64)
        return np.mean((a - t) ** 2) + np.mean((b - t) ** 2)
    eps = 1e-3
    for i in (0, 7, 63):
        d = np.zeros(N)
        d[i] = eps
        g1 = (loss(q1 + d, q2) - loss(q1 - d, q2)) / (2 * eps)
        g2 = (loss(q1, q2 + d) - loss(q1, q2 - d)) / (2 * eps)
        np.testing.assert_allclose(o['d_q1'][i], g1, rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(o['d_q2'][i], g2, rtol=1e-4, atol=1e-7)
    np.testing.assert_array_equal(o['target'][term == 1], rew[term == 1])
    plain = olo.twin_q_td(q1, tq1, rew, term, 0.9, q2=q2, q2_target_next=tq2)
    live = term == 0
    np.testing.assert_allclose(plain['target'][live], rew[live] + 0.9 * np.minimum(tq1, tq2)[live], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose((plain['target'] - o['target'])[live], 0.9 * 0.25 * lp[live], rtol=1e-4, atol=1e-6)
    single = olo.twin_q_td(q1, tq1, rew, term, 0.9)
    assert 'd_q2' not in single and abs(single['loss'] - single['mse1']) == 0.0
