"""GPU test of the IMPALA example's host contract on the device actor pool: a ``@parl.remote_class`` Actor with
sample() / set_weights() / get_metrics() and ``AtariAgent.learn(numpy...)`` driven exactly like the Learner loop of
examples/IMPALA/train.py:165-194 (keys, dtypes and env-major order of examples/IMPALA/actor.py:79-91)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device('cuda:0')


@pytest.mark.parametrize('groups', [1, 2])
def test_remote_actor_and_agent_learn_through_numpy_contract(groups):
    """groups = 2: the pool is split into two column groups whose rollouts overlap the download of the group before;
    the env-major sample dict (and the env side, keyed by the global env index) is the same."""
    import parl_b200 as parl
    from parl_b200.engine.impala_host import DeviceImpalaActor, AtariAgent
    from oracle import envs as oenv
    torch.manual_seed(0)
    B, T, A, seed = 64, 8, 18, 77
    cfg = dict(env_num=B, sample_batch_steps=T, act_dim=A, seed=seed, actor_groups=groups)
    parl.connect('localhost:8010')
    agent = AtariAgent(cfg, device=DEV)
    Actor = parl.remote_class(wait=False)(DeviceImpalaActor)
    actor = Actor(cfg, device=DEV)
    actor.set_weights(agent.get_weights()).get()
    batch = actor.sample().get()
    assert set(batch) == {'obs', 'actions', 'behaviour_logits', 'rewards', 'dones'}
    assert batch['obs'].shape == (B * T, 4, 84, 84) and batch['obs'].dtype == np.uint8
    assert batch['actions'].dtype == np.int64 and batch['behaviour_logits'].shape == (B * T, A)
    assert batch['rewards'].dtype == np.float32 and batch['dones'].dtype == np.bool_
    # env-major order + the env twin: row b*T + t is env b at step t (actor.py:79-89)
    ref = oenv.AtariSynthVec(B, seed, p_done=0.1)
    ref.reset()
    obs_em = batch['obs'].reshape(B, T, 4, 84 * 84)
    rew_em, done_em = batch['rewards'].reshape(B, T), batch['dones'].reshape(B, T)
    for t in range(T):
        assert np.array_equal(obs_em[:, t], ref.obs())
        _, r, d = ref.step()
        assert np.array_equal(rew_em[:, t], r) and np.array_equal(done_em[:, t], d)
    # the actor really runs the learner's weights: its behaviour logits == the agent model's policy on the same obs
    with torch.no_grad():
        lg = agent.alg.model.policy(torch.from_numpy(batch['obs'][:256]).to(DEV)).float().cpu().numpy()
    assert np.abs(lg - batch['behaviour_logits'][:256]).max() < 0.05 * max(1.0, np.abs(lg).max())
    # Learner loop, future mode: the next sample is produced while the agent learns
    w0 = {k: v.copy() for k, v in agent.get_weights().items()}
    fut = actor.sample()
    for _ in range(3):
        out = agent.learn(batch['obs'], batch['actions'], batch['behaviour_logits'], batch['rewards'], batch['dones'],
                          0.001, -0.01)
        assert len(out) == 5 and all(np.isfinite(x) for x in out)
        batch = fut.get()
        actor.set_weights(agent.get_weights())
        fut = actor.sample()
    fut.get()
    w1 = agent.get_weights()
    assert any(not np.array_equal(w0[k], w1[k]) for k in w0)
    m = actor.get_metrics().get()
    assert set(m) == {'episode_rewards', 'episode_steps'} and len(m['episode_rewards']) == len(m['episode_steps']) > 0
    actor.destroy()


def test_slab_pipelined_host_learner_matches_one_shot():
    """AtariAgent.learn with the observation upload pipelined slab by slab against the compute (the default at the
    C3 batch) gives the one-shot path's losses and parameter update up to fp32 summation order."""
    from parl_b200.engine.impala_host import DeviceImpalaActor, AtariAgent
    B, T, A = 256, 8, 18
    cfg = dict(env_num=B, sample_batch_steps=T, act_dim=A, seed=5)
    actor = DeviceImpalaActor(cfg, device=DEV)
    torch.manual_seed(1)
    ag0 = AtariAgent(cfg, device=DEV)
    torch.manual_seed(1)
    ag1 = AtariAgent(cfg, device=DEV)
    ag0.engine.host_slab_samples = 0               # one-shot
    ag1.engine.host_slab_samples = 64 * T          # 4 slabs of 64 env columns
    assert ag0.engine._host_slab_plan() is None and ag1.engine._host_slab_plan() == (4, 64)
    actor.set_weights(ag0.get_weights())
    seen = {}
    ag0.engine.alg.grad_sync = lambda g: seen.__setitem__(0, g.clone())      # called with the flat gradient before the step
    ag1.engine.alg.grad_sync = lambda g: seen.__setitem__(1, g.clone())
    for it in range(2):
        batch = actor.sample()
        args = (batch['obs'], batch['actions'], batch['behaviour_logits'], batch['rewards'], batch['dones'], 0.001, -0.01)
        l0, l1 = ag0.learn(*args), ag1.learn(*args)
        # the two paths run the same kernels on different batch sizes (the fc GEMM switches to split-K at 512 samples),
        # so bf16 activations differ in the last bit here and there: round-off sized differences, nothing structural
        np.testing.assert_allclose(l1[:4], l0[:4], rtol=5e-3, atol=1e-2)
        assert abs(l1[4] - l0[4]) < 1e-5
        g0, g1 = seen[0], seen[1]                  # the gradients the two optimizer steps consumed
        rel = ((g0 - g1).norm() / g0.norm()).item()
        assert rel < 2e-2, (it, rel)               # a dropped / doubled / misplaced slab would be O(1)
        w0, w1 = ag0.get_weights(), ag1.get_weights()
        for k in w0:
            # Adam's first steps move every weight by ~lr = 1e-3 (update ~ lr * sign(g)); the two paths agree far
            # inside that except for components whose gradient is round-off sized
            d = np.abs(w0[k] - w1[k])
            assert d.mean() < 1e-4 and (d > 5e-4).mean() < 2e-2, (it, k, d.mean(), d.max(), (d > 5e-4).mean())
        actor.set_weights(ag0.get_weights())
        ag1.set_weights(ag0.get_weights())         # same starting point for the next round
