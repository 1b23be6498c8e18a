"""CPU tests of the host-side mirror of the reference API (no kernels run): Model/Agent contract,
remote_class / connect façade, futures, schedulers, stats — behaviours listed in SURVEY.md Appendix A."""
import os
import tempfile
import threading
import time

import numpy as np
import pytest
import torch
import torch.nn as nn

import parl_b200 as parl
from parl_b200.remote import exceptions as rex


class TinyModel(parl.Model):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(4, 8)
        self.fc2 = nn.Linear(8, 2)

    def forward(self, x):
        return self.fc2(torch.relu(self.fc1(x)))


def test_model_get_set_weights_and_sync():
    a, b = TinyModel(), TinyModel()
    w = a.get_weights()
    assert list(w.keys()) == list(a.state_dict().keys()) and all(isinstance(v, np.ndarray) for v in w.values())
    b.set_weights(w)
    x = torch.randn(3, 4)
    assert torch.equal(a(x), b(x))
    with pytest.raises(AttributeError):                    # like the reference: no .keys() on a list (model.py:131)
        b.set_weights(list(w.values()))
    bad = dict(w)
    bad['fc1.weight'] = np.zeros((3, 3), np.float32)
    with pytest.raises(RuntimeError):
        b.set_weights(bad)
    # sync_weights_to: target = decay*target + (1-decay)*self
    c = TinyModel()
    before = {k: v.clone() for k, v in c.named_parameters()}
    a.sync_weights_to(c, decay=0.75)
    for k, v in c.named_parameters():
        np.testing.assert_allclose(v.detach().numpy(),
                                   0.75 * before[k].detach().numpy() + 0.25 * dict(a.named_parameters())[k].detach().numpy(),
                                   rtol=1e-6)
    with pytest.raises(AssertionError):
        a.sync_weights_to(a)
    with pytest.raises(AssertionError):
        a.sync_weights_to(c, decay=1.5)


def test_agent_save_restore_train_eval():
    class Alg(parl.Algorithm):
        pass

    class Ag(parl.Agent):
        pass
    m = TinyModel()
    agent = Ag(Alg(m))
    d = tempfile.mkdtemp()
    path = os.path.join(d, 'sub', 'dir', 'model.ckpt')
    agent.save(path)
    assert os.path.exists(path)
    x = torch.randn(2, 4)
    want = m(x).detach().clone()
    with torch.no_grad():
        for p in m.parameters():
            p.add_(1.0)
    agent.restore(path)
    assert torch.equal(m(x), want)
    agent.eval()
    assert agent.training is False and m.training is False
    agent.train()
    assert agent.training is True and m.training is True


def test_algorithm_ctor_contracts():
    # type asserts happen before any CUDA requirement (SURVEY.md §8b signatures)
    from parl_b200.algorithms import IMPALA, PPO, DQN, PolicyGradient
    m = TinyModel()
    with pytest.raises(AssertionError):
        IMPALA(m, sample_batch_steps=50.0, gamma=0.99, vf_loss_coeff=0.5, clip_rho_threshold=1.0,
               clip_pg_rho_threshold=1.0)
    with pytest.raises(AssertionError):
        DQN(m, gamma=1, lr=1e-3)

    class PV(parl.Model):
        def policy(self, x):
            return x

        def value(self, x):
            return x
    with pytest.raises(AssertionError):
        PPO(PV(), clip_param=1)
    with pytest.raises(AssertionError):
        PolicyGradient(PV(), lr=0.1)            # needs an overridden forward
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):       # no silent CPU fallback
            DQN(m, gamma=0.99, lr=1e-3)


def test_remote_class_decorator_contract():
    parl.remote.disconnect()
    with pytest.raises(AssertionError):
        @parl.remote_class
        def not_a_class():
            pass
    with pytest.raises(AssertionError):
        parl.remote_class(cpu=3)

    @parl.remote_class
    class Actor(object):
        def __init__(self, k):
            self.k = k
            self.add = 'shadow'           # instance attribute shadows the method below

        def add(self, a):
            return a + self.k

        def mul(self, a):
            self.created = a * self.k
            return self.created

        def boom(self):
            raise ValueError('x')
    with pytest.raises(AssertionError):
        Actor(1)                          # before parl.connect
    parl.connect('localhost:8010')
    a = Actor(3)
    assert a.k == 3 and a.add == 'shadow'
    assert a.mul(5) == 15 and a.created == 15
    a.k = 4
    assert a.mul(5) == 20
    with pytest.raises(rex.RemoteError):
        a.boom()

    @parl.remote_class
    class Bad(object):
        def __init__(self):
            raise RuntimeError('init')
    with pytest.raises(rex.RemoteError):
        Bad()
    assert Actor(2).mul(2) == 4            # still usable afterwards


def test_remote_class_future_mode():
    parl.connect('localhost:8010')

    @parl.remote_class(wait=False)
    class Actor(object):
        def __init__(self):
            self.n = 0

        def work(self, dt, v):
            time.sleep(dt)
            self.n += 1
            return v

        def fail(self):
            raise KeyError('nope')
    actors = [Actor() for _ in range(4)]
    t0 = time.time()
    futs = [a.work(0.2, i) for i, a in enumerate(actors)]
    assert time.time() - t0 < 0.15           # calls return immediately
    assert [f.get() for f in futs] == [0, 1, 2, 3]
    assert time.time() - t0 < 0.6            # and ran concurrently
    with pytest.raises(rex.FutureGetRepeatedlyError):
        futs[0].get()
    f = actors[0].fail()
    with pytest.raises(rex.FutureFunctionError):
        f.get()
    assert actors[0].n == 1                  # attribute read waits for queued calls and returns the value
    f = actors[1].work(0.3, 'x')
    with pytest.raises(rex.FutureObjectEmpty):
        f.get_nowait()
    assert f.get(timeout=2) == 'x' and f.empty()
    for a in actors:
        a.destroy()


def test_schedulers_and_stats():
    from parl_b200.utils import PiecewiseScheduler, LinearDecayScheduler, WindowStat, TimeStat
    s = PiecewiseScheduler([(0, 0.001), (20000, 0.0005), (40000, 0.0001)])     # impala_config.py:36
    assert s.step(1000) == 0.001 and s.step(19000) == 0.0005 and s.step(30000) == 0.0001
    s2 = PiecewiseScheduler([(0, 1.0), (10, 2.0), (20, 3.0)])
    assert s2.step(25) == 2.0 and s2.step(1) == 3.0        # the reference moves one segment per call (scheduler.py:52-60)
    ld = LinearDecayScheduler(0.001, 100)
    assert abs(ld.step(50) - 0.0005) < 1e-12 and ld.step(100) == 0.0
    w = WindowStat(3)
    assert w.mean is None
    for v in (1, 2, 3, 4):
        w.add(v)
    assert w.mean == 3.0 and w.min == 2.0 and w.max == 4.0
    ts = TimeStat(5)
    with ts:
        time.sleep(0.01)
    assert 0.005 < ts.mean < 0.5


def test_calc_gae_host_matches_reference_fixture(golden):
    from parl_b200.utils import calc_gae
    g = golden('gae')
    for c in range(int(g['n_cases'])):
        p = 'c%d_' % c
        adv = calc_gae(g[p + 'rewards'], g[p + 'values'], float(g[p + 'next_value']), float(g[p + 'gamma']),
                       float(g[p + 'lam']))
        np.testing.assert_allclose(adv, g[p + 'adv'], rtol=1e-12)


def test_vector_env_auto_reset_contract():
    from parl_b200.env import VectorEnv

    class E(object):
        def __init__(self):
            self.t = 0

        def reset(self):
            self.t = 0
            return 'reset'

        def step(self, a):
            self.t += 1
            return 'obs%d' % self.t, 1.0, self.t == 2, {}
    v = VectorEnv([E(), E()])
    assert v.reset() == ['reset', 'reset']
    o, r, d, _ = v.step([0, 0])
    assert o == ['obs1', 'obs1'] and d == [False, False]
    o, r, d, _ = v.step([0, 0])
    assert o == ['reset', 'reset'] and d == [True, True] and r == [1.0, 1.0]     # vector_env.py:56-57


def test_install_as_parl():
    parl.install_as_parl()
    import parl as p2
    from parl.utils import logger, ReplayMemory          # noqa: F401
    from parl.env.vector_env import VectorEnv            # noqa: F401
    from parl.algorithms import IMPALA                   # noqa: F401
    assert p2 is parl


def test_host_slab_plan_and_actor_groups():
    from parl_b200.engine.impala import host_slab_plan
    from parl_b200.engine.impala_host import _actor_groups
    assert host_slab_plan(4096, 50) == (16, 256)                 # the C3 batch on one GPU
    assert host_slab_plan(2048, 50) == (8, 256) and host_slab_plan(1024, 50) == (4, 256)
    assert host_slab_plan(512, 50) is None and host_slab_plan(4096, 50, samples=0) is None
    assert host_slab_plan(256, 8, samples=512) == (4, 64)
    n, cols = host_slab_plan(3000, 50)                           # slabs are equal and cover every column
    assert n * cols == 3000 and cols <= 327
    assert _actor_groups(dict(env_num=4096)) == 4 and _actor_groups(dict(env_num=1024)) == 2
    assert _actor_groups(dict(env_num=512)) == 1 and _actor_groups(dict(env_num=64, actor_groups=2)) == 2
    with pytest.raises(AssertionError):
        _actor_groups(dict(env_num=10, actor_groups=3))
