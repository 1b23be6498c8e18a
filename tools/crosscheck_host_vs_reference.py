"""Build-container tool: run the host-side mirrors of parl_b200 next to the REFERENCE modules (imported from
/root/reference with the stub recipe of SURVEY.md 8c) on the same random inputs.  CPU only; prints one line per
check.  (The GPU box has no /root/reference: what must travel is distilled into tests/test_host_api.py.)
    python tools/crosscheck_host_vs_reference.py
"""
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from make_golden import _setup_reference_import  # noqa

_setup_reference_import()
sys.path.insert(0, ROOT)
import warnings  # noqa
warnings.filterwarnings('ignore')
import numpy as np  # noqa
import torch  # noqa
import torch.nn as nn  # noqa
import parl as ref  # noqa  (the reference)
import parl_b200 as ours  # noqa

ok = True


def report(name, cond, detail=''):
    global ok
    ok &= bool(cond)
    print('%-58s %s %s' % (name, 'OK ' if cond else 'MISMATCH', detail))


# ---- schedulers / window stats ---------------------------------------------------------------------------------
from parl.utils.scheduler import PiecewiseScheduler as RP, LinearDecayScheduler as RL  # noqa
from parl.utils.window_stat import WindowStat as RW  # noqa
from parl_b200.utils import PiecewiseScheduler as OP, LinearDecayScheduler as OL, WindowStat as OW  # noqa
random.seed(0)
good = True
for _ in range(200):
    bs = sorted(random.sample(range(0, 200), random.randint(1, 5)))
    sl = [(b, random.random()) for b in bs]
    a, b = OP(list(sl)), RP(list(sl))
    good &= all(a.step(k) == b.step(k) for k in [random.randint(1, 25) for _ in range(40)])
    a, b = OL(0.3, 77), RL(0.3, 77)
    good &= all(abs(a.step(k) - b.step(k)) < 1e-15 for k in [random.randint(1, 5) for _ in range(40)])
report('PiecewiseScheduler / LinearDecayScheduler', good)
good = True
for _ in range(50):
    n = random.randint(1, 7)
    a, b = OW(n), RW(n)
    for _ in range(random.randint(0, 20)):
        v = random.random()
        a.add(v), b.add(v)
        good &= abs(a.mean - b.mean) < 1e-12 and a.min == b.min and a.max == b.max and a.count == b.count
    if b.count == 0:
        good &= a.mean is None and b.mean is None
report('WindowStat', good)


# ---- Model: get_weights / set_weights / sync_weights_to -----------------------------------------------------------
def make(base):
    class Net(base):
        def __init__(self):
            super(Net, self).__init__()
            self.fc1, self.fc2 = nn.Linear(4, 8), nn.Linear(8, 3)
            self.bn = nn.BatchNorm1d(3)

        def forward(self, x):
            return self.bn(self.fc2(torch.relu(self.fc1(x))))

        def policy(self, x):
            return self.forward(x)

        def value(self, x):
            return self.forward(x).sum(-1)
    return Net


torch.manual_seed(0)
RN, ON = make(ref.Model), make(ours.Model)
rm, om = RN(), ON()
om.load_state_dict(rm.state_dict())
rw, ow = rm.get_weights(), om.get_weights()
report('Model.get_weights keys/order', list(rw.keys()) == list(ow.keys()), str(list(ow.keys())[:3]))
report('Model.get_weights values/dtypes', all(np.array_equal(rw[k], ow[k]) and rw[k].dtype == ow[k].dtype for k in rw))
rt, ot = RN(), ON()
ot.load_state_dict(rt.state_dict())
rm.sync_weights_to(rt, decay=0.3), om.sync_weights_to(ot, decay=0.3)
report('Model.sync_weights_to(decay=0.3)', all(np.allclose(a, b, atol=0, rtol=0) for a, b in
                                                zip(rt.get_weights().values(), ot.get_weights().values())))
for bad in ([1, 2], None):
    try:
        rm.set_weights(bad)
        re_ = None
    except Exception as e:
        re_ = type(e).__name__
    try:
        om.set_weights(bad)
        oe_ = None
    except Exception as e:
        oe_ = type(e).__name__
    report('Model.set_weights(%r) error type' % (bad, ), re_ == oe_, '%s vs %s' % (re_, oe_))
w2 = {k: np.asarray(v + 1) for k, v in rw.items()}
rm.set_weights(w2), om.set_weights(w2)
report('Model.set_weights round trip', all(np.array_equal(a, b) for a, b in
                                           zip(rm.get_weights().values(), om.get_weights().values())))
bad_shape = dict(w2)
k0 = list(bad_shape)[0]
bad_shape[k0] = np.zeros((2, 2), np.float32)
errs = []
for m_ in (rm, om):
    try:
        m_.set_weights(bad_shape)
        errs.append(None)
    except Exception as e:
        errs.append(type(e).__name__)
report('Model.set_weights(shape mismatch) error type', errs[0] == errs[1], str(errs))


# ---- Algorithm / Agent: save, restore, train/eval ------------------------------------------------------------------
class RAlg(ref.Algorithm):
    def __init__(self, model):
        super(RAlg, self).__init__(model)
        self.model = model

    def predict(self, obs):
        return self.model.policy(obs)

    def learn(self, *a):
        return None


class OAlg(ours.Algorithm):
    def __init__(self, model):
        super(OAlg, self).__init__(model)
        self.model = model

    def predict(self, obs):
        return self.model.policy(obs)

    def learn(self, *a):
        return None


class RAgent(ref.Agent):
    def predict(self, obs):
        return self.alg.predict(obs)

    def learn(self, *a):
        return None


class OAgent(ours.Agent):
    def predict(self, obs):
        return self.alg.predict(obs)

    def learn(self, *a):
        return None


try:
    ra, oa = RAgent(RAlg(rm)), OAgent(OAlg(om))
    d = tempfile.mkdtemp()
    p1, p2 = os.path.join(d, 'a', 'b', 'ref.ckpt'), os.path.join(d, 'c', 'd', 'ours.ckpt')
    ra.save(p1), oa.save(p2)
    report('Agent.save creates missing directories', os.path.exists(p1) and os.path.exists(p2))
    sd1, sd2 = torch.load(p1), torch.load(p2)
    report('Agent.save file = model state_dict (same keys)', list(sd1.keys()) == list(sd2.keys()))
    oa2 = OAgent(OAlg(ON()))
    oa2.restore(p1)                                   # our agent restores a checkpoint written by the reference
    report('Agent.restore(reference checkpoint)', all(np.array_equal(a, b) for a, b in
                                                      zip(oa2.alg.model.get_weights().values(), rm.get_weights().values())))
    ra.eval(), oa.eval()
    report('Agent.eval()/train() flags', ra.training == oa.training is False and oa.alg.model.training is False)
    ra.train(), oa.train()
    report('Agent.train() flags', ra.training == oa.training is True and oa.alg.model.training is True)
    rwa, owa = ra.get_weights(), oa.get_weights()
    report('Agent.get_weights structure', type(rwa) == type(owa) and list(rwa.keys()) == list(owa.keys()))
except Exception as e:   # the torch backend's Agent may want a device: report instead of dying
    report('Agent checks', False, repr(e))

# ---- calc_gae / misc -----------------------------------------------------------------------------------------------
from parl.utils import calc_gae as rgae  # noqa
from parl_b200.utils import calc_gae as ogae  # noqa
rng = np.random.RandomState(0)
good = True
for n in (1, 2, 17, 200):
    r, v = rng.randn(n), rng.randn(n)
    good &= np.allclose(rgae(r, v, 0.3, 0.99, 0.95), ogae(r, v, 0.3, 0.99, 0.95), rtol=1e-12, atol=1e-12)
report('calc_gae', good)

print('ALL OK' if ok else 'SOME MISMATCHES')
