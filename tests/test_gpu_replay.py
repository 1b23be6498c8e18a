"""GPU parity: device sum-tree / PER and replay gathers vs fixtures recorded from the reference."""
import numpy as np
import pytest
import torch

from oracle import replay as orp

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def test_per_trace_golden(golden):
    from parl_b200 import kernels as K
    g = golden('per')
    cap, seg = int(g['capacity']), int(g['seg_num'])
    alpha, eps = float(g['alpha']), float(g['eps'])
    tree = K.DeviceSumTree(cap, DEV)
    sd = g['store_delta'].astype(np.float32)
    # first 40 stores use delta=None (-> max_priority), the rest explicit deltas; feed in two batches
    tree.store(0, 40, alpha, eps)
    tree.store(40, cap - 40, alpha, eps, delta=torch.as_tensor(sd[40:]).to(DEV))
    np.testing.assert_allclose(tree.tree.cpu().numpy(), g['tree_after_store'], rtol=1e-6)
    for rnd in range(g['u'].shape[0]):
        u = torch.as_tensor(g['u'][rnd].astype(np.float32)).to(DEV)
        tidx, eidx, w = tree.sample(seg, 0.5 + 0.1 * rnd, cap, u=u)
        # float32 uniforms vs the fixture's float64 ones: indices must still agree (no draw sits on a boundary)
        assert np.array_equal(tidx.cpu().numpy(), g['indices'][rnd])
        assert np.array_equal(eidx.cpu().numpy(), g['indices'][rnd] - cap + 1)
        np.testing.assert_allclose(w.cpu().numpy(), g['weights'][rnd], rtol=1e-5)
        tree.update(tidx, torch.as_tensor(g['new_priorities'][rnd].astype(np.float32)).to(DEV), alpha, eps)
    np.testing.assert_allclose(tree.tree.cpu().numpy(), g['tree_final'], rtol=1e-6)
    st = tree.state.cpu().numpy()
    np.testing.assert_allclose(st[0], float(g['min_final']), rtol=1e-6)
    np.testing.assert_allclose(st[1], float(g['max_priority_final']), rtol=1e-6)


def test_per_large_random_vs_oracle():
    from parl_b200 import kernels as K
    cap, seg = 100000, 256            # non power of two: leaves on two levels
    rng = np.random.RandomState(0)
    ref = orp.ProportionalPER(alpha=0.6, seg_num=seg, size=cap)
    tree = K.DeviceSumTree(cap, DEV)
    delta = (rng.rand(cap) * 2 + 0.01).astype(np.float32)
    for d in delta:
        ref.store(float(d))
    for lo in range(0, cap, 4096):
        n = min(4096, cap - lo)
        tree.store(lo, n, 0.6, 0.01, delta=torch.as_tensor(delta[lo:lo + n]).to(DEV))
    np.testing.assert_allclose(tree.tree[0].item(), ref.elements.total_p, rtol=1e-10)
    for rnd in range(3):
        u = rng.rand(seg).astype(np.float32)
        tidx, eidx, w = tree.sample(seg, 0.7, cap, u=torch.as_tensor(u).to(DEV))
        ridx, rw = ref.sample(u.astype(np.float64), beta=0.7)
        assert np.array_equal(tidx.cpu().numpy(), ridx)
        np.testing.assert_allclose(w.cpu().numpy(), rw, rtol=1e-5)
        newp = rng.rand(seg).astype(np.float32)
        tree.update(tidx, torch.as_tensor(newp).to(DEV), 0.6, 0.01)
        ref.update(ridx, newp.astype(np.float64))
        np.testing.assert_allclose(tree.tree[0].item(), ref.elements.total_p, rtol=1e-9)
    np.testing.assert_allclose(tree.tree.cpu().numpy(), ref.elements.tree, rtol=1e-7, atol=1e-9)
    # Philox-driven draws stay inside their strata
    tidx, eidx, w = tree.sample(seg, 1.0, cap, seed=5, draw=3)
    assert eidx.min().item() >= 0 and eidx.max().item() < cap


def test_atari_replay_gather_golden(golden):
    from parl_b200 import kernels as K
    g = golden('atari_replay')
    size, ctx = int(g['size']), int(g['ctx'])
    ref = orp.AtariReplay(size, g['frames'].shape[1:], ctx)
    for f, a, r, o in zip(g['frames'], g['actions'], g['rewards'], g['overs']):
        ref.append(f, a, r, o)
    # pad 6-byte frames to one 16-byte block so the device gather's alignment contract holds
    HW = 16
    frames = np.zeros((size, HW), np.uint8)
    frames[:, :6] = ref.obs.reshape(size, 6)
    idx = ref.batch_indices(g['raw']).astype(np.int32)
    out = K.replay_gather_frames(torch.as_tensor(frames).to(DEV), torch.as_tensor(ref.isOver).to(DEV),
                                 torch.as_tensor(idx).to(DEV), ref._curr_size, ctx)
    got = out.cpu().numpy()[:, :, :6].reshape(len(idx), ctx + 1, 3, 2)
    np.testing.assert_array_equal(got, g['obs'])
    real = (idx + ctx - 1) % ref._curr_size
    a = K.gather_rows(torch.as_tensor(ref.action.reshape(-1, 1)).to(DEV), torch.as_tensor(real.astype(np.int32)).to(DEV))
    np.testing.assert_array_equal(a.cpu().numpy().reshape(-1), g['action'])


def test_gather_rows_uniform_replay():
    from parl_b200 import kernels as K
    rng = np.random.RandomState(1)
    obs = rng.randn(1000, 17).astype(np.float32)
    idx = rng.randint(0, 1000, 256).astype(np.int32)
    out = K.gather_rows(torch.as_tensor(obs).to(DEV), torch.as_tensor(idx).to(DEV))
    np.testing.assert_array_equal(out.cpu().numpy(), obs[idx])      # parl/utils/replay_memory.py:61-66
