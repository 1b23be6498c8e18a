"""Host logic of the one-launch operand refresh (parl_b200/engine/packing.py) on CPU: the index permutation built by
evaluating the nets' pack expressions on index tensors must reproduce those expressions applied to the parameters
themselves — out[i] = flat[idx[i]] (0 where idx < 0) is what rl_gather_cast computes on the device."""
import torch

from parl_b200.engine.actor_net import AtariActorNet
from parl_b200.engine.nets import AtariActorCritic
from parl_b200.engine.packing import PackedOperands
from parl_b200.engine.train_net import AtariTrainNet


class _Stub(object):
    A = 18
    window_form = True


SPECS = dict(
    train=(AtariTrainNet._sources, [('w1', (32, 256)), ('w2', (64, 512)), ('w3', (64, 576)), ('wfc', (512, 5184)),
                                    ('wpi', (18, 512)), ('wv', (1, 512)), ('wfcT', (5184, 512)), ('whT', (512, 32)),
                                    ('w3T', (64, 576)), ('w2T', (128, 256))]),
    actor=(AtariActorNet._sources, [('w1', (32, 256)), ('w2', (64, 512)), ('w3', (64, 576)), ('wfc', (512, 5184)),
                                    ('wpi', (18, 512)), ('wv', (1, 512))]))
BIASES = [('b1', (32, )), ('b2', (64, )), ('b3', (64, )), ('bfc', (512, )), ('bpi', (18, )), ('bv', (1, ))]


def _flatten(model):
    """Re-home the parameters into one flat float32 buffer the way FlatAdam does (16-byte aligned views)."""
    params = list(model.parameters())
    offs, total = [], 0
    for p in params:
        offs.append(total)
        total += (p.numel() + 3) // 4 * 4
    flat = torch.zeros(total)
    with torch.no_grad():
        for p, o in zip(params, offs):
            flat[o:o + p.numel()].copy_(p.reshape(-1))
            p.data = flat[o:o + p.numel()].view(p.shape)
    return flat


def test_index_permutation_reproduces_the_pack_expressions():
    torch.manual_seed(0)
    model = AtariActorCritic(18)
    with torch.no_grad():
        for p in model.parameters():
            p.normal_(0, 1)
    flat = _flatten(model)
    for kind, (src_fn, weights) in SPECS.items():
        ops = PackedOperands('cpu')
        for name, shape in weights:
            ops.declare(name, shape, torch.bfloat16)
        for name, shape in BIASES:
            ops.declare(name, shape, torch.float32)
        holder = _Stub()
        ops.materialize(holder)
        sources = lambda P, full, fn=src_fn: fn(holder, P, full)
        assert ops.bind_flat(flat, model, sources)
        params = dict(model.named_parameters())
        exprs = dict(sources(lambda n: params[n], lambda shape: torch.zeros(shape)))
        for dtype, idx in ops.idx.items():
            gathered = torch.where(idx >= 0, flat[idx.clamp(min=0).long()], torch.zeros(()))
            for name, (dt, off, n, shape) in ops.slots.items():
                if dt != dtype:
                    continue
                got = gathered[off:off + n].view(shape)
                assert torch.equal(got, exprs[name].float()), (kind, name)
                assert idx[off:off + n].max().item() < flat.numel()
        # padding between carved operands is never read from the flat buffer
        covered = torch.zeros(ops.arenas[torch.bfloat16].numel(), dtype=torch.bool)
        for name, (dt, off, n, shape) in ops.slots.items():
            if dt == torch.bfloat16:
                covered[off:off + n] = True
        assert (ops.idx[torch.bfloat16][~covered] == -1).all()


def test_bind_refuses_parameters_outside_the_flat_buffer():
    model = AtariActorCritic(18)
    flat = torch.zeros(16)
    ops = PackedOperands('cpu')
    ops.declare('w1', (32, 256), torch.bfloat16)
    ops.materialize(_Stub())
    assert not ops.bind_flat(flat, model, lambda P, full: [('w1', P('conv1.weight').reshape(32, 256))])
    assert ops.flat is None
