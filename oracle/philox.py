"""Philox4x32-10 counter-based RNG + the exact-arithmetic helpers of the RNG contract.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The device kernels
(parl_b200/csrc/philox.cuh) implement the same functions; because every
operation below is either integer arithmetic or a single correctly-rounded
IEEE-754 binary32 operation (no fused multiply-add, no libm), the two sides
agree bit for bit.

Algorithm: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as
1, 2, 3" (SC'11), Philox-4x32 with 10 rounds (Random123 v1.14 constants).
The reference itself uses unseeded ``np.random`` (parl/tests/gym.py:117-207,
examples/IMPALA/atari_agent.py:39-40), so the contract is ours; SURVEY.md §7
"Hard parts" requires a shared counter-based RNG for bit-exact actions/dones.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

# counter word 3 = stream id
STREAM_FRAME = 0      # Atari-synth frame pixels        ctr = (env, step, block, 0)
STREAM_REWDONE = 1    # reward / done draws             ctr = (env, step, 0, 1)
STREAM_ACTION = 2     # action sampling uniforms        ctr = (env, step, 0, 2)
STREAM_OBS = 3        # MuJoCo-synth / CartPole obs     ctr = (env, step, block, 3)
STREAM_GAUSS = 4      # diag-Gaussian action noise      ctr = (env, step, block, 4)
STREAM_REPLAY = 5     # replay / PER sampling uniforms  ctr = (slot, draw, 0, 5)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-R. All inputs broadcastable uint32 arrays/ints.

    Returns four uint32 arrays.
    """
    c0, c1, c2, c3 = np.broadcast_arrays(
        *[np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3)])
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1,
                          hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def split_seed(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def u01_24(x):
    """uint32 -> float32 uniform in [0,1) with 24 random bits (exact)."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def prob_threshold(p):
    """Bernoulli(p) as an integer compare: event <=> x < threshold (uint32)."""
    return min(int(p * 4294967296.0), 0xFFFFFFFF)


# ---------------------------------------------------------------------------
# exact exp: every step is one IEEE binary32 op -> bit-identical on CPU and GPU
# ---------------------------------------------------------------------------
_LOG2E = np.float32(1.4426950408889634)
_LN2 = 0.6931471805599453
_EXP2_COEF = [np.float32(_LN2 ** k / float(np.prod(np.arange(1, k + 1)) if k else 1)) for k in range(8)]


def exp_exact(x):
    """exp(x) for x <= 0 in binary32 using only mul/add/rint (no FMA, no libm).

    t = x*log2e; n = rint(t); f = t - n in [-.5,.5]; 2^f by a degree-7 Horner
    polynomial (separate multiply and add, each rounded); result = p * 2^n.
    Relative error <= ~4e-6 (dominated by rounding of x*log2e at |x|~87); inputs
    below -87 flush to 0.
    """
    x = np.asarray(x, dtype=np.float32)
    t = x * _LOG2E
    t = np.maximum(t, np.float32(-126.0))
    n = np.rint(t).astype(np.float32)
    f = (t - n).astype(np.float32)
    p = np.full_like(f, _EXP2_COEF[7])
    for k in range(6, -1, -1):
        p = (p * f).astype(np.float32)
        p = (p + _EXP2_COEF[k]).astype(np.float32)
    scale = ((n.astype(np.int32) + 127) << 23).astype(np.int32).view(np.float32)
    out = (p * scale).astype(np.float32)
    return np.where(x < np.float32(-87.0), np.float32(0.0), out).astype(np.float32)


def sample_categorical_exact(logits, u):
    """Inverse-CDF categorical sample with exact arithmetic.

    logits [N, A] float32, u [N] float32 in [0,1).  w_j = exp_exact(l_j - max),
    c_j = sequential float32 prefix sum, thr = u * c_{A-1},
    action = #{j : c_j <= thr} clamped to A-1.
    Replaces the reference's per-row ``np.random.choice(len(prob), 1, p=prob)``
    (examples/IMPALA/atari_agent.py:39-40) / ``Categorical.sample()``
    (parl/algorithms/torch/a2c.py:75, ppo.py:173).
    """
    logits = np.asarray(logits, dtype=np.float32)
    m = logits.max(axis=-1, keepdims=True)
    w = exp_exact((logits - m).astype(np.float32))
    c = np.zeros_like(w)
    acc = np.zeros(w.shape[:-1], dtype=np.float32)
    for j in range(w.shape[-1]):
        acc = (acc + w[..., j]).astype(np.float32)
        c[..., j] = acc
    thr = (np.asarray(u, dtype=np.float32) * c[..., -1]).astype(np.float32)
    a = (c <= thr[..., None]).sum(axis=-1)
    return np.minimum(a, logits.shape[-1] - 1).astype(np.int32)


def action_uniforms(seed, env_ids, step):
    k0, k1 = split_seed(seed)
    x0, _, _, _ = philox4x32(env_ids, step, 0, STREAM_ACTION, k0, k1)
    return u01_24(x0)
