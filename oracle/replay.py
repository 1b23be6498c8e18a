"""Replay memories — CPU restatement (TEST INFRASTRUCTURE).

SumTree / ProportionalPER follow benchmark/fluid/Prioritized_DQN/proportional_per.py:18-157
with the RNG draws made explicit (``u`` arguments) so the device sampler can be
checked bit-exactly on indices; AtariReplay follows
benchmark/torch/dqn/replay_memory.py:22-113 (uint8 frame ring, context stacking with
episode-boundary zeroing).  Pinned by tests/golden/make_golden.py (reference run).
"""
import numpy as np


class SumTree(object):
    def __init__(self, capacity):
        self.capacity = capacity
        self.tree = np.zeros(2 * capacity - 1, np.float64)
        self._ptr = 0
        self._min = 10.0                                   # proportional_per.py:24
        self.filled = 0

    def add(self, priority):
        tree_idx = self._ptr + self.capacity - 1
        self.update(tree_idx, priority)
        self._ptr = (self._ptr + 1) % self.capacity
        self.filled = min(self.filled + 1, self.capacity)
        return tree_idx

    def update(self, tree_idx, priority):                  # :36-42
        diff = priority - self.tree[tree_idx]
        self.tree[tree_idx] = priority
        while tree_idx != 0:
            tree_idx = (tree_idx - 1) >> 1
            self.tree[tree_idx] += diff
        self._min = min(self._min, priority)

    def retrieve(self, value):                             # :44-60
        parent = 0
        n = len(self.tree)
        while True:
            left = 2 * parent + 1
            if left >= n:
                leaf = parent
                break
            if value <= self.tree[left]:
                parent = left
            else:
                value -= self.tree[left]
                parent = left + 1
        return leaf, self.tree[leaf]

    @property
    def total_p(self):
        return self.tree[0]


class ProportionalPER(object):
    def __init__(self, alpha, seg_num, size, eps=0.01):
        self.alpha, self.seg_num, self.size, self.eps = alpha, seg_num, int(size), eps
        self.elements = SumTree(self.size)
        self._max_priority = 1.0

    def store(self, delta=None):                           # :105-111
        if not delta:
            delta = self._max_priority
        ps = np.power(delta + self.eps, self.alpha)
        return self.elements.add(ps)

    def update(self, indices, priorities):                 # :113-118
        priorities = np.array(priorities) + self.eps
        pa = np.power(priorities, self.alpha)
        for idx, p in zip(indices, pa):
            self.elements.update(idx, p)
            self._max_priority = max(p, self._max_priority)

    def sample(self, u, beta=1.0):
        """:126-157 with np.random.uniform(low, high) replaced by low + u*(high-low)."""
        total = self.elements.total_p
        seg = total / self.seg_num
        idxs, prios = [], []
        for i in range(self.seg_num):
            low, high = seg * i, seg * (i + 1)
            val = low + float(u[i]) * (high - low)
            leaf, p = self.elements.retrieve(val)
            idxs.append(leaf)
            prios.append(p)
        probs = self.size * np.array(prios) / total
        min_prob = self.size * self.elements._min / total
        w = np.power(probs / min_prob, -beta)
        return np.array(idxs), w


class AtariReplay(object):
    """uint8 single-frame ring + (context_len+1)-frame sampling
    (benchmark/torch/dqn/replay_memory.py:22-113)."""

    def __init__(self, max_size, obs_shape, context_len):
        self.max_size, self.obs_shape, self.context_len = int(max_size), obs_shape, int(context_len)
        self.obs = np.zeros((self.max_size, ) + obs_shape, 'uint8')
        self.action = np.zeros((self.max_size, ), 'int32')
        self.reward = np.zeros((self.max_size, ), 'float32')
        self.isOver = np.zeros((self.max_size, ), 'bool')
        self._curr_size = 0
        self._curr_pos = 0

    def append(self, obs, action, reward, isOver):
        p = self._curr_pos
        self.obs[p], self.action[p], self.reward[p], self.isOver[p] = obs, action, reward, isOver
        self._curr_size = min(self._curr_size + 1, self.max_size)
        self._curr_pos = (p + 1) % self.max_size

    def sample(self, idx):                                 # :59-85
        obs = np.zeros((self.context_len + 1, ) + self.obs_shape, np.uint8)
        obs_idx = np.arange(idx, idx + self.context_len + 1) % self._curr_size
        has_last = False
        for k in range(self.context_len - 2, -1, -1):
            if self.isOver[obs_idx[k]]:
                has_last = True
                obs_idx = obs_idx[k + 1:]
                obs[k + 1:] = self.obs[obs_idx]
                break
        if not has_last:
            obs = self.obs[obs_idx]
        real = (idx + self.context_len - 1) % self._curr_size
        return obs, self.reward[real], self.action[real], self.isOver[real]

    def batch_indices(self, raw):
        """:103-107: raw = randint(curr_size - context_len - 1, size=batch)."""
        return (self._curr_pos + np.asarray(raw)) % self._curr_size

    def sample_batch_by_raw(self, raw):
        exps = [self.sample(i) for i in self.batch_indices(raw)]
        return (np.asarray([e[0] for e in exps], 'uint8'), np.asarray([e[2] for e in exps], 'int32'),
                np.asarray([e[1] for e in exps], 'float32'), np.asarray([e[3] for e in exps], 'bool'))
