"""Sliding-window statistics of scalar samples — same surface as parl/utils/window_stat.py:20-54
(``add``, ``mean``, ``min``, ``max``, ``count``; the three statistics are ``None`` before the first sample)."""
from collections import deque

__all__ = ['WindowStat']


def _stat(reduce_fn):
    def getter(self):
        return float(reduce_fn(self._window)) if self._window else None
    return property(getter)


class WindowStat(object):
    def __init__(self, window_size):
        self._window = deque(maxlen=int(window_size))     # the deque drops the oldest sample by itself
        self.count = 0                                     # samples ever added

    def add(self, obj):
        self._window.append(float(obj))
        self.count += 1

    mean = _stat(lambda w: sum(w) / len(w))
    min = _stat(min)
    max = _stat(max)
