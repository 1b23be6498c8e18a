"""``with stat: ...`` wall-clock spans with sliding-window statistics — same surface as
parl/utils/time_stat.py:21-52 (e.g. ``with learn_time_stat: agent.learn(...)`` -> the ``learn_time_s`` metric)."""
from time import perf_counter

from .window_stat import WindowStat

__all__ = ['TimeStat']


class TimeStat(object):
    def __init__(self, window_size=1):
        self.time_samples = WindowStat(window_size)
        self._t0 = None

    def __enter__(self):
        self._t0 = perf_counter()
        return self

    def __exit__(self, exc_type, exc, tb):
        self.time_samples.add(perf_counter() - self._t0)
        return False

    def __getattr__(self, name):                           # mean / min / max of the recorded spans
        if name in ('mean', 'min', 'max'):
            return getattr(self.time_samples, name)
        raise AttributeError(name)
