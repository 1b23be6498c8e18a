"""Step-driven hyper-parameter schedules with the surface of parl/utils/scheduler.py (used by the IMPALA / A2C
learners, examples/IMPALA/train.py:60-66): ``PiecewiseScheduler([(boundary_step, value), ...]).step(n)`` and
``LinearDecayScheduler(start_value, max_steps).step(n)`` advance an internal step counter by ``n`` and return
the value in force."""
__all__ = ['PiecewiseScheduler', 'LinearDecayScheduler']


def _advance(counter, step_num):
    assert isinstance(step_num, int) and step_num >= 1, 'step_num must be a positive int'
    return counter + step_num


class PiecewiseScheduler(object):
    """Piecewise-constant schedule.  Like the reference (scheduler.py:52-60) a call moves on by AT MOST ONE
    segment, so a single jump across two boundaries reaches the later value one call late — kept, because
    drop-in means the learning-rate sequence of an unmodified training script does not change."""

    def __init__(self, scheduler_list):
        assert len(scheduler_list) > 0
        self._boundaries = [b for b, _ in scheduler_list]
        self._values = [v for _, v in scheduler_list]
        assert all(a < b for a, b in zip(self._boundaries, self._boundaries[1:])), 'boundaries must increase'
        self.cur_step = 0
        self._segment = 0

    def step(self, step_num=1):
        self.cur_step = _advance(self.cur_step, step_num)
        nxt = self._segment + 1
        if nxt < len(self._boundaries) and self.cur_step >= self._boundaries[nxt]:
            self._segment = nxt
        return self._values[self._segment]


class LinearDecayScheduler(object):
    """start_value * (1 - step / max_steps), held at 0 once max_steps is reached."""

    def __init__(self, start_value, max_steps):
        assert max_steps > 0
        self.start_value, self.max_steps, self.cur_step = start_value, max_steps, 0

    def step(self, step_num=1):
        self.cur_step = min(_advance(self.cur_step, step_num), self.max_steps)
        return self.start_value * (1.0 - self.cur_step / float(self.max_steps))
