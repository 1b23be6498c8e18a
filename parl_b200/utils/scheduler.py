"""Schedulers used by the IMPALA / A2C learners (parl/utils/scheduler.py):
PiecewiseScheduler([(step, value), ...]).step(n) and LinearDecayScheduler(start, max_steps).step(n)."""

__all__ = ['PiecewiseScheduler', 'LinearDecayScheduler']


class PiecewiseScheduler(object):
    def __init__(self, scheduler_list):
        assert len(scheduler_list) > 0
        for i in range(len(scheduler_list) - 1):
            assert scheduler_list[i][0] < scheduler_list[i + 1][0]
        self.scheduler_list = scheduler_list
        self.cur_index = 0
        self.cur_step = 0
        self.cur_value = scheduler_list[0][1]
        self.scheduler_num = len(scheduler_list)

    def step(self, step_num=1):
        assert isinstance(step_num, int) and step_num >= 1
        self.cur_step += step_num
        while self.cur_index < self.scheduler_num - 1 and self.cur_step >= self.scheduler_list[self.cur_index + 1][0]:
            self.cur_index += 1
            self.cur_value = self.scheduler_list[self.cur_index][1]
        return self.cur_value


class LinearDecayScheduler(object):
    def __init__(self, start_value, max_steps):
        assert max_steps > 0
        self.cur_step = 0
        self.max_steps = max_steps
        self.start_value = start_value

    def step(self, step_num=1):
        assert isinstance(step_num, int) and step_num >= 1
        self.cur_step = min(self.cur_step + step_num, self.max_steps)
        return self.start_value * (1.0 - (self.cur_step * 1.0 / self.max_steps))
