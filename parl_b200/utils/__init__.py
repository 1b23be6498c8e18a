"""Host-side utilities the target examples import from ``parl.utils``
(logger, summary, schedulers, TimeStat/WindowStat, ReplayMemory, calc_gae, check_model_method)."""
from .logger import logger
from . import summary
from .scheduler import PiecewiseScheduler, LinearDecayScheduler
from .time_stat import TimeStat
from .window_stat import WindowStat
from .misc import check_model_method, get_gpu_count, is_gpu_available
from .rl_utils import calc_gae, calc_discount_sum_rewards
from .replay_memory import ReplayMemory
from .rollout_storage import RolloutStorage
from . import atari_replay_memory

__all__ = ['logger', 'summary', 'PiecewiseScheduler', 'LinearDecayScheduler', 'TimeStat', 'WindowStat',
           'check_model_method', 'get_gpu_count', 'is_gpu_available', 'calc_gae', 'calc_discount_sum_rewards',
           'ReplayMemory', 'RolloutStorage', 'atari_replay_memory']
