"""``parl.utils.logger`` surface (parl/utils/logger.py:83-242): module-level info/warning/error,
set_dir / auto_set_dir / get_dir, ``DEBUG`` env var."""
import logging
import os
import sys


class _Logger(object):
    def __init__(self):
        self._log = logging.getLogger('parl_b200')
        self._log.propagate = False
        self._log.setLevel(logging.DEBUG if os.environ.get('DEBUG') else logging.INFO)
        if not self._log.handlers:
            h = logging.StreamHandler(sys.stdout)
            h.setFormatter(logging.Formatter('[%(asctime)s %(filename)s:%(lineno)d] %(levelname)s %(message)s',
                                             datefmt='%m-%d %H:%M:%S'))
            self._log.addHandler(h)
        self._dir = None
        self._file_handler = None
        for name in ('info', 'warning', 'error', 'critical', 'debug', 'exception'):
            setattr(self, name, getattr(self._log, name))
        self.warn = self._log.warning

    def set_level(self, level):
        self._log.setLevel(level)

    def set_dir(self, dirname):
        os.makedirs(dirname, exist_ok=True)
        if self._file_handler is not None:
            self._log.removeHandler(self._file_handler)
        self._dir = dirname
        self._file_handler = logging.FileHandler(os.path.join(dirname, 'log.log'), encoding='utf-8')
        self._file_handler.setFormatter(logging.Formatter('[%(asctime)s] %(levelname)s %(message)s'))
        self._log.addHandler(self._file_handler)

    def auto_set_dir(self, action=None):
        main = sys.modules.get('__main__')
        base = os.path.splitext(os.path.basename(getattr(main, '__file__', 'run')))[0]
        self.set_dir(os.path.join('train_log', base))

    def get_dir(self):
        return self._dir


logger = _Logger()
