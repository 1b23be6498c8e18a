"""FutureObject — contract of parl/remote/future_mode/future_object.py:57-97: ``get(block, timeout)``,
``get_nowait()``, ``empty()``; a second ``get`` raises FutureGetRepeatedlyError; a failed remote call
surfaces as FutureFunctionError on ``get``."""
import queue

from .exceptions import FutureFunctionError, FutureGetRepeatedlyError, FutureObjectEmpty


class _Failure(object):
    def __init__(self, exc):
        self.exc = exc


class FutureObject(object):
    def __init__(self, func_name):
        self._func_name = func_name
        self._q = queue.Queue(maxsize=1)
        self._already_get = False

    # producer side
    def _set_result(self, value):
        self._q.put(value)

    def _set_exception(self, exc):
        self._q.put(_Failure(exc))

    # consumer side
    def get(self, block=True, timeout=None):
        if self._already_get:
            raise FutureGetRepeatedlyError(self._func_name)
        try:
            result = self._q.get(block=block, timeout=timeout)
        except queue.Empty:
            raise FutureObjectEmpty(self._func_name)
        self._already_get = True
        if isinstance(result, _Failure):
            raise FutureFunctionError(self._func_name) from result.exc
        return result

    def get_nowait(self):
        return self.get(block=False)

    def empty(self):
        return self._q.empty()
