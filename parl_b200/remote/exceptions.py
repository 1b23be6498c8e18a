"""Exception surface of parl/remote/exceptions.py:16-110 kept for drop-in scripts."""
import traceback as _tb


class ResourceError(Exception):
    """No capacity left to host another remote object (reference: no vacant CPU in the cluster)."""


class RemoteError(Exception):
    """An exception raised inside a remote object's __init__ or method (wait mode)."""

    def __init__(self, func_name, error_info):
        self.func_name, self.error_info = func_name, error_info
        super(RemoteError, self).__init__(func_name, error_info)

    def __str__(self):
        return "[PARL remote error when calling function `{}`]:\n{}".format(self.func_name, self.error_info)


class FutureFunctionError(Exception):
    """Raised by ``future.get()`` when the remote call failed (future mode)."""

    def __init__(self, func_name):
        self.func_name = func_name
        super(FutureFunctionError, self).__init__(func_name)

    def __str__(self):
        return "[PARL remote error when calling function `{}`]".format(self.func_name)


class RemoteSerializeError(RemoteError):
    pass


class RemoteDeserializeError(RemoteError):
    pass


class RemoteAttributeError(RemoteError):
    pass


class FutureGetRepeatedlyError(Exception):
    def __init__(self, func_name):
        super(FutureGetRepeatedlyError, self).__init__(
            "Calling `get` function of the FutureObject returned by `{}` repeatedly is not allowed.".format(func_name))


class FutureObjectEmpty(Exception):
    def __init__(self, func_name):
        super(FutureObjectEmpty, self).__init__(
            "The FutureObject returned by `{}` is not ready yet.".format(func_name))


def format_current_exception():
    return _tb.format_exc()
