"""``@parl.remote_class`` façade — decorator contract of parl/remote/remote_decorator.py:25-113:
bare or with ``max_memory=``, ``wait=``, ``n_gpu=`` (any other kwarg -> AssertionError; decorating
a non-class -> AssertionError; instantiating before ``parl.connect`` -> AssertionError); names
starting ``_xparl`` are reserved; inside a hosted object (``XPARL`` env) the decorator is a no-op.

The decorated class is hosted IN THIS PROCESS next to the GPU it drives (the on-device actor pool
replaces remote CPU jobs), so a call is a plain method call instead of a cloudpickle+ZeroMQ round
trip (parl/remote/remote_wrapper.py:178-227).  ``wait=True``: calls run synchronously and an
exception surfaces as RemoteError; ``wait=False``: every call (and construction) is queued on the
object's own worker thread and returns a FutureObject (proxy_wrapper_nowait.py:32-230); attribute
reads go to the hosted instance, an instance attribute shadows a same-named method
(get_set_attribute_test.py:25-76)."""
import os
import queue
import threading
import traceback

from . import client as _client_mod
from .exceptions import RemoteError, RemoteAttributeError
from .future_object import FutureObject

XPARL_RESERVED_PREFIX = '_xparl'


class _WaitProxy(object):
    def __init__(self, cls, args, kwargs, max_memory, n_gpu):
        client = _client_mod.get_global_client()
        object.__setattr__(self, '_xparl_device', client.allocate_device(n_gpu))
        try:
            obj = cls(*args, **kwargs)
        except Exception:
            raise RemoteError('__init__', traceback.format_exc())
        object.__setattr__(self, '_xparl_obj', obj)

    def __getattr__(self, name):
        obj = object.__getattribute__(self, '_xparl_obj')
        if name in obj.__dict__:                      # instance attribute shadows a method of the same name
            return obj.__dict__[name]
        try:
            attr = getattr(obj, name)
        except AttributeError:
            raise RemoteAttributeError(name, traceback.format_exc())
        if not callable(attr):
            return attr

        def _call(*a, **k):
            try:
                return attr(*a, **k)
            except Exception:
                raise RemoteError(name, traceback.format_exc())
        return _call

    def __setattr__(self, name, value):
        setattr(object.__getattribute__(self, '_xparl_obj'), name, value)


class _NoWaitProxy(object):
    def __init__(self, cls, args, kwargs, max_memory, n_gpu):
        client = _client_mod.get_global_client()
        object.__setattr__(self, '_xparl_device', client.allocate_device(n_gpu))
        object.__setattr__(self, '_xparl_calls', queue.Queue())
        object.__setattr__(self, '_xparl_obj', None)
        object.__setattr__(self, '_xparl_init_error', None)
        object.__setattr__(self, '_xparl_closed', False)
        t = threading.Thread(target=self._xparl_loop, args=(cls, args, kwargs), daemon=True)
        object.__setattr__(self, '_xparl_thread', t)
        t.start()

    def _xparl_loop(self, cls, args, kwargs):
        try:
            object.__setattr__(self, '_xparl_obj', cls(*args, **kwargs))
        except Exception as e:
            object.__setattr__(self, '_xparl_init_error', e)
        calls = object.__getattribute__(self, '_xparl_calls')
        while True:
            item = calls.get()
            if item is None:
                return
            kind, name, a, k, fut = item
            err = object.__getattribute__(self, '_xparl_init_error')
            if err is not None:
                fut._set_exception(err)
                continue
            obj = object.__getattribute__(self, '_xparl_obj')
            try:
                if kind == 'call':
                    fut._set_result(getattr(obj, name)(*a, **k))
                elif kind == 'probe':
                    if name in obj.__dict__:
                        fut._set_result((True, obj.__dict__[name]))
                    else:
                        v = getattr(obj, name)
                        fut._set_result((not callable(v), v))
                else:
                    setattr(obj, name, a[0])
                    fut._set_result(None)
            except Exception as e:
                fut._set_exception(e)

    def _xparl_submit(self, kind, name, a=(), k=None):
        fut = FutureObject(name)
        object.__getattribute__(self, '_xparl_calls').put((kind, name, a, k or {}, fut))
        return fut

    def __getattr__(self, name):
        # like the reference (proxy_wrapper_nowait.py:150-202): wait for the calls queued so far, then
        # an attribute read returns the VALUE, a method returns a wrapper producing FutureObjects.
        is_attr, value = self._xparl_submit('probe', name).get()
        if is_attr:
            return value

        def _call(*a, **k):
            return self._xparl_submit('call', name, a, k)
        return _call

    def __setattr__(self, name, value):
        self._xparl_submit('set', name, (value, )).get()

    def destroy(self):
        if not object.__getattribute__(self, '_xparl_closed'):
            object.__setattr__(self, '_xparl_closed', True)
            object.__getattribute__(self, '_xparl_calls').put(None)


def remote_class(*args, **kwargs):
    def decorator(cls):
        assert isinstance(cls, type), "Only classes can be decorated by `parl.remote_class`."
        if os.environ.get('XPARL') == 'True':          # nested decoration inside a hosted object
            return cls
        for name in list(vars(cls)):
            assert not name.startswith(XPARL_RESERVED_PREFIX), \
                "attribute names starting with `{}` are reserved".format(XPARL_RESERVED_PREFIX)
        max_memory = kwargs.get('max_memory')
        n_gpu = kwargs.get('n_gpu', 0)
        wait = kwargs.get('wait', True)
        proxy_base = _WaitProxy if wait else _NoWaitProxy

        class RemoteProxy(proxy_base):
            _original = cls

            def __init__(self, *a, **k):
                proxy_base.__init__(self, cls, a, k, max_memory, n_gpu)

        RemoteProxy.__name__ = cls.__name__
        RemoteProxy.__qualname__ = getattr(cls, '__qualname__', cls.__name__)
        RemoteProxy.__doc__ = cls.__doc__
        RemoteProxy.__module__ = cls.__module__
        return RemoteProxy

    if len(args) == 1 and len(kwargs) == 0 and callable(args[0]):
        assert isinstance(args[0], type), "Only classes can be decorated by `parl.remote_class`."
        return decorator(args[0])
    assert len(args) == 0, "`parl.remote_class` takes keyword arguments only"
    for key in kwargs:
        assert key in ('max_memory', 'wait', 'n_gpu'), \
            "unsupported argument `{}` for parl.remote_class (max_memory, wait, n_gpu)".format(key)
    return decorator
