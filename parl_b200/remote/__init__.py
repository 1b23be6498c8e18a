from .exceptions import (RemoteError, ResourceError, RemoteSerializeError, RemoteDeserializeError,
                         RemoteAttributeError, FutureFunctionError, FutureGetRepeatedlyError,
                         FutureObjectEmpty)
from .remote_decorator import remote_class
from .client import connect, disconnect, get_global_client
from .future_object import FutureObject

__all__ = ['remote_class', 'connect', 'disconnect', 'FutureObject', 'RemoteError', 'ResourceError',
           'RemoteSerializeError', 'RemoteDeserializeError', 'RemoteAttributeError', 'FutureFunctionError',
           'FutureGetRepeatedlyError', 'FutureObjectEmpty']
