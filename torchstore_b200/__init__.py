"""torchstore_b200 -- B200-native weight-sync path behind torchstore's API.

``import torchstore_b200 as ts`` exposes the names of the reference package
(torchstore/__init__.py:47-69): initialize / initialize_spmd / put / get / put_batch / get_batch /
delete / delete_batch / keys / exists / client / shutdown / put_state_dict / get_state_dict /
LocalRankStrategy / HostStrategy / ControllerStorageVolumes / TorchStoreStrategy / reset_client.
"""

from torchstore_b200 import spmd
from torchstore_b200.api import (
    client,
    delete,
    delete_batch,
    exists,
    get,
    get_batch,
    get_state_dict,
    initialize,
    keys,
    put,
    put_batch,
    put_state_dict,
    reset_client,
    shutdown,
)
from torchstore_b200.logging import init_logging
from torchstore_b200.strategy import ControllerStorageVolumes, HostStrategy, LocalRankStrategy, TorchStoreStrategy
from torchstore_b200.transport import TransportType
from torchstore_b200.transport.types import Request, TensorSlice

initialize_spmd = spmd.initialize

__all__ = [
    "initialize",
    "initialize_spmd",
    "init_logging",
    "put",
    "put_batch",
    "get",
    "get_batch",
    "delete",
    "delete_batch",
    "keys",
    "exists",
    "client",
    "shutdown",
    "TorchStoreStrategy",
    "LocalRankStrategy",
    "HostStrategy",
    "ControllerStorageVolumes",
    "put_state_dict",
    "get_state_dict",
    "reset_client",
    "spmd",
    "TransportType",
    "TensorSlice",
    "Request",
]
