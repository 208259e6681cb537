"""Transport registry.  The B200 build ships exactly one data plane, the NVLink/HBM transport
(``transport/hbm.py``); the enum keeps the reference's member names (transport/__init__.py:34-42)
so existing ``Strategy(default_transport_type=...)`` call sites still import, but selecting a
transport that is not part of this build fails loudly instead of falling back."""

from __future__ import annotations

from enum import Enum, auto
from typing import TYPE_CHECKING

from torchstore_b200.transport.buffers import TransportBuffer
from torchstore_b200.transport.types import Request, TensorSlice

if TYPE_CHECKING:
    from torchstore_b200.strategy import StorageVolumeRef


class TransportType(Enum):
    Unset = auto()  # resolved lazily: NVLink on this build
    MonarchRPC = auto()
    MonarchRDMA = auto()
    TorchComms = auto()
    TorchCommsRDMA = TorchComms
    Gloo = auto()
    SharedMemory = auto()
    NVLink = auto()  # HBM arena volumes + P2P copy_rects over NVSwitch (this repo)


def get_available_transport(storage_volume_ref: "StorageVolumeRef") -> TransportType:
    return TransportType.NVLink


def create_transport_buffer(storage_volume_ref: "StorageVolumeRef") -> TransportBuffer:
    transport_type = storage_volume_ref.default_transport_type
    if transport_type == TransportType.Unset:
        transport_type = get_available_transport(storage_volume_ref)
    if transport_type != TransportType.NVLink:
        raise RuntimeError(
            f"transport {transport_type.name} is not part of the B200 build; use TransportType.NVLink "
            "(or leave default_transport_type unset)"
        )
    from torchstore_b200.transport.hbm import HbmTransportBuffer

    return HbmTransportBuffer(storage_volume_ref)


__all__ = ["Request", "TensorSlice", "TransportType", "create_transport_buffer", "get_available_transport"]
