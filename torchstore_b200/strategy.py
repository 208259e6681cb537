"""Client -> storage-volume mapping (reference torchstore/strategy.py:29-245).

Kept surface: ``TorchStoreStrategy`` (get_volume_id / get_client_id / select_storage_volume /
get_storage_volume / set_storage_volumes), ``LocalRankStrategy``, ``HostStrategy``,
``ControllerStorageVolumes``, ``StorageVolumeRef`` and the ``default_transport_type`` ctor arg.
On one NVSwitch box a "volume" is an HBM arena owned by one rank's process (LocalRankStrategy) or
by the box (HostStrategy).
"""

from __future__ import annotations

import logging
import os
import socket
from typing import TYPE_CHECKING

from torchstore_b200.transport import TransportType
from torchstore_b200.transport.buffers import TransportContext

if TYPE_CHECKING:
    from torchstore_b200.storage_volume import StorageVolume

logger = logging.getLogger(__name__)

# rank of the process whose volume is being constructed (monarch.actor.current_rank stand-in)
_spawn_rank: list[int] = [0]


def current_spawn_rank() -> int:
    return _spawn_rank[0]


class StorageVolumeRef:
    __slots__ = ("volume", "volume_id", "transport_context", "default_transport_type", "volume_hostname", "epoch_board")

    def __init__(self, volume: "StorageVolume", volume_id: str, transport_context: TransportContext,
                 default_transport_type: TransportType, volume_hostname: str | None = None, epoch_board=None):
        self.volume = volume
        self.volume_id = volume_id
        self.transport_context = transport_context  # survives across requests: caches live here
        self.default_transport_type = default_transport_type
        self.volume_hostname = volume_hostname
        self.epoch_board = epoch_board  # (shm name, {volume_id: slot}) or None


class TorchStoreStrategy:
    """Assigns volume ids, maps each client process to its volume, hands out volume refs."""

    def __init__(self, default_transport_type: TransportType = TransportType.Unset):
        self.default_transport_type = default_transport_type
        logger.info("Initializing %s with default_transport_type=%s", type(self).__name__, default_transport_type)
        self.storage_volumes = None
        self.volume_id_to_coord: dict = {}
        self.volume_id_to_hostname: dict = {}
        self.transport_context = TransportContext()
        self.epoch_board = None  # (shm name, {volume_id: slot}) published by the controller, or None

    def __str__(self) -> str:
        n = len(self.storage_volumes) if self.storage_volumes is not None else 0
        return f"{self.__class__.__name__}(storage_volume_len={n})"

    def __getstate__(self):
        # the strategy travels controller -> client; caches are per process
        state = self.__dict__.copy()
        state["transport_context"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self.transport_context = TransportContext()

    @classmethod
    def get_volume_id(cls):
        raise NotImplementedError(f"{cls.__name__} must implement 'get_volume_id'")

    @classmethod
    def get_client_id(cls):
        raise NotImplementedError(f"{cls.__name__} must implement 'get_client_id'")

    async def set_storage_volumes(self, storage_volumes) -> None:
        """Ask every volume of the mesh who it is: [(coord, (volume_id, hostname)), ...]."""
        listing = await storage_volumes.get_id.call()
        self.storage_volumes = storage_volumes
        self.volume_id_to_coord = {vid: coord for coord, (vid, _host) in listing}
        self.volume_id_to_hostname = {vid: host for _coord, (vid, host) in listing}

    def _ref(self, actor, volume_id: str) -> StorageVolumeRef:
        return StorageVolumeRef(actor, volume_id, self.transport_context, self.default_transport_type,
                                volume_hostname=self.volume_id_to_hostname.get(volume_id),
                                epoch_board=getattr(self, "epoch_board", None))

    def select_storage_volume(self) -> StorageVolumeRef:
        mine = self.get_client_id()  # a client's id is the id of the volume it writes to
        if mine not in self.volume_id_to_coord:
            raise KeyError(f"No corresponding storage volume found for {mine} {self.volume_id_to_coord=}")
        return self.get_storage_volume(mine)

    def get_storage_volume(self, volume_id: str) -> StorageVolumeRef:
        return self._ref(self.storage_volumes.slice(**self.volume_id_to_coord[volume_id]), volume_id)


class HostStrategy(TorchStoreStrategy):
    """One volume per host, addressed by $HOSTNAME."""

    @classmethod
    def get_volume_id(cls):
        return os.environ.get("HOSTNAME", socket.gethostname())

    @classmethod
    def get_client_id(cls):
        return os.environ["HOSTNAME"]


class LocalRankStrategy(TorchStoreStrategy):
    """One volume per rank.  Volumes are numbered by the rank of the process that hosts them;
    clients use $RANK when set, else $LOCAL_RANK (reference strategy.py:164-188)."""

    @classmethod
    def get_volume_id(cls):
        return str(current_spawn_rank())

    @classmethod
    def get_client_id(cls):
        rank = os.environ.get("RANK")
        return rank if rank is not None else os.environ["LOCAL_RANK"]


class ControllerStorageVolumes(TorchStoreStrategy):
    """Single volume living next to the controller (default when num_storage_volumes == 1)."""

    @classmethod
    def get_volume_id(cls):
        return "0"

    @classmethod
    def get_client_id(cls):
        return "0"

    async def set_storage_volumes(self, storage_volumes) -> None:
        vid, host = await storage_volumes.get_id.call_one()
        self.storage_volumes = storage_volumes
        self.volume_id_to_coord = {"0": {}}  # one bare actor ref, no mesh coordinate
        self.volume_id_to_hostname = {vid: host}

    def get_storage_volume(self, volume_id: str) -> StorageVolumeRef:
        return self._ref(self.storage_volumes, volume_id)
