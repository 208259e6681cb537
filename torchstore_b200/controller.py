"""The metadata index: which volume holds which key, in which form, and whether a sharded key has
been put by every mesh coordinate (reference torchstore/controller.py:22-296).

On one box this is a plain object served by the process that called ``initialize`` (rank 0 under
SPMD); every endpoint is microseconds of dict work, so it never sits on the bandwidth path.
"""

from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from enum import Enum, auto
from itertools import product

from torchstore_b200.rpc import Actor, ActorMesh, endpoint
from torchstore_b200.strategy import ControllerStorageVolumes, TorchStoreStrategy
from torchstore_b200.transport.types import Request, TensorSlice


class ObjectType(Enum):
    OBJECT = auto()
    TENSOR = auto()
    TENSOR_SLICE = auto()

    @classmethod
    def from_request(cls, request: Request) -> "ObjectType":
        if request.is_object:
            return cls.OBJECT
        return cls.TENSOR_SLICE if request.tensor_slice is not None else cls.TENSOR


@dataclass
class StorageInfo:
    object_type: ObjectType
    tensor_slices: set[TensorSlice | None] = field(default_factory=set)

    def update(self, other: "StorageInfo") -> None:
        assert self.object_type == other.object_type, (
            "Particularly dangerous to change storage type of an existing key, are you sure? Raise an issue if so."
        )
        self.tensor_slices.update(other.tensor_slices)


class KeyIndex(dict):
    """key -> {volume_id -> StorageInfo}; prefix queries follow the reference's '.'-separated trie
    (storage_utils/trie.py:35-42): a prefix matches whole path components."""

    def filter_by_prefix(self, prefix: str) -> list[str]:
        want = prefix.split(".")
        n = len(want)
        return [k for k in self if k.split(".")[:n] == want]


class Controller(Actor):
    def __init__(self) -> None:
        self.keys_to_storage_volumes = KeyIndex()
        self.is_initialized = False
        self.strategy: TorchStoreStrategy | None = None
        self.storage_volumes = None
        self.num_storage_volumes: int | None = None
        # bumped whenever the index changes SHAPE (new key / volume / slice, any delete); overwriting
        # an indexed key in place leaves it alone.  Clients validate cached get plans against it.
        self.epoch = 0
        self._board = None  # EpochBoard owned by this controller (slot 0 = self.epoch)

    def _bump(self) -> None:
        self.epoch += 1
        if self._board is not None:
            self._board.write(0, self.epoch)

    def assert_initialized(self) -> None:
        assert self.is_initialized, "Please call torchstore.initialize before attempting to use store."

    def _is_dtensor_fully_committed(self, key: str, volume_map: dict[str, StorageInfo]) -> bool:
        """Every coordinate of the put-side mesh has stored its shard."""
        have = set()
        mesh_shape = None
        for info in volume_map.values():
            if info.object_type != ObjectType.TENSOR_SLICE:
                return True
            for ts in info.tensor_slices:
                have.add(ts.coordinates)
                if mesh_shape is None:
                    mesh_shape = ts.mesh_shape
                else:
                    assert mesh_shape == ts.mesh_shape, "Inconsistent mesh shapes in stored slices"
        return have == set(product(*(range(s) for s in mesh_shape)))

    @endpoint
    async def init(self, strategy: TorchStoreStrategy, num_storage_volumes: int, storage_volumes) -> None:
        if self.is_initialized:
            raise RuntimeError("TorchStore is already initialized")
        if isinstance(strategy, ControllerStorageVolumes) and num_storage_volumes > 1:
            warnings.warn("ControllerStorageVolumes serves a single volume", DeprecationWarning, stacklevel=2)
        self.strategy = strategy
        self.storage_volumes = storage_volumes
        self.num_storage_volumes = num_storage_volumes
        await self.strategy.set_storage_volumes(self.storage_volumes)
        await self._publish_epoch_board()
        self.is_initialized = True

    async def _publish_epoch_board(self) -> None:
        """One shm page with this controller's epoch and every volume's (epoch_board.py): clients on
        the box validate replayable sessions without an RPC.  Best effort."""
        from torchstore_b200.epoch_board import MAX_SLOTS, EpochBoard

        volume_ids = sorted(self.strategy.volume_id_to_coord)
        if len(volume_ids) + 1 > MAX_SLOTS:
            return
        board = EpochBoard.create()
        if board is None:
            return
        board.write(0, self.epoch)
        slots = {vid: i + 1 for i, vid in enumerate(volume_ids)}
        try:
            for vid, slot in slots.items():
                # an ActorMesh is addressed by coordinate; ControllerStorageVolumes hands over one bare ref
                ref = self.storage_volumes.slice(**self.strategy.volume_id_to_coord[vid]) if isinstance(self.storage_volumes, ActorMesh) \
                    else self.storage_volumes
                ok = await ref.attach_epoch_board.call_one(board.name, slot)
                if not ok:
                    raise RuntimeError(f"volume {vid} could not attach")
        except Exception as e:  # a volume on another host, an old volume ...: stay on RPC epochs
            import logging

            logging.getLogger(__name__).debug("epoch board disabled: %s", e)
            board.close()
            return
        self._board = board
        self.strategy.epoch_board = (board.name, slots)

    @endpoint
    async def get_controller_strategy(self) -> TorchStoreStrategy:
        self.assert_initialized()
        assert self.strategy is not None, "Strategy is not set"
        return self.strategy

    @endpoint
    async def locate_volumes(self, keys: list[str], missing_ok: bool = False,
                             require_fully_committed: bool = True) -> dict[str, dict[str, StorageInfo]]:
        self.assert_initialized()
        found = {}
        for key in keys:
            volume_map = self.keys_to_storage_volumes.get(key)
            if volume_map is None:
                if missing_ok:
                    continue
                raise KeyError(f"Unable to locate {key} in any storage volumes.")
            if require_fully_committed and not self._is_dtensor_fully_committed(key, volume_map):
                raise KeyError(
                    f"DTensor '{key}' is only partially committed. Not all shards have been stored yet. "
                    "Please ensure all ranks complete their put() operations."
                )
            found[key] = volume_map
        return found

    @endpoint
    async def notify_put_batch(self, requests: list[Request], storage_volume_id: str) -> None:
        self.assert_initialized()
        for request in requests:
            self._notify_put(request, storage_volume_id)

    def _notify_put(self, request: Request, storage_volume_id: str) -> None:
        assert request.tensor_val is None, (
            "request should not contain tensor data, as this will significantly increase e2e latency"
        )
        volume_map = self.keys_to_storage_volumes.setdefault(request.key, {})
        info = StorageInfo(ObjectType.from_request(request), {request.tensor_slice})
        if storage_volume_id in volume_map:
            known = volume_map[storage_volume_id]
            if request.tensor_slice not in known.tensor_slices:
                self._bump()
            known.update(info)
        else:
            volume_map[storage_volume_id] = info
            self._bump()

    @endpoint
    async def get_epoch(self) -> int:
        return self.epoch

    @endpoint
    async def teardown(self) -> None:
        self.is_initialized = False
        self.keys_to_storage_volumes = KeyIndex()
        self._bump()
        self.strategy = None
        if self.storage_volumes is not None:
            await self.storage_volumes.reset.call()
        if self._board is not None:
            from torchstore_b200 import epoch_board

            epoch_board.forget(self._board.name)  # this process's reader attachment (client + local volumes)
            self._board.close()                    # the owner's mapping; unlinks the segment
            self._board = None
        self.storage_volumes = None
        self.num_storage_volumes = None

    @endpoint
    async def keys(self, prefix=None) -> list[str]:
        if prefix is None:
            return list(self.keys_to_storage_volumes.keys())
        return self.keys_to_storage_volumes.filter_by_prefix(prefix)

    @endpoint
    async def notify_delete(self, key: str, storage_volume_id: str) -> None:
        self.assert_initialized()
        self._notify_delete(key, storage_volume_id)

    def _notify_delete(self, key: str, storage_volume_id: str, missing_ok: bool = False) -> None:
        volume_map = self.keys_to_storage_volumes.get(key)
        if volume_map is None:
            if missing_ok:
                return
            raise KeyError(f"Unable to locate {key} in any storage volumes.")
        if storage_volume_id not in volume_map:
            if missing_ok:
                return
            raise KeyError(f"Unable to locate {key} in storage volume {storage_volume_id}.")
        del volume_map[storage_volume_id]
        self._bump()
        if not volume_map:
            del self.keys_to_storage_volumes[key]

    @endpoint
    async def notify_delete_batch(self, volume_to_keys: dict[str, list[str]]) -> None:
        self.assert_initialized()
        for volume_id, keys in volume_to_keys.items():
            for key in keys:
                self._notify_delete(key, volume_id, missing_ok=True)

    def get_keys_to_storage_volumes(self):
        return self.keys_to_storage_volumes
