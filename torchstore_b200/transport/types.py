"""Request / TensorSlice: the metadata that travels with every put/get.

Same public surface and semantics as the reference's ``torchstore/transport/types.py``
(TensorSlice :20-55, _is_dtensor_fully_local :58-85, Request :88-218); the implementation is
ours.  A ``TensorSlice`` is one axis-aligned hyper-rectangle of a global tensor -- exactly what
the reshard kernel moves -- so this module is also where DTensor layouts become rectangles.
"""

from __future__ import annotations

import copy
from dataclasses import dataclass
from logging import getLogger
from typing import Any

import torch
from torch.distributed.tensor import DTensor
from torch.distributed.tensor._utils import _compute_local_shape_and_global_offset
from torch.distributed.tensor.placement_types import Replicate

logger = getLogger(__name__)


@dataclass
class TensorSlice:
    """One shard's rectangle inside the global tensor.

    offsets / local_shape locate the rectangle, global_shape is the full tensor, coordinates /
    mesh_shape identify which member of the device mesh owns it (used by the controller to
    decide when a sharded key is fully committed).
    """

    offsets: tuple
    coordinates: tuple
    global_shape: tuple
    local_shape: tuple
    mesh_shape: tuple

    def __post_init__(self):
        if self.coordinates is not None:
            self.coordinates = tuple(self.coordinates)

    def _key(self):
        shape = self.local_shape
        if hasattr(shape, "__iter__"):
            shape = tuple(shape)
        return (self.offsets, self.coordinates, self.global_shape, shape, self.mesh_shape)

    def __hash__(self):
        return hash(self._key())

    # convenience used by the planner (not part of the reference surface)
    def end(self, dim: int) -> int:
        return self.offsets[dim] + self.local_shape[dim]

    @property
    def ndim(self) -> int:
        return len(self.global_shape)


def _is_dtensor_fully_local(dtensor: DTensor) -> bool:
    """A DTensor that is not really distributed: one-device mesh, or Replicate() everywhere.
    Such tensors are stored as plain tensors (reference types.py:58-85)."""
    if dtensor.device_mesh.size() == 1:
        return True
    return all(isinstance(p, Replicate) for p in dtensor.placements)


@dataclass
class Request:
    """What a client asks a storage volume to store or return for one key."""

    key: str = ""
    tensor_val: torch.Tensor | None = None
    tensor_slice: TensorSlice | None = None
    objects: Any | None = None
    is_object: bool = False

    @classmethod
    def from_any(cls, key: str, value: "torch.Tensor | DTensor | None", tensor_slice: TensorSlice | None = None) -> "Request":
        """Tensor / DTensor / None -> Request (objects go through ``from_objects``)."""
        if isinstance(value, DTensor):
            if tensor_slice is not None:
                raise ValueError(
                    "Cannot specify tensor_slice with a DTensor since DTensor already has its own sharding info."
                )
            if _is_dtensor_fully_local(value):
                logger.debug("DTensor %s is fully local; storing as a regular tensor", tuple(value.shape))
                return cls.from_tensor(key, value._local_tensor)
            return cls.from_dtensor(key, value)
        if isinstance(value, torch.Tensor):
            if tensor_slice is not None and tensor_slice.local_shape != value.shape:
                raise ValueError(
                    f"Requested tensor slice shape {tensor_slice.local_shape} does not match tensor shape {value.shape}"
                )
            req = cls.from_tensor(key, value)
            req.tensor_slice = tensor_slice
            return req
        if value is None:
            return cls(key=key, tensor_slice=tensor_slice)
        raise TypeError(
            f"from_any accepts None, torch.Tensor, or DTensor, got {type(value)}. "
            "For arbitrary objects, use Request.from_objects() instead."
        )

    @classmethod
    def from_dtensor(cls, key: str, dtensor: DTensor) -> "Request":
        mesh = dtensor.device_mesh
        coordinates = mesh.get_coordinate()
        _, offsets = _compute_local_shape_and_global_offset(
            dtensor.shape, mesh_shape=mesh.shape, my_coordinate=coordinates, placements=dtensor.placements
        )
        local = dtensor._local_tensor
        return cls(
            key=key,
            tensor_val=local,
            tensor_slice=TensorSlice(offsets, coordinates, dtensor.shape, local.shape, mesh.shape),
        )

    @classmethod
    def from_tensor(cls, key: str, tensor: torch.Tensor) -> "Request":
        return cls(key=key, tensor_val=tensor)

    @classmethod
    def from_objects(cls, key: str, objects) -> "Request":
        return cls(key=key, objects=objects, is_object=True)

    @classmethod
    def from_tensor_slice(cls, key: str, tensor_slice: TensorSlice) -> "Request":
        return cls(key=key, tensor_slice=copy.deepcopy(tensor_slice))

    def meta_only(self) -> "Request":
        """Copy without the tensor payload (what the controller and the volume RPCs carry)."""
        return Request(key=self.key, tensor_val=None, tensor_slice=self.tensor_slice, objects=self.objects,
                       is_object=self.is_object)
