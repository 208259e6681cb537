"""Index math of the reshard path: rectangle intersection, destination views, gather.

Public functions mirror the reference's ``torchstore/utils.py`` (to_byte_view :25-33,
get_destination_view :36-98, tensors_overlap_in_memory :101-120, get_local_tensor :142-155,
assemble_tensor :158-212, get_target_tensor_shape_and_offset :215-245,
get_slice_intersection :248-307).  These are metadata-only; bytes move in csrc/copy_rects.cu.
"""

from __future__ import annotations

import math
import os
import socket
from logging import getLogger
from typing import TYPE_CHECKING, Sequence

import torch

if TYPE_CHECKING:
    from torchstore_b200.transport.types import TensorSlice

logger = getLogger(__name__)


def to_byte_view(tensor: torch.Tensor) -> torch.Tensor:
    """Flat uint8 alias of a (contiguous) tensor; 0-d tensors are unsqueezed first."""
    if tensor.dim() == 0:
        tensor = tensor.unsqueeze(0)
    return tensor.view(torch.uint8).flatten()


def get_local_hostname() -> str:
    return os.environ.get("HOSTNAME", socket.gethostname())


def _box(offsets: Sequence[int], shape: Sequence[int]) -> tuple[slice, ...]:
    return tuple(slice(o, o + s) for o, s in zip(offsets, shape, strict=True))


def get_local_tensor(global_tensor: torch.Tensor, local_shape, global_offset) -> torch.Tensor:
    """The view of ``global_tensor`` a shard with this shape/offset covers."""
    return global_tensor[_box(global_offset, local_shape)]


def get_slice_intersection(tensor_slice: "TensorSlice", dtensor_slice: "TensorSlice") -> "TensorSlice | None":
    """Overlap of a stored rectangle with a wanted rectangle, or None.

    The result keeps the *stored* slice's coordinates and mesh_shape (it names a sub-rectangle of
    that shard).  Different global shapes never intersect.
    """
    from torchstore_b200.transport.types import TensorSlice

    if tensor_slice.global_shape != dtensor_slice.global_shape:
        return None
    lo, extent = [], []
    for dim in range(len(tensor_slice.global_shape)):
        start = max(tensor_slice.offsets[dim], dtensor_slice.offsets[dim])
        stop = min(
            tensor_slice.offsets[dim] + tensor_slice.local_shape[dim],
            dtensor_slice.offsets[dim] + dtensor_slice.local_shape[dim],
        )
        if stop <= start:
            return None
        lo.append(start)
        extent.append(stop - start)
    return TensorSlice(
        offsets=tuple(lo),
        coordinates=tensor_slice.coordinates,
        global_shape=tensor_slice.global_shape,
        local_shape=tuple(extent),
        mesh_shape=tensor_slice.mesh_shape,
    )


def local_box(outer: "TensorSlice | None", outer_shape, inner: "TensorSlice") -> tuple[slice, ...] | None:
    """Index of rectangle ``inner`` (global coordinates) inside a tensor that holds rectangle
    ``outer`` (None: the tensor *is* the global tensor).  None if ``inner`` sticks out."""
    idx = []
    for dim in range(len(inner.global_shape)):
        base = outer.offsets[dim] if outer is not None else 0
        limit = outer.local_shape[dim] if outer is not None else outer_shape[dim]
        start = inner.offsets[dim] - base
        stop = start + inner.local_shape[dim]
        if start < 0 or stop > limit:
            return None
        idx.append(slice(start, stop))
    return tuple(idx)


def get_destination_view(dest_tensor: torch.Tensor, dest_slice: "TensorSlice | None", fetch_slice: "TensorSlice"):
    """View of ``dest_tensor`` where ``fetch_slice`` lands, or None when it is out of bounds, the
    destination is not contiguous, or the view is not contiguous (the reference's rule for
    in-place transports, utils.py:94-96)."""
    if not dest_tensor.is_contiguous():
        return None
    idx = local_box(dest_slice, tuple(dest_tensor.shape), fetch_slice)
    if idx is None:
        return None
    view = dest_tensor[idx]
    return view if view.is_contiguous() else None


def get_destination_region(dest_tensor: torch.Tensor, dest_slice: "TensorSlice | None", fetch_slice: "TensorSlice"):
    """Like get_destination_view but without the contiguity restriction: the reshard kernel writes
    strided sub-rectangles directly, so every in-bounds region is an in-place target."""
    idx = local_box(dest_slice, tuple(dest_tensor.shape), fetch_slice)
    return None if idx is None else dest_tensor[idx]


def tensors_overlap_in_memory(tensors, base_tensor: torch.Tensor) -> bool:
    """True when every (tensor, meta) pair starts inside ``base_tensor``'s bytes."""
    if not tensors:
        return False
    lo = base_tensor.data_ptr()
    hi = lo + base_tensor.nbytes
    return all(lo <= t.data_ptr() < hi for t, _ in tensors)


def get_target_tensor_shape_and_offset(local_tensor_shapes, global_offsets):
    """Bounding box of the parts: (shape, offset of its origin in global coordinates)."""
    target_offset = min(global_offsets)
    ndim = len(global_offsets[0])
    ends = [max(off[d] + shp[d] for off, shp in zip(global_offsets, local_tensor_shapes)) for d in range(ndim)]
    target_shape = [max(0, end - start) for start, end in zip(target_offset, ends)]
    have = sum(math.prod(s) for s in local_tensor_shapes)
    need = math.prod(target_shape)
    assert have >= need, (
        "Local tensor sizes doesn't match target tensor. "
        f"Local tensors total size: {have}, Target tensor size: {need}"
    )
    return target_shape, target_offset


def assemble_tensor(local_tensors: list[torch.Tensor], global_offsets, device=None) -> torch.Tensor:
    """Gather parts into their bounding box (later parts win on overlap).

    The reference always builds the result on the CPU (utils.py:199-202); ``device`` lets the
    HBM path keep it on the GPU that holds the parts.
    """
    assert local_tensors
    shape, origin = get_target_tensor_shape_and_offset([t.shape for t in local_tensors], global_offsets)
    out = torch.empty(shape, dtype=local_tensors[0].dtype, device=device)
    for part, offset in zip(local_tensors, global_offsets, strict=True):
        out[_box([o - b for o, b in zip(offset, origin, strict=True)], part.shape)] = part
    return out
