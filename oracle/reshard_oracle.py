"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the weight-sync hot path (numpy, no torch).

A restatement, in plain numpy / Python integers, of the reference algorithms on the path
(paths relative to the reference root, commit ed2ddb67):

  * shard layout            torch.distributed.tensor._utils._compute_local_shape_and_global_offset
                            as called from torchstore/transport/types.py:176-191
  * slice intersection      torchstore/utils.py:248-307   (get_slice_intersection)
  * destination view        torchstore/utils.py:36-98     (get_destination_view)
  * assemble                torchstore/utils.py:158-245   (assemble_tensor)
  * direct-sync plan        torchstore/direct_weight_sync.py:221-317 (_build_plan)
  * direct-sync pull        torchstore/direct_weight_sync.py:319-350 (+ MockRDMABuffer semantics,
                            tests/test_direct_weight_sync.py:27-37)
  * store put / reshard get torchstore/client.py:239-373, storage_volume.py:239-359,
                            transport/shared_memory.py:328-380,438-480
  * fp32->bf16 cast         torch .to(bfloat16) (c10 round_to_nearest_even)

PARITY PINNING: every function here is checked against outputs of the *real* reference
executed in the build container (oracle/gen_golden.py -> tests/golden/*.json|npz) by
tests/test_oracle_golden.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product package (torchstore_b200/) never does.
"""

from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------------------------
# slices
# --------------------------------------------------------------------------------------------


@dataclass(frozen=True)
class Slice:
    """The reference's TensorSlice (transport/types.py:20-55) as an immutable value."""

    offsets: tuple
    coordinates: tuple | None
    global_shape: tuple
    local_shape: tuple
    mesh_shape: tuple | None


def shard_layout(global_shape, mesh_shape, coordinate, placements):
    """(local_shape, global_offset) of one mesh coordinate.

    placements: one entry per mesh dim, either ("S", tensor_dim) or ("R",).
    torch semantics (torch.chunk style): a dim of size n sharded k ways uses chunks of
    ceil(n/k); trailing shards may be short or empty; nested shards of the same tensor dim
    apply to the already-sharded extent, offsets accumulate.
    """
    local = list(global_shape)
    offset = [0] * len(global_shape)
    empty = [False] * len(global_shape)
    for mesh_dim, placement in enumerate(placements):
        if placement[0] != "S":
            continue
        dim = placement[1]
        k = mesh_shape[mesh_dim]
        idx = coordinate[mesh_dim]
        n = local[dim]
        chunk = -(-n // k) if k else 0
        start = min(n, chunk * idx)
        size = max(0, min(n, chunk * (idx + 1)) - start)
        if size == 0:
            empty[dim] = True
        local[dim] = size
        offset[dim] += start
    for dim, is_empty in enumerate(empty):
        if is_empty:
            # torch reports an empty shard at the end of the GLOBAL dimension ("zero_global_offset")
            offset[dim] = global_shape[dim]
    return tuple(local), tuple(offset)


def make_slice(global_shape, mesh_shape, coordinate, placements) -> Slice:
    local, off = shard_layout(global_shape, mesh_shape, coordinate, placements)
    return Slice(tuple(off), tuple(coordinate), tuple(global_shape), tuple(local), tuple(mesh_shape))


def full_slice(shape) -> Slice:
    """Synthetic slice of a plain tensor (direct_weight_sync.py:61-74)."""
    shape = tuple(shape)
    n = len(shape)
    return Slice((0,) * n, (0,) * n, shape, shape, (1,) * n)


def slice_intersection(stored: Slice, wanted: Slice) -> Slice | None:
    if stored.global_shape != wanted.global_shape:
        return None
    offs, shape = [], []
    for d in range(len(stored.global_shape)):
        lo = max(stored.offsets[d], wanted.offsets[d])
        hi = min(stored.offsets[d] + stored.local_shape[d], wanted.offsets[d] + wanted.local_shape[d])
        if lo >= hi:
            return None
        offs.append(lo)
        shape.append(hi - lo)
    return Slice(tuple(offs), stored.coordinates, stored.global_shape, tuple(shape), stored.mesh_shape)


def _contiguous_strides(shape):
    st, s = [], 1
    for e in reversed(shape):
        st.append(s)
        s *= e
    return tuple(reversed(st))


def _view_is_contiguous(parent_shape, idx):
    """torch's is_contiguous() of parent[idx] for a contiguous parent and unit-step slices."""
    shape = [sl.stop - sl.start for sl in idx]
    strides = _contiguous_strides(parent_shape)
    expected = 1
    for e, st in zip(reversed(shape), reversed(strides)):
        if e == 1:
            continue
        if st != expected:
            return False
        expected *= e
    return True


def destination_view(dest_shape, dest_is_contiguous, dest_slice: Slice | None, fetch: Slice):
    """Index (tuple of slices) into the destination tensor where `fetch` lands, or None."""
    if not dest_is_contiguous:
        return None
    dest_shape = tuple(dest_shape)
    if dest_slice is None:
        dest_slice = Slice((0,) * len(dest_shape), None, dest_shape, dest_shape, None)
    idx = []
    for d in range(len(fetch.global_shape)):
        lo = fetch.offsets[d] - dest_slice.offsets[d]
        hi = lo + fetch.local_shape[d]
        if lo < 0 or hi > dest_slice.local_shape[d]:
            return None
        idx.append(slice(lo, hi))
    idx = tuple(idx)
    if any(e == 0 for e in dest_shape):
        return idx
    if not _view_is_contiguous(dest_shape, idx):
        return None
    return idx


def target_shape_and_offset(shapes, offsets):
    target_offset = min(tuple(o) for o in offsets)
    ends = tuple(max(o[i] + s[i] for o, s in zip(offsets, shapes)) for i in range(len(offsets[0])))
    shape = [max(0, e - o) for o, e in zip(target_offset, ends)]
    assert sum(math.prod(s) for s in shapes) >= math.prod(shape), "Local tensor sizes doesn't match target tensor."
    return shape, target_offset


def assemble(parts, offsets):
    """utils.py:158-212; later parts overwrite earlier ones; result dtype = parts[0].dtype."""
    assert parts
    shape, toff = target_shape_and_offset([p.shape for p in parts], offsets)
    out = np.empty(shape, dtype=parts[0].dtype)
    for p, off in zip(parts, offsets):
        idx = tuple(slice(o - t, o - t + s) for o, t, s in zip(off, toff, p.shape))
        out[idx] = p
    return out


# --------------------------------------------------------------------------------------------
# casts (integer arithmetic on bit patterns)
# --------------------------------------------------------------------------------------------


def f32_bits_to_bf16_bits(u32: np.ndarray, nan_mode: str = "cpu") -> np.ndarray:
    """torch .to(bfloat16): round to nearest even; NaN -> 0x7FC0 (CPU, c10) or 0x7FFF (CUDA cvt.rn)."""
    u = u32.astype(np.uint64)
    bias = ((u >> 16) & 1) + 0x7FFF
    out = ((u + bias) >> 16).astype(np.uint16)
    nan = (u32 & np.uint32(0x7FFFFFFF)) > np.uint32(0x7F800000)
    out[nan] = 0x7FC0 if nan_mode == "cpu" else 0x7FFF
    return out


# --------------------------------------------------------------------------------------------
# direct weight sync
# --------------------------------------------------------------------------------------------


@dataclass
class PlanOp:
    """One _TransferOp (direct_weight_sync.py:184-206) in value form."""

    name: str
    source_index: int  # position of the handle in all_handles[name]
    source_rank: int
    exact: bool
    src_index: tuple | None  # slices into the FULL source shard (None for exact)
    dest_index: tuple | None  # slices into the dest local tensor (None for exact)
    source_shape: tuple = ()


def build_plan(all_handles: dict, dest: dict) -> list[PlanOp]:
    """all_handles: name -> list of (Slice, source_rank); dest: name -> Slice of the dest shard.

    Restates _build_plan: per dest param, per source handle in order; skip empty intersections;
    dedup on (intersection.offsets, intersection.local_shape); exact iff offsets and local_shape
    of the source equal the dest's.
    """
    ops: list[PlanOp] = []
    for name, dslice in dest.items():
        handles = all_handles.get(name)
        if not handles:
            continue
        seen = set()
        for si, (sslice, srank) in enumerate(handles):
            inter = slice_intersection(sslice, dslice)
            if inter is None:
                continue
            key = (inter.offsets, inter.local_shape)
            if key in seen:
                continue
            seen.add(key)
            exact = sslice.offsets == dslice.offsets and sslice.local_shape == dslice.local_shape
            if exact:
                ops.append(PlanOp(name, si, srank, True, None, None, tuple(sslice.local_shape)))
            else:
                nd = len(inter.offsets)
                src_idx = tuple(
                    slice(inter.offsets[d] - sslice.offsets[d], inter.offsets[d] - sslice.offsets[d] + inter.local_shape[d])
                    for d in range(nd)
                )
                dst_idx = tuple(
                    slice(inter.offsets[d] - dslice.offsets[d], inter.offsets[d] - dslice.offsets[d] + inter.local_shape[d])
                    for d in range(nd)
                )
                ops.append(PlanOp(name, si, srank, False, src_idx, dst_idx, tuple(sslice.local_shape)))
    return ops


def pull(plan: list[PlanOp], sources: dict, dest_arrays: dict) -> None:
    """Execute a plan: sources[name][source_index] is the full source shard (numpy array, already in
    the transfer dtype); dest_arrays[name] is written in place.  Same-itemsize byte semantics:
    the reference reads raw bytes into a buffer of the DEST dtype (direct_weight_sync.py:280-286)."""
    for op in plan:
        src = sources[op.name][op.source_index]
        dst = dest_arrays[op.name]
        if op.exact:
            # byte copy of the whole shard into param memory
            flat_dst = dst.reshape(-1).view(np.uint8)
            flat_dst[...] = np.ascontiguousarray(src).reshape(-1).view(np.uint8)
        else:
            recv = np.ascontiguousarray(src).reshape(-1).view(np.uint8).view(dst.dtype).reshape(op.source_shape)
            dst[op.dest_index] = recv[op.src_index]


def plan_stats(plan: list[PlanOp], itemsize_of: dict) -> dict:
    """Counts used by BASELINE.md: ops, exact ops, algorithmic bytes, reference-direct read bytes."""
    n_exact = sum(1 for op in plan if op.exact)
    algo = 0
    ref_read = 0
    for op in plan:
        isz = itemsize_of[op.name]
        full = math.prod(op.source_shape) * isz
        ref_read += full
        if op.exact:
            algo += full
        else:
            algo += math.prod(s.stop - s.start for s in op.dest_index) * isz
    return {"ops": len(plan), "exact_ops": n_exact, "algorithmic_bytes": algo, "reference_read_bytes": ref_read}


# --------------------------------------------------------------------------------------------
# store path (SharedMemory transport semantics, restated)
# --------------------------------------------------------------------------------------------


@dataclass
class _Stored:
    kind: str  # "tensor" | "object" | "sharded"
    tensor: np.ndarray | None = None
    obj: object = None
    shards: dict = field(default_factory=dict)  # coordinates -> (Slice, ndarray)


class OracleStore:
    """Controller index + per-volume kv + the two copies of the shm path (put: shard -> segment,
    get: segment view -> destination).  One instance models the whole single-host store."""

    def __init__(self, num_volumes: int):
        self.volumes: list[dict[str, _Stored]] = [dict() for _ in range(num_volumes)]
        # key -> {volume_id -> ("tensor"|"object"|"slice", set(Slice|None))}
        self.index: dict[str, dict[int, tuple[str, set]]] = {}
        self.copies = 0  # number of array copies performed (the cost model of the CPU baseline)

    # -- put ---------------------------------------------------------------------------------
    def put(self, volume: int, key: str, value, tslice: Slice | None = None) -> None:
        kv = self.volumes[volume]
        if not isinstance(value, np.ndarray):
            kv[key] = _Stored("object", obj=value)
            kind = "object"
        elif tslice is None:
            cur = kv.get(key)
            if cur is not None and cur.kind == "tensor" and cur.tensor.shape == value.shape and cur.tensor.dtype == value.dtype:
                cur.tensor[...] = value  # in-place overwrite of the existing segment
            else:
                kv[key] = _Stored("tensor", tensor=np.array(value, copy=True))
            self.copies += 1
            kind = "tensor"
        else:
            assert tuple(value.shape) == tuple(tslice.local_shape)
            cur = kv.get(key)
            if cur is None or cur.kind != "sharded":
                cur = kv[key] = _Stored("sharded")
            old = cur.shards.get(tslice.coordinates)
            if old is not None and old[1].shape == value.shape and old[1].dtype == value.dtype:
                old[1][...] = value
                cur.shards[tslice.coordinates] = (tslice, old[1])
            else:
                cur.shards[tslice.coordinates] = (tslice, np.array(value, copy=True))
            self.copies += 1
            kind = "slice"
        ent = self.index.setdefault(key, {})
        if volume in ent:
            assert ent[volume][0] == kind, "storage type of an existing key changed"
            ent[volume][1].add(tslice)
        else:
            ent[volume] = (kind, {tslice})

    # -- locate ------------------------------------------------------------------------------
    def _fully_committed(self, key) -> bool:
        coords, mesh = set(), None
        for kind, slices in self.index[key].values():
            if kind != "slice":
                return True
            for s in slices:
                coords.add(s.coordinates)
                mesh = mesh or s.mesh_shape
        return coords == set(itertools.product(*(range(m) for m in mesh)))

    def locate(self, key):
        if key not in self.index:
            raise KeyError(f"Unable to locate {key} in any storage volumes.")
        if not self._fully_committed(key):
            raise KeyError(f"DTensor '{key}' is only partially committed.")
        return self.index[key]

    # -- get ---------------------------------------------------------------------------------
    def _volume_read(self, volume: int, key: str, want: Slice | None):
        st = self.volumes[volume][key]
        if st.kind == "object":
            return st.obj
        if st.kind == "tensor":
            if want is None:
                return st.tensor
            return st.tensor[tuple(slice(o, o + s) for o, s in zip(want.offsets, want.local_shape))]
        if want is None:
            raise RuntimeError(f"Key '{key}' contains sharded tensor but no tensor_slice was requested")
        for sslice, arr in st.shards.values():
            inter = slice_intersection(sslice, want)
            if inter is None or inter.local_shape != want.local_shape or inter.offsets != want.offsets:
                continue
            idx = tuple(slice(inter.offsets[d] - sslice.offsets[d], inter.offsets[d] - sslice.offsets[d] + inter.local_shape[d])
                        for d in range(len(inter.offsets)))
            return arr[idx]
        raise RuntimeError(f"Tensor slice {want} not found in any stored shards for {key}")

    def get(self, key: str, dest: np.ndarray | None = None, want: Slice | None = None):
        """client.get semantics: returns `dest` (filled in place) when given, else a fresh array
        (or the object)."""
        vmap = self.locate(key)
        parts = []  # (array_or_view_into_dest, Slice)
        whole = None
        for volume, (kind, slices) in vmap.items():
            if kind == "object":
                return self._volume_read(volume, key, None)
            if kind == "tensor":
                data = self._volume_read(volume, key, want)
                if dest is not None:
                    dest[...] = data
                    self.copies += 1
                    return dest
                self.copies += 1
                return np.array(data, copy=True)
            for stored in slices:
                fetch = stored if want is None else slice_intersection(stored, want)
                if fetch is None:
                    continue
                data = self._volume_read(volume, key, fetch)
                view_idx = None
                if dest is not None:
                    view_idx = destination_view(dest.shape, dest.flags["C_CONTIGUOUS"], want, fetch)
                if view_idx is not None:
                    dest[view_idx] = data
                    self.copies += 1
                    parts.append((None, fetch))
                else:
                    self.copies += 1
                    parts.append((np.array(data, copy=True), fetch))
        if not parts:
            raise RuntimeError(f"No results found for key '{key}'.")
        if dest is not None and all(p is None for p, _ in parts):
            return dest
        arrays = []
        for p, f in parts:
            if p is None:  # landed in dest already; re-read it for the assemble fallback
                idx = destination_view(dest.shape, True, want, f)
                p = dest[idx]
            arrays.append(p)
        whole = assemble(arrays, [f.offsets for _, f in parts])
        self.copies += 1
        if dest is not None:
            dest[...] = whole
            self.copies += 1
            return dest
        return whole
