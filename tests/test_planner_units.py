"""StridedMem / rect-descriptor unit tests (host logic, CPU)."""

import ctypes as C

import pytest
import torch

from torchstore_b200 import _native
from torchstore_b200.planner import StridedMem, _collapse, _span_elems, build_rects, fill_rect


def test_sub_is_pointer_arithmetic():
    t = torch.arange(6 * 8 * 4, dtype=torch.float32).reshape(6, 8, 4)
    m = StridedMem.from_tensor(t)
    w = m.sub((slice(2, 5), slice(1, 7), slice(0, 4)))
    v = t[2:5, 1:7, 0:4]
    assert w.ptr == v.data_ptr() and w.shape == tuple(v.shape) and w.stride == tuple(v.stride())
    assert m.is_contiguous() and not w.is_contiguous()
    assert w.sub((slice(0, 1), slice(0, 6), slice(0, 4))).is_contiguous()
    with pytest.raises(ValueError):
        m.sub((slice(0, 6, 2), slice(0, 8), slice(0, 4)))
    with pytest.raises(ValueError):
        m.sub((slice(0, 6),))
    with pytest.raises(ValueError):
        w.flat_bytes()
    fb = m.flat_bytes()
    assert fb.shape == (6 * 8 * 4 * 4,) and fb.dtype == torch.uint8 and fb.ptr == t.data_ptr()


def test_collapse_merges_only_jointly_contiguous_dims():
    # [4, 8] window of a [4, 8] tensor on both sides: one dim
    assert _collapse((4, 8), (32, 4), (32, 4)) == [(32, 4, 4)]
    # source pitch 64 B, dest pitch 32 B: cannot merge
    assert _collapse((4, 8), (64, 4), (32, 4)) == [(4, 64, 32), (8, 4, 4)]
    # unit dims vanish
    assert _collapse((1, 4, 1, 8), (999, 32, 999, 4), (7, 32, 7, 4)) == [(32, 4, 4)]
    assert _collapse((1, 1), (4, 4), (4, 4)) == []


def test_fill_rect_layout_and_dtype_codes():
    src = torch.zeros(16, 32, dtype=torch.float32)
    dst = torch.zeros(16, 8, dtype=torch.bfloat16)
    r = _native.Rect()
    assert fill_rect(r, StridedMem.from_tensor(src[:, 8:16]), StridedMem.from_tensor(dst))
    assert r.ndim == 2 and list(r.extent[:2]) == [16, 8]
    assert list(r.src_stride[:2]) == [128, 4] and list(r.dst_stride[:2]) == [16, 2]
    assert r.src_dtype == _native.TSB_F32 and r.dst_dtype == _native.TSB_BF16
    assert r.src == src[:, 8:16].data_ptr() and r.dst == dst.data_ptr()
    # shape mismatch / unsupported cast / too many dims
    with pytest.raises(ValueError):
        fill_rect(r, StridedMem.from_tensor(src), StridedMem.from_tensor(dst))
    with pytest.raises(NotImplementedError):
        fill_rect(r, StridedMem.from_tensor(torch.zeros(4, dtype=torch.int32)), StridedMem.from_tensor(torch.zeros(4)))
    big = torch.zeros((3,) * 7)
    win = big[(slice(0, 2),) * 7]  # 7 dims, none mergeable
    with pytest.raises(NotImplementedError):
        fill_rect(r, StridedMem.from_tensor(win), StridedMem.from_tensor(torch.zeros((3,) * 7)[(slice(1, 3),) * 7]))
    # bool / int8 map to opaque byte codes
    rb = _native.Rect()
    assert fill_rect(rb, StridedMem.from_tensor(torch.zeros(5, dtype=torch.bool)), StridedMem.from_tensor(torch.zeros(5, dtype=torch.bool)))
    assert rb.src_dtype == rb.dst_dtype == _native.TSB_U8


def test_build_rects_skips_empty_and_counts():
    a, b = torch.zeros(4, 4), torch.zeros(4, 4)
    rects, n = build_rects([(StridedMem.from_tensor(a[:0]), StridedMem.from_tensor(b[:0])),
                            (StridedMem.from_tensor(a), StridedMem.from_tensor(b))])
    assert n == 1 and rects[0].extent[0] == 16  # fully collapsed


def test_span_and_region_blob_roundtrip():
    assert _span_elems((4, 8), (8, 1)) == 32
    assert _span_elems((4, 4), (16, 1)) == 52  # strided window
    assert _span_elems((0, 4), (4, 1)) == 0
    reg = _native.Region()
    reg.offset, reg.nbytes, reg.device, reg.pid, reg.boot_id = 4096, 1234, 3, 77, 0xDEADBEEF
    back = _native.region_from_bytes(_native.region_to_bytes(reg))
    assert (back.offset, back.nbytes, back.device, back.pid, back.boot_id) == (4096, 1234, 3, 77, 0xDEADBEEF)
    assert C.sizeof(back) == 120
    with pytest.raises(ValueError):
        _native.region_from_bytes(b"short")


def test_dtype_code_table():
    assert _native.dtype_code(torch.bfloat16) == _native.TSB_BF16
    assert _native.dtype_code(torch.int8) == _native.TSB_U8
    assert _native.dtype_code(torch.int16) == _native.TSB_U16
    assert _native.dtype_code(torch.int32) == _native.TSB_U32
    assert _native.dtype_code(torch.int64) == _native.TSB_U64
    assert _native.dtype_code(torch.complex64) == _native.TSB_U64
    assert _native.cast_supported(torch.float32, torch.bfloat16)
    assert _native.cast_supported(torch.float64, torch.float32)
    assert not _native.cast_supported(torch.int32, torch.float32)
    assert not _native.cast_supported(torch.float64, torch.bfloat16)
