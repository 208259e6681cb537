"""GPU parity of the copy_rects kernel through the C-ABI: identical rectangle descriptors are run by
the CUDA kernel (on device copies) and by the C oracle (on host copies) and compared byte for byte,
including the bytes OUTSIDE every destination window."""

import os
import random

import numpy as np
import pytest
import torch

from oracle import c_oracle
from tests.helpers import (CAST_PAIRS, assert_equal_modulo_nan, bytes_of, materialise, random_case, rects_for)
from torchstore_b200 import _native
from torchstore_b200.planner import StridedMem, build_rects

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _init():
    _native.init()
    yield


def run_product(pairs, device=0, flags=0):
    rects, n = rects_for(pairs)
    plan = _native.plan_create(device, rects, n, flags)
    info = _native.plan_info(plan)
    before = _native.launch_count()
    _native.plan_run(plan, None)
    _native.stream_sync(device, None)
    assert _native.launch_count() == before + (1 if info.num_tiles else 0)
    _native.plan_destroy(plan)
    return info


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("cast", [False, True])
@pytest.mark.parametrize("force_generic", [False, True])
def test_random_rects_match_oracle(seed, cast, force_generic, monkeypatch):
    if force_generic:
        monkeypatch.setenv("TSB_FORCE_GENERIC", "1")
    else:
        monkeypatch.delenv("TSB_FORCE_GENERIC", raising=False)
    rng = random.Random(1000 * seed + cast)
    cases = [random_case(rng, cast=cast, max_elems=1 << 18) for _ in range(80)]
    host = materialise(cases, "cpu", seed)
    dev = materialise(cases, "cuda:0", seed)
    rects, n = rects_for(host)
    c_oracle.copy_rects(rects, n, nan_mode=1)
    run_product(dev)
    for (hs, hd, hbase, _), (ds, dd, dbase, _), case in zip(host, dev, cases):
        # window and everything around it
        assert_equal_modulo_nan(dbase, hbase, cast=case.src_dtype != case.dst_dtype)


@pytest.mark.parametrize("src_dtype,dst_dtype", CAST_PAIRS)
def test_cast_matches_torch_cuda_bit_exact(src_dtype, dst_dtype):
    """The cast path must equal what the reference computes on a GPU-resident param:
    tensor.to(transfer_dtype) on CUDA (direct_weight_sync.py:133), NaNs included."""
    gen = torch.Generator().manual_seed(3)
    n = 1 << 20
    raw = torch.randint(0, 256, (n * src_dtype.itemsize,), dtype=torch.uint8, generator=gen).view(src_dtype).cuda()
    # add edge values
    edge = torch.tensor([0.0, -0.0, 1.0, float("inf"), float("-inf"), float("nan"), 65504.0, 65520.0, 1e-8, 3.3895e38],
                        dtype=torch.float64).to(src_dtype).cuda()
    src = torch.cat([edge, raw])[: n - (n % 8)]
    want = src.to(dst_dtype)
    got = torch.zeros_like(want)
    run_product([(src, got)])
    if dst_dtype == torch.float32 and src_dtype == torch.float64:
        assert_equal_modulo_nan(got, want, cast=True)  # x86/CUDA differ in NaN payload propagation only
    else:
        assert torch.equal(got.view(torch.uint8), want.view(torch.uint8))
    # scalar path (misaligned views)
    got2 = torch.zeros(want.numel() + 1, dtype=dst_dtype, device="cuda")[1:]
    run_product([(src, got2)])
    assert_equal_modulo_nan(got2, want, cast=True)


def test_large_contiguous_int32_is_bit_exact():
    """Config #2 shape in miniature (1 GiB): integer/index tensors must be bit-exact."""
    n = 1 << 28
    src = torch.arange(n, dtype=torch.int32, device="cuda")
    dst = torch.zeros_like(src)
    info = run_product([(src, dst)])
    assert info.payload_bytes == n * 4 and info.num_vector_rects == 1
    assert torch.equal(src, dst)
    # properties at full size: sum of a checksum of checksums
    assert int(dst.view(torch.int64).sum().item()) == int(src.view(torch.int64).sum().item())


def test_narrow_row_reshard_shapes():
    """wo / w2 style rectangles: 512 x 512 and 512 x 1792 bf16 windows with 8 KiB / 28 KiB source
    pitch, written into a [4096, 512] / [4096, 1792] destination."""
    for cols_total, cols in ((4096, 512), (14336, 1792)):
        srcs = [torch.randn(512, cols_total, device="cuda").to(torch.bfloat16) for _ in range(8)]
        for r in (0, 3, 7):
            dst = torch.zeros(4096, cols, dtype=torch.bfloat16, device="cuda")
            pairs = [(srcs[s][:, r * cols:(r + 1) * cols], dst[s * 512:(s + 1) * 512]) for s in range(8)]
            info = run_product(pairs)
            assert info.num_vector_rects == 8
            want = torch.cat([s[:, r * cols:(r + 1) * cols] for s in srcs], dim=0)
            assert torch.equal(dst, want)


def test_many_tiny_and_one_huge_in_one_launch():
    tiny_src = [torch.randn(512, device="cuda").to(torch.bfloat16) for _ in range(300)]
    tiny_dst = [torch.zeros(512, dtype=torch.bfloat16, device="cuda") for _ in range(300)]
    big_src = torch.randn(64 << 20, device="cuda").to(torch.bfloat16)
    big_dst = torch.zeros_like(big_src)
    info = run_product(list(zip(tiny_src, tiny_dst)) + [(big_src, big_dst)])
    assert info.num_rects == 301
    assert torch.equal(big_src, big_dst)
    assert all(torch.equal(a, b) for a, b in zip(tiny_src, tiny_dst))


def test_one_shot_copy_rects_and_events():
    a = torch.arange(1 << 20, dtype=torch.int64, device="cuda")
    b = torch.zeros_like(a)
    rects, n = build_rects([(StridedMem.from_tensor(a), StridedMem.from_tensor(b))])
    torch.cuda.synchronize()
    start = _native.Event(0, timing=True).record(None)
    _native.copy_rects(0, rects, n)
    done = _native.Event(0, timing=True).record(None)
    done.synchronize()
    assert done.query()
    assert start.elapsed_ms(done) > 0
    assert torch.equal(a, b)


def test_caller_stream_and_torch_interop():
    """Launching on torch's current stream orders with torch work on that stream."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        a = torch.full((1 << 22,), 7, dtype=torch.int32, device="cuda")
        b = torch.zeros_like(a)
        rects, n = build_rects([(StridedMem.from_tensor(a), StridedMem.from_tensor(b))])
        plan = _native.plan_create(0, rects, n)
        _native.plan_run(plan, s.cuda_stream)
        c = b + 1
    s.synchronize()
    assert int(c.sum().item()) == 8 * (1 << 22)
    _native.plan_destroy(plan)


def test_arena_alloc_free_ring():
    arena = _native.arena_create(0, 64 << 20)
    st = _native.arena_stats(arena)
    assert st.capacity == 64 << 20 and st.in_use == 0
    ptrs = [_native.arena_alloc(arena, 1 << 20) for _ in range(32)]
    assert len(set(ptrs)) == 32 and all(p % 256 == 0 for p in ptrs)
    assert _native.arena_stats(arena).in_use == 32 << 20
    for p in ptrs[::2]:
        _native.arena_free(arena, p)
    big = _native.arena_alloc(arena, 30 << 20)  # fits only in the untouched tail
    with pytest.raises(_native.TsbError):
        _native.arena_alloc(arena, 40 << 20)
    _native.arena_free(arena, big)
    for p in ptrs[1::2]:
        _native.arena_free(arena, p)
    assert _native.arena_stats(arena).in_use == 0
    whole = _native.arena_alloc(arena, 64 << 20)  # everything coalesced back
    _native.arena_free(arena, whole)
    # arena memory is exportable and usable by the kernel
    p = _native.arena_alloc(arena, 1 << 20)
    reg = _native.export_region(p, 1 << 20)
    assert reg.nbytes == 1 << 20 and _native.import_region(reg, 0) == p
    _native.arena_destroy(arena)


def test_export_import_same_process_torch_suballocation():
    """A tensor carved by torch's caching allocator exports as (allocation handle, offset)."""
    pad = torch.empty(1000, device="cuda")  # noqa: F841 -- make the next tensor a sub-allocation
    t = torch.arange(4096, dtype=torch.float32, device="cuda")
    reg = _native.export_region(t.data_ptr(), t.numel() * 4)
    assert reg.nbytes == 4096 * 4 and reg.alloc_bytes >= reg.offset + reg.nbytes
    assert _native.import_region(reg, 0) == t.data_ptr()
    blob = _native.region_to_bytes(reg)
    assert len(blob) == 112 and _native.region_from_bytes(blob).offset == reg.offset
