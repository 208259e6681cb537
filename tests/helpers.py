"""Shared helpers for the test-suite: seeded random rectangle workloads that can be materialised on
the CPU (for the C oracle) and on CUDA (for the product), plus Llama-3 layouts."""

from __future__ import annotations

import random

import numpy as np
import torch

from torchstore_b200 import _native
from torchstore_b200.planner import StridedMem, build_rects

SAME_DTYPES = [torch.uint8, torch.int16, torch.bfloat16, torch.float16, torch.int32, torch.float32, torch.int64, torch.float64]
CAST_PAIRS = [
    (torch.float32, torch.bfloat16),
    (torch.float32, torch.float16),
    (torch.bfloat16, torch.float32),
    (torch.float16, torch.float32),
    (torch.bfloat16, torch.float16),
    (torch.float16, torch.bfloat16),
    (torch.float64, torch.float32),
    (torch.float32, torch.float64),
]


def random_bits(shape, dtype: torch.dtype, gen: torch.Generator, finite_only: bool = False) -> torch.Tensor:
    """Tensor of `dtype` filled with random bit patterns (so NaN payloads, denormals, etc. occur)."""
    n = int(np.prod(shape)) if len(shape) else 1
    raw = torch.randint(0, 256, (n * dtype.itemsize,), dtype=torch.uint8, generator=gen)
    t = raw.view(dtype).reshape(shape)
    if finite_only and dtype.is_floating_point:
        t = torch.where(torch.isfinite(t), t, torch.zeros((), dtype=dtype))
    return t.clone()


class RectCase:
    """One (src window, dst window) pair described relative to two base buffers, so the same case
    can be instantiated on any device."""

    def __init__(self, src_base_shape, dst_base_shape, src_dtype, dst_dtype, src_view, dst_view):
        self.src_base_shape = tuple(src_base_shape)
        self.dst_base_shape = tuple(dst_base_shape)
        self.src_dtype = src_dtype
        self.dst_dtype = dst_dtype
        self.src_view = src_view  # callable(base) -> view
        self.dst_view = dst_view


def _rand_window(rng: random.Random, shape, base_pad):
    """Pick a base shape >= shape and a slice window of `shape` inside it."""
    base = [s + rng.choice([0, 0, 1, 3, 8]) if base_pad else s for s in shape]
    starts = [rng.randint(0, b - s) for b, s in zip(base, shape)]
    idx = tuple(slice(a, a + s) for a, s in zip(starts, shape))
    return tuple(base), idx


def random_case(rng: random.Random, cast: bool = False, max_elems: int = 1 << 16) -> RectCase:
    if cast:
        sdt, ddt = rng.choice(CAST_PAIRS)
    else:
        sdt = ddt = rng.choice(SAME_DTYPES)
    ndim = rng.choice([1, 1, 2, 2, 2, 3, 4])
    while True:
        shape = [rng.choice([1, 2, 3, 5, 7, 8, 16, 17, 31, 32, 64, 100, 128, 257, 512, 1000]) for _ in range(ndim)]
        if int(np.prod(shape)) <= max_elems:
            break
    sbase, sidx = _rand_window(rng, shape, rng.random() < 0.7)
    dbase, didx = _rand_window(rng, shape, rng.random() < 0.7)
    s_perm = list(range(ndim))
    d_perm = list(range(ndim))
    if ndim >= 2 and rng.random() < 0.2:
        rng.shuffle(s_perm)
    if ndim >= 2 and rng.random() < 0.15:
        rng.shuffle(d_perm)
    # views: base is allocated in permuted order, then permuted back, then windowed
    def mk(perm, base, idx):
        inv = [perm.index(i) for i in range(len(perm))]
        alloc_shape = tuple(base[p] for p in perm)

        def view(buf):
            return buf.reshape(alloc_shape).permute(*inv)[idx]

        return alloc_shape, view

    s_alloc, s_view = mk(s_perm, sbase, sidx)
    d_alloc, d_view = mk(d_perm, dbase, didx)
    return RectCase(s_alloc, d_alloc, sdt, ddt, s_view, d_view)


def materialise(cases, device, seed: int, finite_only: bool = False):
    """Allocate bases on `device` (filled from a seeded CPU generator so CPU and CUDA copies hold
    identical bytes) and return [(src_view, dst_view, dst_base)]."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for c in cases:
        # +2 elements of slack and an optional odd offset exercise unaligned bases
        src_base = random_bits(c.src_base_shape, c.src_dtype, gen, finite_only).to(device)
        dst_base = random_bits(c.dst_base_shape, c.dst_dtype, gen, True).to(device)
        out.append((c.src_view(src_base), c.dst_view(dst_base), dst_base, src_base))
    return out


def rects_for(pairs):
    """[(src_view, dst_view, ...)] -> (ctypes rect array, n)."""
    return build_rects([(StridedMem.from_tensor(p[0]), StridedMem.from_tensor(p[1])) for p in pairs])


def bytes_of(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().contiguous().view(torch.uint8).numpy().copy()


NAN_CHECK = {
    torch.bfloat16: (np.uint16, 0x7F80, 0x007F),
    torch.float16: (np.uint16, 0x7C00, 0x03FF),
    torch.float32: (np.uint32, 0x7F800000, 0x007FFFFF),
    torch.float64: (np.uint64, 0x7FF0000000000000, 0x000FFFFFFFFFFFFF),
}


def assert_equal_modulo_nan(got: torch.Tensor, want: torch.Tensor, cast: bool):
    """Bit-exact; for cast outputs any NaN encoding equals any other NaN."""
    g, w = bytes_of(got), bytes_of(want)
    if not cast or got.dtype not in NAN_CHECK:
        assert np.array_equal(g, w)
        return
    np_t, emask, mmask = NAN_CHECK[got.dtype]
    gb, wb = g.view(np_t), w.view(np_t)
    gn = ((gb & np_t(emask)) == np_t(emask)) & ((gb & np_t(mmask)) != 0)
    wn = ((wb & np_t(emask)) == np_t(emask)) & ((wb & np_t(mmask)) != 0)
    assert np.array_equal(gn, wn)
    assert np.array_equal(gb[~wn], wb[~wn])


# ---- Llama-3 layouts (torchtitan TP plan; SURVEY.md section 8 header) -------------------------------
LLAMA3_8B = dict(dim=4096, n_layers=32, n_heads=32, n_kv_heads=8, vocab=128256, ffn=14336)
LLAMA3_70B = dict(dim=8192, n_layers=80, n_heads=64, n_kv_heads=8, vocab=128256, ffn=28672)


def llama_layout(cfg=LLAMA3_8B, n_layers: int | None = None, scale: int = 1):
    """name -> (global shape, TP placement).  `scale` divides every dimension (small test models)."""
    d = cfg["dim"] // scale
    kv = cfg["dim"] // cfg["n_heads"] * cfg["n_kv_heads"] // scale
    ffn = cfg["ffn"] // scale
    vocab = cfg["vocab"] // scale
    layers = cfg["n_layers"] if n_layers is None else n_layers
    out = {"tok_embeddings.weight": ((vocab, d), ("S", 0))}
    for i in range(layers):
        p = f"layers.{i}."
        out[p + "attention.wq.weight"] = ((d, d), ("S", 0))
        out[p + "attention.wk.weight"] = ((kv, d), ("S", 0))
        out[p + "attention.wv.weight"] = ((kv, d), ("S", 0))
        out[p + "attention.wo.weight"] = ((d, d), ("S", 1))
        out[p + "feed_forward.w1.weight"] = ((ffn, d), ("S", 0))
        out[p + "feed_forward.w2.weight"] = ((d, ffn), ("S", 1))
        out[p + "feed_forward.w3.weight"] = ((ffn, d), ("S", 0))
        out[p + "attention_norm.weight"] = ((d,), ("R",))
        out[p + "ffn_norm.weight"] = ((d,), ("R",))
    out["norm.weight"] = ((d,), ("R",))
    out["output.weight"] = ((vocab, d), ("S", 0))
    return out
