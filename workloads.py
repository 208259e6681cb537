"""Synthetic weight-sync workloads shared by bench.py, tools/ and the tests: Llama-3 state-dict
layouts (shapes + the torchtitan TP plan of SURVEY.md section 8) and their FSDP/TP rectangles.
Metadata only -- no tensors are created here."""

from __future__ import annotations

import math

LLAMA3_8B = dict(name="llama3-8b", dim=4096, n_layers=32, n_heads=32, n_kv_heads=8, vocab=128256, ffn=14336)
LLAMA3_70B = dict(name="llama3-70b", dim=8192, n_layers=80, n_heads=64, n_kv_heads=8, vocab=128256, ffn=28672)


def llama_layout(cfg=LLAMA3_8B, n_layers: int | None = None, scale: int = 1, with_embeddings: bool = True):
    """name -> (global shape, TP placement) with placement ("S", dim) or ("R",).

    ColwiseParallel (Shard(0)): wq wk wv w1 w3, output; RowwiseParallel embedding: Shard(0);
    RowwiseParallel linear (Shard(1)): wo w2; norms replicated.  `scale` divides every dimension."""
    d = cfg["dim"] // scale
    kv = cfg["dim"] // cfg["n_heads"] * cfg["n_kv_heads"] // scale
    ffn = cfg["ffn"] // scale
    vocab = cfg["vocab"] // scale
    layers = cfg["n_layers"] if n_layers is None else n_layers
    out = {}
    if with_embeddings:
        out["tok_embeddings.weight"] = ((vocab, d), ("S", 0))
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
    if with_embeddings:
        out["norm.weight"] = ((d,), ("R",))
        out["output.weight"] = ((vocab, d), ("S", 0))
    return out


def shard_box(shape, n: int, rank: int, placement):
    """(offsets, local_shape) of `rank` on a 1-D mesh of size n; even or torch.chunk-style uneven."""
    offsets = [0] * len(shape)
    local = list(shape)
    if placement[0] == "S":
        dim = placement[1]
        size = shape[dim]
        chunk = -(-size // n)
        start = min(size, chunk * rank)
        stop = min(size, chunk * (rank + 1))
        offsets[dim] = start if stop > start else size
        local[dim] = max(0, stop - start)
    return tuple(offsets), tuple(local)


def state_dict_bytes(layout, itemsize: int = 2) -> int:
    return sum(math.prod(shape) for shape, _ in layout.values()) * itemsize


def fsdp_to_tp_rects(layout, n: int, dest_rank: int, itemsize: int = 2):
    """Rectangles dest_rank pulls for FSDP(n) Shard(0) -> TP(n): list of
    (name, src_rank, src_index, dst_index, exact) with index = tuple of (start, stop) per dim,
    in the reference's plan order (dest param major, source rank minor, replicated dedup)."""
    out = []
    for name, (shape, tp) in layout.items():
        d_off, d_shape = shard_box(shape, n, dest_rank, tp) if n > 1 else ((0,) * len(shape), tuple(shape))
        seen = set()
        for s in range(n):
            s_off, s_shape = shard_box(shape, n, s, ("S", 0))
            lo = [max(a, b) for a, b in zip(s_off, d_off)]
            hi = [min(a + x, b + y) for a, x, b, y in zip(s_off, s_shape, d_off, d_shape)]
            if any(h <= l for l, h in zip(lo, hi)):
                continue
            key = (tuple(lo), tuple(h - l for l, h in zip(lo, hi)))
            if key in seen:
                continue
            seen.add(key)
            exact = tuple(s_off) == tuple(d_off) and tuple(s_shape) == tuple(d_shape)
            src_idx = tuple((l - o, h - o) for l, h, o in zip(lo, hi, s_off))
            dst_idx = tuple((l - o, h - o) for l, h, o in zip(lo, hi, d_off))
            out.append((name, s, src_idx, dst_idx, exact))
    return out
