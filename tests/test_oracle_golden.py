"""Pin the CPU oracle (oracle/) to the real reference: every check here compares the oracle with
fixtures that oracle/gen_golden.py recorded from the UNMODIFIED reference run in-process.
CPU only."""

import hashlib
import itertools
import json
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle import reshard_oracle as ro
from torchstore_b200 import _native

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def S(j):
    if j is None:
        return None
    return ro.Slice(tuple(j["offsets"]), None if j["coordinates"] is None else tuple(j["coordinates"]),
                    tuple(j["global_shape"]), tuple(j["local_shape"]),
                    None if j["mesh_shape"] is None else tuple(j["mesh_shape"]))


def test_shard_layout_matches_torch():
    for c in load("slice_math.json")["layouts"]:
        pl = [tuple(p) for p in c["placements"]]
        shp, off = ro.shard_layout(tuple(c["global_shape"]), tuple(c["mesh_shape"]), tuple(c["coordinate"]), pl)
        assert list(shp) == c["local_shape"] and list(off) == c["offsets"], c


def test_slice_intersection_matches_reference():
    for c in load("slice_math.json")["intersections"]:
        got = ro.slice_intersection(S(c["stored"]), S(c["wanted"]))
        assert got == S(c["result"]), c


def test_destination_view_matches_reference():
    for c in load("slice_math.json")["dest_views"]:
        got = ro.destination_view(c["dest_shape"], c["dest_contiguous"], S(c["dest_slice"]), S(c["fetch"]))
        want = c["result"]
        if want is None:
            assert got is None, c
        else:
            assert got is not None, c
            assert [[s.start, s.stop] for s in got] == want, c


def test_assemble_matches_reference():
    d = load("slice_math.json")
    for c in d["assemble"]:
        parts = [np.array(p) for p in c["parts"]]
        got = ro.assemble(parts, [tuple(o) for o in c["offsets"]])
        assert got.tolist() == c["result"], c
    for c in d["get_local_tensor"]:
        g = np.array(c["global"])
        idx = tuple(slice(o, o + s) for o, s in zip(c["offset"], c["shape"]))
        assert g[idx].tolist() == c["result"]


def _np_dtype(name):
    return {"float32": np.float32, "int32": np.int32}[name]


def _sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _case_inputs(params):
    """Rebuild sources / dest slices of one direct_plan case from its layout description."""
    all_handles, dest_slices, sources, dest_arrays = {}, {}, {}, {}
    for name, p in params.items():
        shape = tuple(p["global_shape"])
        full = (np.arange(int(np.prod(shape))) + p["arange_start"]).astype(_np_dtype(p["dtype"])).reshape(shape)
        smesh = tuple(p["src_mesh"])
        spl = [tuple(x) for x in p["src_placements"]]
        hl, src = [], []
        for rank, coord in enumerate(itertools.product(*(range(m) for m in smesh))):
            sl = ro.make_slice(shape, smesh, coord, spl)
            hl.append((sl, rank))
            src.append(np.ascontiguousarray(full[tuple(slice(o, o + s) for o, s in zip(sl.offsets, sl.local_shape))]))
        all_handles[name] = hl
        sources[name] = src
        if p["dst_mesh"] is None:
            dsl = ro.full_slice(shape)
        else:
            dmesh = tuple(p["dst_mesh"])
            coord = list(itertools.product(*(range(m) for m in dmesh)))[p["dst_rank"]]
            dsl = ro.make_slice(shape, dmesh, coord, [tuple(x) for x in p["dst_placements"]])
        dest_slices[name] = dsl
        dest_arrays[name] = np.zeros(dsl.local_shape, dtype=full.dtype)
    return all_handles, dest_slices, sources, dest_arrays


def test_direct_plan_and_pull_match_reference():
    d = load("direct_plan.json")
    assert len(d["cases"]) >= 30
    for case in d["cases"]:
        all_handles, dest_slices, sources, dest_arrays = _case_inputs(case["params"])
        plan = ro.build_plan(all_handles, dest_slices)
        got = [
            {
                "name": op.name,
                "source_index": op.source_index,
                "source_rank": op.source_rank,
                "exact": op.exact,
                "src_index": None if op.src_index is None else [[s.start, s.stop] for s in op.src_index],
                "dest_index": None if op.dest_index is None else [[s.start, s.stop] for s in op.dest_index],
            }
            for op in plan
        ]
        assert got == case["ops"], case["label"]
        ro.pull(plan, sources, dest_arrays)
        for name, digest in case["dest_sha256"].items():
            assert _sha(dest_arrays[name]) == digest, (case["label"], name)


LLAMA3_8B = dict(dim=4096, n_layers=32, n_heads=32, n_kv_heads=8, vocab=128256, ffn=14336)


def llama_layout(cfg=LLAMA3_8B):
    d, kv = cfg["dim"], cfg["dim"] // cfg["n_heads"] * cfg["n_kv_heads"]
    out = {"tok_embeddings.weight": ((cfg["vocab"], d), ("S", 0))}
    for i in range(cfg["n_layers"]):
        p = f"layers.{i}."
        out[p + "attention.wq.weight"] = ((d, d), ("S", 0))
        out[p + "attention.wk.weight"] = ((kv, d), ("S", 0))
        out[p + "attention.wv.weight"] = ((kv, d), ("S", 0))
        out[p + "attention.wo.weight"] = ((d, d), ("S", 1))
        out[p + "feed_forward.w1.weight"] = ((cfg["ffn"], d), ("S", 0))
        out[p + "feed_forward.w2.weight"] = ((d, cfg["ffn"]), ("S", 1))
        out[p + "feed_forward.w3.weight"] = ((cfg["ffn"], d), ("S", 0))
        out[p + "attention_norm.weight"] = ((d,), ("R",))
        out[p + "ffn_norm.weight"] = ((d,), ("R",))
    out["norm.weight"] = ((d,), ("R",))
    out["output.weight"] = ((cfg["vocab"], d), ("S", 0))
    return out


def test_llama3_8b_plan_statistics_match_reference():
    """Op counts / algorithmic bytes / reference read amplification of FSDP(N)->TP(N)."""
    layout = llama_layout()
    assert len(layout) == 291
    for entry in load("direct_plan.json")["llama3_8b_fsdp_to_tp"]:
        n = entry["n"]
        for drank, want in enumerate(entry["per_dest_rank"]):
            handles, dest = {}, {}
            for name, (shape, tp) in layout.items():
                handles[name] = [(ro.make_slice(shape, (n,), (r,), [("S", 0)]), r) for r in range(n)]
                dest[name] = ro.full_slice(shape) if n == 1 else ro.make_slice(shape, (n,), (drank,), [tp])
            stats = ro.plan_stats(ro.build_plan(handles, dest), {k: 2 for k in layout})
            assert stats == want, (n, drank)


def test_store_reshard_matches_reference():
    d = load("store_reshard.json")
    for case in d["cases"]:
        shape = tuple(case["global_shape"])
        full = np.arange(int(np.prod(shape)), dtype=np.float32).reshape(shape)
        smesh, dmesh = tuple(case["src_mesh"]), tuple(case["dst_mesh"])
        spl = [tuple(p) for p in case["src_placements"]]
        dpl = [tuple(p) for p in case["dst_placements"]]
        nvol = max(int(np.prod(smesh)), int(np.prod(dmesh)))
        store = ro.OracleStore(nvol)
        all_rep = all(p[0] == "R" for p in spl)
        for rank, coord in enumerate(itertools.product(*(range(m) for m in smesh))):
            sl = ro.make_slice(shape, smesh, coord, spl)
            local = np.ascontiguousarray(full[tuple(slice(o, o + s) for o, s in zip(sl.offsets, sl.local_shape))])
            store.put(rank, "test_key", local, None if all_rep else sl)
        for r in case["per_rank"]:
            coord = list(itertools.product(*(range(m) for m in dmesh)))[r["rank"]]
            want = ro.make_slice(shape, dmesh, coord, dpl)
            assert list(want.local_shape) == r["local_shape"] and list(want.offsets) == r["offsets"]
            dest = np.zeros(want.local_shape, dtype=np.float32)
            got = store.get("test_key", dest, want)
            assert got is dest
            assert _sha(dest) == r["sha256"], (case, r["rank"])
        assert _sha(store.get("test_key")) == case["full_get_sha256"]
    # explicit slice get + partial commit
    ts = d["tensor_slice_get"]
    store = ro.OracleStore(1)
    t = np.arange(100 * 100, dtype=np.float32).reshape(100, 100)
    store.put(0, "t", t)
    spec = ro.Slice(tuple(ts["offsets"]), (), tuple(ts["global_shape"]), tuple(ts["local_shape"]), ())
    assert _sha(store.get("t", want=spec)) == ts["sha256"]
    store2 = ro.OracleStore(2)
    small = np.arange(48, dtype=np.float32).reshape(8, 6)
    store2.put(0, "p", small[:4].copy(), ro.Slice((0, 0), (0,), (8, 6), (4, 6), (2,)))
    with pytest.raises(KeyError, match=d["partial_commit_error_contains"]):
        store2.get("p")


# ---- casts: numpy oracle and C oracle against torch-CPU bit patterns -------------------------------
def _cast_vectors():
    return np.load(os.path.join(GOLDEN, "cast_vectors.npz"))


def _nan16(bits, exp_mask, man_mask):
    return ((bits & exp_mask) == exp_mask) & ((bits & man_mask) != 0)


NAN_MASKS = {
    "bf16": lambda b: _nan16(b, np.uint16(0x7F80), np.uint16(0x007F)),
    "f16": lambda b: _nan16(b, np.uint16(0x7C00), np.uint16(0x03FF)),
    "f32": lambda b: (b & np.uint32(0x7FFFFFFF)) > np.uint32(0x7F800000),
    "f64": lambda b: (b & np.uint64(0x7FFFFFFFFFFFFFFF)) > np.uint64(0x7FF0000000000000),
}


def assert_bits_equal_modulo_nan(got, want, kind):
    """Bit-exact everywhere except that any NaN encoding matches any other: torch itself produces
    0x7FC0 (c10 scalar), 0xFFFF (ATen AVX512 path) or 0x7FFF (CUDA cvt.rn) for the same bf16 NaN."""
    gn, wn = NAN_MASKS[kind](got), NAN_MASKS[kind](want)
    assert np.array_equal(gn, wn)
    assert np.array_equal(got[~wn], want[~wn])


def test_numpy_cast_matches_torch_cpu():
    v = _cast_vectors()
    assert_bits_equal_modulo_nan(ro.f32_bits_to_bf16_bits(v["f32"]), v["f32_to_bf16"], "bf16")


@pytest.mark.parametrize(
    "src_key,src_code,dst_key,dst_code,dst_np,kind",
    [
        ("f32", _native.TSB_F32, "f32_to_bf16", _native.TSB_BF16, np.uint16, "bf16"),
        ("f32", _native.TSB_F32, "f32_to_f16", _native.TSB_F16, np.uint16, "f16"),
        ("all16", _native.TSB_BF16, "bf16_to_f32", _native.TSB_F32, np.uint32, "f32"),
        ("all16", _native.TSB_F16, "f16_to_f32", _native.TSB_F32, np.uint32, "f32"),
        ("all16", _native.TSB_BF16, "bf16_to_f16", _native.TSB_F16, np.uint16, "f16"),
        ("all16", _native.TSB_F16, "f16_to_bf16", _native.TSB_BF16, np.uint16, "bf16"),
        ("f64", _native.TSB_F64, "f64_to_f32", _native.TSB_F32, np.uint32, "f32"),
        ("f32", _native.TSB_F32, "f32_to_f64", _native.TSB_F64, np.uint64, "f64"),
    ],
)
def test_c_oracle_cast_matches_torch_cpu(src_key, src_code, dst_key, dst_code, dst_np, kind):
    v = _cast_vectors()
    src = np.ascontiguousarray(v[src_key])
    want = v[dst_key]
    for nan_mode in (0, 1):
        got = np.zeros(src.shape[0], dtype=dst_np)
        c_oracle.convert(src.ctypes.data, src_code, got.ctypes.data, dst_code, src.shape[0], nan_mode=nan_mode)
        assert_bits_equal_modulo_nan(got, want, kind)
