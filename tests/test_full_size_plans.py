"""Full-size properties of the product's planner on the BASELINE workloads, checked on the CPU with
metadata only (fake, suitably aligned device addresses; no memory is touched):

  * Llama-3-8B FSDP(N)->TP(N): the rectangles the product derives equal the reference's plan
    statistics (tests/golden/direct_plan.json: op counts, exact ops, algorithmic bytes) -- i.e. no
    read amplification -- and the compiled tile table covers every unit of every rectangle exactly
    once, in an order that keeps all source GPUs interleaved.
"""

import json
import os

import numpy as np
import pytest
import torch

import workloads
from torchstore_b200 import _native
from torchstore_b200.planner import StridedMem, build_rects

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
BASE = 0x7F0000000000


def contig(shape):
    out, s = [], 1
    for e in reversed(shape):
        out.append(s)
        s *= e
    return tuple(reversed(out))


def fake_windows(layout, n, dest_rank):
    """(src window, dst window) per reference plan op, laid out like bench.py's flat buffers."""
    src_base = {}  # (name, rank) -> (ptr, shape)
    cursor = {r: BASE + (r << 36) for r in range(n)}
    for name, (shape, tp) in layout.items():
        for r in range(n):
            off, shp = workloads.shard_box(shape, n, r, ("S", 0)) if n > 1 else ((0,) * len(shape), tuple(shape))
            src_base[(name, r)] = (cursor[r], shp)
            cursor[r] += (int(np.prod(shp)) * 2 + 127) // 128 * 128
    dst_cursor = BASE + (15 << 36)
    pairs, exact_ops = [], 0
    dst_of = {}
    for name, (shape, tp) in layout.items():
        off, shp = workloads.shard_box(shape, n, dest_rank, tp) if n > 1 else ((0,) * len(shape), tuple(shape))
        dst_of[name] = (dst_cursor, shp)
        dst_cursor += (int(np.prod(shp)) * 2 + 127) // 128 * 128
    for name, srank, s_idx, d_idx, exact in workloads.fsdp_to_tp_rects(layout, n, dest_rank):
        sp, sshape = src_base[(name, srank)]
        dp, dshape = dst_of[name]
        src = StridedMem(sp, sshape, contig(sshape), torch.bfloat16, srank)
        dst = StridedMem(dp, dshape, contig(dshape), torch.bfloat16, dest_rank)
        pairs.append((src.sub(tuple(slice(a, b) for a, b in s_idx)), dst.sub(tuple(slice(a, b) for a, b in d_idx))))
        exact_ops += bool(exact)
    return pairs, exact_ops


DEV_RECT = np.dtype([("src", "<u8"), ("dst", "<u8"), ("ss", "<i8", 6), ("ds", "<i8", 6), ("ext", "<u4", 6),
                     ("n_outer", "<u4"), ("rows", "<u4"), ("upr", "<u4"), ("magic", "<u4"), ("wide", "<u4"),
                     ("split", "<u4"), ("mode", "<u4"), ("sub", "<u4"), ("dub", "<u4"), ("tile_units", "<u4"),
                     ("link", "<u4"), ("pad", "<u4", 3)])
assert DEV_RECT.itemsize == 192


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_llama3_8b_plans_match_reference_statistics_and_cover_exactly(n):
    gold = {e["n"]: e["per_dest_rank"] for e in json.load(open(os.path.join(GOLDEN, "direct_plan.json")))["llama3_8b_fsdp_to_tp"]}
    layout = workloads.llama_layout()
    assert workloads.state_dict_bytes(layout) == 16060522496
    for dest_rank in sorted({0, n - 1, n // 2}):
        pairs, exact_ops = fake_windows(layout, n, dest_rank)
        want = gold[n][dest_rank]
        assert len(pairs) == want["ops"] and exact_ops == want["exact_ops"]
        rects, k = build_rects(pairs)
        assert k == len(pairs)
        table, tiles, info = _native.plan_compile_host(dest_rank, rects, k, 0, 0)
        # no read amplification: bytes read == bytes written == the reference's algorithmic bytes
        assert info.payload_bytes == want["algorithmic_bytes"] == info.src_bytes
        assert info.num_vector_rects == info.num_rects  # everything moves as 16-byte units
        local = sum(int(np.prod(s.shape)) * 2 for s, _ in pairs if s.device == dest_rank)
        assert info.remote_src_bytes == info.src_bytes - local
        dr = table.view(DEV_RECT)
        # two queues: remote 16-byte rects ride the link queue (TMA bulk ring, small tiles), the rest
        # the copy queue; the tile list is copy queue first, then link queue
        is_link = dr["link"] == 1
        src_dev = np.array([p[0].device for p in pairs])
        assert len(dr) == len(pairs)  # rects map 1:1 to pairs here (no splitting at these sizes)
        assert np.array_equal(is_link, src_dev != dest_rank)
        assert info.link_bytes == info.remote_src_bytes
        if n > 1:  # ring depth follows the fan-in: 6 stages from 4 source GPUs up, 3 below
            assert info.link_stages == (6 if n - 1 >= 4 else 3)
        tile_units = np.where(is_link, info.link_tile_bytes // 16, info.tile_bytes // 16)
        assert np.array_equal(dr["tile_units"], tile_units)
        assert len(tiles) == info.num_tiles + info.num_link_tiles
        assert not is_link[tiles[: info.num_tiles, 0]].any() and is_link[tiles[info.num_tiles:, 0]].all()
        # a link tile is never larger than one ring stage
        stage_units = info.link_tile_bytes // 16
        lk = dr[is_link]
        assert np.all(np.where(lk["wide"] == 1, stage_units, lk["split"].astype(np.int64) * lk["upr"]) <= stage_units)
        assert np.all((lk["wide"] == 1) | (lk["n_outer"] <= 1))
        # per-rect tile counts, every (rect, tile) exactly once
        per_rect = np.bincount(tiles[:, 0], minlength=len(dr))
        wide = dr["wide"] == 1
        expect = np.where(wide, dr["rows"].astype(np.int64) * dr["split"], -(-dr["rows"].astype(np.int64) // dr["split"]))
        assert np.array_equal(per_rect, expect)
        key = tiles[:, 0].astype(np.int64) * (1 << 32) + tiles[:, 1]
        assert len(np.unique(key)) == len(key)
        # units covered == units of the rect (wide: segments of tile_units; narrow: whole rows)
        assert int((dr["rows"].astype(np.int64) * dr["upr"] * dr["dub"]).sum()) == info.payload_bytes
        assert np.all(dr["split"][wide] == -(-dr["upr"][wide].astype(np.int64) // tile_units[wide]))
        assert np.all(dr["upr"][~wide] < tile_units[~wide])
        # multiply-high division exact over the index range a narrow tile can see
        for r in dr[~wide][:50]:
            u = int(r["upr"])
            if u == 1:
                continue
            idx = np.arange(0, int(r["split"]) * u, dtype=np.uint64)
            assert np.array_equal((idx * np.uint64(r["magic"])) >> np.uint64(32), idx // np.uint64(u))
        # source interleave: within any window of 4*n consecutive tiles (while all sources still have
        # work) every source GPU appears
        if n > 1:
            # link queue: granules of 8 tiles (one claim) rotate over the n-1 source GPUs
            devs = src_dev[tiles[info.num_tiles:, 0]]
            assert (devs != dest_rank).all()
            head = devs[: len(devs) // 2]
            win = 8 * 4 * n
            for start in range(0, max(1, len(head) - win), max(1, len(head) // 64)):
                assert len(set(head[start:start + win].tolist())) == n - 1


def test_config2_single_4gib_rect_compiles_to_one_wide_row():
    src = StridedMem(BASE, (2147483648,), (1,), torch.bfloat16, 0)
    dst = StridedMem(BASE + (1 << 40), (2147483648,), (1,), torch.bfloat16, 1)
    rects, k = build_rects([(src, dst)])
    table, tiles, info = _native.plan_compile_host(1, rects, k, 0, 0)
    assert info.payload_bytes == 4 << 30 and info.remote_src_bytes == 4 << 30
    assert info.num_rects == 1 and info.num_tiles == 0 and info.num_link_tiles == (4 << 30) // info.link_tile_bytes
    dr = table.view(DEV_RECT)[0]
    assert dr["link"] == 1
    assert dr["wide"] == 1 and dr["rows"] == 1 and dr["upr"] == (4 << 30) // 16 and dr["mode"] == 4


def test_oversize_rects_are_split():
    # a 40 GiB contiguous region exceeds 2^30 units per row -> split into column chunks
    n = 40 << 30
    src = StridedMem(BASE, (n,), (1,), torch.uint8, 0)
    dst = StridedMem(BASE + (1 << 41), (n,), (1,), torch.uint8, 0)
    rects, k = build_rects([(src, dst)])
    table, tiles, info = _native.plan_compile_host(0, rects, k, 0, 0)
    assert info.payload_bytes == n and info.num_rects == 3
    dr = table.view(DEV_RECT)
    assert int((dr["upr"].astype(np.int64) * 16).sum()) == n
    assert list(dr["src"]) == [BASE, BASE + (16 << 30), BASE + (32 << 30)]
