"""Sweep copy_rects launch parameters on the real state-dict plans (development tool).

  python tools/sweep_plan.py --mode n1      Llama-3-8B 1->1 (291 exact copies, 16 GB payload)
  python tools/sweep_plan.py --mode tp8     dest rank 0 of FSDP(8)->TP(8) with all 8 source shards resident
                                            on this GPU (1194 rects, 2 GB payload, narrow rows)
  python tools/sweep_plan.py --mode x2 --n 8   2-GPU emulation of the N-rank sync: GPU0 plays dest rank 0, GPU1
                                            plays dest rank 1; each one's N-1 "remote" source shards live on the
                                            OTHER GPU, so every launch has the real local/NVLink byte mix and both
                                            HBMs serve a peer while copying (needs 2 GPUs; --env sweeps env settings)
"""

from __future__ import annotations

import argparse
import asyncio
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import workloads  # noqa: E402
from torchstore_b200 import _native  # noqa: E402
from torchstore_b200.direct_weight_sync import DirectWeightSyncDest, NvlinkBuffer, RDMAWeightHandle  # noqa: E402
from torchstore_b200.transport.types import TensorSlice  # noqa: E402


def build(mode: str):
    dev = torch.device("cuda", 0)
    layout = workloads.llama_layout()
    n = 1 if mode == "n1" else 8
    handles, dest, dslices, keep = {}, {}, {}, []
    for name, (shape, tp) in layout.items():
        hl = []
        for r in range(n):
            off, shp = workloads.shard_box(shape, n, r, ("S", 0)) if n > 1 else ((0,) * len(shape), tuple(shape))
            t = torch.empty(shp, dtype=torch.bfloat16, device=dev).normal_(0, 0.02)
            keep.append(t)
            hl.append(RDMAWeightHandle(NvlinkBuffer(t), TensorSlice(off, (r,), tuple(shape), shp, (n,)), r))
        handles[name] = hl
        doff, dshp = workloads.shard_box(shape, n, 0, tp) if n > 1 else ((0,) * len(shape), tuple(shape))
        dest[name] = torch.zeros(dshp, dtype=torch.bfloat16, device=dev)
        if n > 1:
            dslices[name] = TensorSlice(doff, (0,), tuple(shape), dshp, (n,))
    return handles, dest, dslices or None, keep


def build_x2(n: int):
    """Two destination ranks (0 on GPU0, 1 on GPU1); rank r's local source shard on its own GPU, the
    other n-1 source shards on the other GPU."""
    layout = workloads.llama_layout()
    sides = []
    keep = []
    for drank, (own, other) in enumerate(((0, 1), (1, 0))):
        handles, dest, dslices = {}, {}, {}
        for name, (shape, tp) in layout.items():
            hl = []
            for r in range(n):
                off, shp = workloads.shard_box(shape, n, r, ("S", 0))
                dev = torch.device("cuda", own if r == drank else other)
                t = torch.empty(shp, dtype=torch.bfloat16, device=dev).normal_(0, 0.02)
                keep.append(t)
                hl.append(RDMAWeightHandle(NvlinkBuffer(t), TensorSlice(off, (r,), tuple(shape), shp, (n,)), r))
            handles[name] = hl
            doff, dshp = workloads.shard_box(shape, n, drank, tp)
            dest[name] = torch.zeros(dshp, dtype=torch.bfloat16, device=torch.device("cuda", own))
            dslices[name] = TensorSlice(doff, (drank,), tuple(shape), dshp, (n,))
        sides.append((own, handles, dest, dslices))
    return sides, keep


def check_x2(sides):
    """Bit-exact, order-sensitive: every destination equals the slices of the sources it was built from."""
    for own, handles, dest, dslices in sides:
        for name, d in dest.items():
            ds = dslices[name]
            want = torch.empty_like(d)
            for h in handles[name]:
                ss = h.tensor_slice
                lo = [max(a, b) for a, b in zip(ss.offsets, ds.offsets)]
                hi = [min(a + x, b + y) for a, x, b, y in zip(ss.offsets, ss.local_shape, ds.offsets, ds.local_shape)]
                if any(h_ <= l_ for l_, h_ in zip(lo, hi)):
                    continue
                src = h.rdma_buffer._keepalive
                s_idx = tuple(slice(l_ - o, h_ - o) for l_, h_, o in zip(lo, hi, ss.offsets))
                d_idx = tuple(slice(l_ - o, h_ - o) for l_, h_, o in zip(lo, hi, ds.offsets))
                want[d_idx] = src[s_idx].to(want.device)
            assert torch.equal(want.view(torch.int16), d.view(torch.int16)), f"mismatch in {name} on GPU {own}"


def run_x2(args):
    assert torch.cuda.device_count() >= 2, "--mode x2 needs 2 GPUs"
    _native.enable_peer_access(0, 1)
    sides, keep = build_x2(args.n)
    rows = []
    envs = [e for e in args.env.split(";") if e] or [""]
    for env in envs:
        saved = {}
        for kv in env.split(","):
            if kv:
                k, v = kv.split("=")
                saved[k] = os.environ.get(k)
                os.environ[k] = v
        syncs = [DirectWeightSyncDest() for _ in sides]
        for own, _, dest, _ in sides:
            for d in dest.values():
                d.zero_()
        times = []

        async def both():
            for sync, (own, handles, dest, dslices) in zip(syncs, sides):
                if sync._plan is None:
                    with torch.cuda.device(own):
                        await sync.pull(handles, dest, dslices)
            torch.cuda.synchronize(0)
            torch.cuda.synchronize(1)
            active = syncs[:1] if args.solo else syncs   # --solo: GPU1 stays idle while GPU0 pulls from it
            for sync in active:
                sync.launch()
            for sync in active:
                await sync.wait()

        for i in range(args.iters + 2):
            asyncio.run(both())
            if i >= 2:
                times.append(max(s.last_pull_ms[own] for s, (own, *_r) in list(zip(syncs, sides))[: 1 if args.solo else 2]))
        check_x2(sides)
        info = syncs[0].plan_info()[0]
        for s in syncs:
            s.close()
        times.sort()
        med = times[len(times) // 2]
        local_read = info["src_bytes"] - info["remote_src_bytes"]
        row = {"mode": f"x2_n{args.n}" + ("_solo" if args.solo else ""), "env": env, "ms_median": round(med, 4), "ms_min": round(times[0], 4),
               "nvlink_in_GBps": round(info["remote_src_bytes"] / med / 1e6, 1),
               # HBM traffic of one GPU: own local reads + own writes + the peer's reads of our memory
               "hbm_GBps": round((local_read + info["payload_bytes"] + info["remote_src_bytes"]) / med / 1e6, 1),
               "payload": info["payload_bytes"], "remote": info["remote_src_bytes"], "link_bytes": info["link_bytes"],
               "tiles": info["num_tiles"], "link_tiles": info["num_link_tiles"], "grid": info["grid"], "block": info["block"],
               "parity": "bit-exact"}
        rows.append(row)
        print(json.dumps(row), flush=True)
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="n1")
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--env", default="", help="x2 mode: ';'-separated settings, each 'K=V,K=V' (empty = defaults)")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--solo", action="store_true",
                    help="x2 mode: launch only GPU0's plan (its peer is idle): the kernel without source-side contention")
    ap.add_argument("--grid", default="2x65536,3x32768,3x65536,3x131072,4x32768,4x65536,4x131072,6x65536,8x32768,8x65536")
    args = ap.parse_args()
    _native.init()
    if args.mode == "x2":
        run_x2(args)
        return
    handles, dest, dslices, keep = build(args.mode)
    payload = sum(v.numel() * 2 for v in dest.values())
    rows = []
    for combo in args.grid.split(","):
        per_sm, tile = combo.split("x")
        os.environ["TSB_CTAS_PER_SM"] = per_sm
        os.environ["TSB_TILE_BYTES"] = tile
        sync = DirectWeightSyncDest()
        times = []
        for i in range(args.iters + 2):
            asyncio.run(sync.pull(handles, dest, dslices))
            if i >= 2:
                times.append(sync.last_pull_ms[0])
        info = sync.plan_info()[0]
        sync.close()
        times.sort()
        med = times[len(times) // 2]
        row = {"mode": args.mode, "ctas_per_sm": int(per_sm), "tile_bytes": int(tile), "ms_median": round(med, 4),
               "ms_min": round(times[0], 4), "payload_GBps": round(payload / med / 1e6, 1),
               "rw_GBps": round(2 * payload / med / 1e6, 1), "tiles": info["num_tiles"], "rects": info["num_rects"]}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
