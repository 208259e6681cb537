"""Sweep copy_rects launch parameters on the real state-dict plans (development tool).

  python tools/sweep_plan.py --mode n1      Llama-3-8B 1->1 (291 exact copies, 16 GB payload)
  python tools/sweep_plan.py --mode tp8     dest rank 0 of FSDP(8)->TP(8) with all 8 source shards resident
                                            on this GPU (1194 rects, 2 GB payload, narrow rows)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="n1")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--grid", default="2x65536,3x32768,3x65536,3x131072,4x32768,4x65536,4x131072,6x65536,8x32768,8x65536")
    args = ap.parse_args()
    _native.init()
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
