"""Kernel-level microbenchmarks for copy_rects (development tool, not the contract bench).

  python tools/microbench.py [--out gpurun_out/microbench.json]

Measures, with CUDA events on the launching stream and an L2 flush between iterations:
  * big contiguous D2D copy (HBM roofline point) vs torch copy_ (the measured-peak method)
  * narrow-row reshard rectangles (wo / w2 shapes of Llama-3-8B FSDP8->TP8)
  * fused fp32->bf16 cast
  * sweeps of TSB_CTAS_PER_SM and TSB_TILE_BYTES
"""

from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from torchstore_b200 import _native  # noqa: E402
from torchstore_b200.planner import StridedMem, build_rects  # noqa: E402


def time_plan(pairs, iters=10, warmup=3, flush=None):
    rects, n = build_rects([(StridedMem.from_tensor(s), StridedMem.from_tensor(d)) for s, d in pairs])
    plan = _native.plan_create(0, rects, n)
    info = _native.plan_info(plan).as_dict()
    stream = _native.torch_stream(0)
    times = []
    for i in range(warmup + iters):
        if flush is not None:
            flush.add_(1)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.plan_run(plan, stream)
        e1.record()
        e1.synchronize()
        if i >= warmup:
            times.append(e0.elapsed_time(e1))
    _native.plan_destroy(plan)
    times.sort()
    return times[len(times) // 2], times[0], info


def time_torch_copy(dst, src, iters=10, warmup=3, flush=None):
    times = []
    for i in range(warmup + iters):
        if flush is not None:
            flush.add_(1)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        dst.copy_(src)
        e1.record()
        e1.synchronize()
        if i >= warmup:
            times.append(e0.elapsed_time(e1))
    times.sort()
    return times[len(times) // 2], times[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    _native.init()
    res = {"device": torch.cuda.get_device_name(0), "results": []}
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def rec(name, ms_med, ms_min, payload, extra=None):
        row = {"name": name, "ms_median": round(ms_med, 4), "ms_min": round(ms_min, 4),
               "payload_GBps": round(payload / ms_med / 1e6, 1), "rw_GBps": round(2 * payload / ms_med / 1e6, 1)}
        if extra:
            row.update(extra)
        res["results"].append(row)
        print(json.dumps(row), flush=True)

    # ---- 1. big contiguous copy: ours vs torch copy_ ------------------------------------------
    n = 2 << 30  # 2 GiB payload
    src = torch.empty(n, dtype=torch.uint8, device="cuda").random_(0, 255)
    dst = torch.empty_like(src)
    med, mn = time_torch_copy(dst, src, flush=flush)
    rec("torch_copy_2GiB", med, mn, n)
    for per_sm in ([4] if args.quick else [1, 2, 3, 4, 6, 8]):
        for tile in ([65536] if args.quick else [16384, 32768, 65536, 131072, 262144]):
            os.environ["TSB_CTAS_PER_SM"] = str(per_sm)
            os.environ["TSB_TILE_BYTES"] = str(tile)
            med, mn, info = time_plan([(src, dst)], flush=flush)
            rec(f"copy_rects_2GiB_contig", med, mn, n, {"ctas_per_sm": per_sm, "tile_bytes": tile, "grid": info["grid"]})
    assert torch.equal(src, dst)
    os.environ["TSB_CTAS_PER_SM"] = "4"
    os.environ["TSB_TILE_BYTES"] = "65536"
    del src, dst

    # ---- 2. narrow-row reshard rectangles (bf16) ------------------------------------------------
    for label, cols_total, cols, nlayers in (("wo_1KiB_rows", 4096, 512, 32), ("w2_3.5KiB_rows", 14336, 1792, 16)):
        srcs = [torch.randn(nlayers, 512, cols_total, device="cuda").to(torch.bfloat16) for _ in range(8)]
        dst = torch.zeros(nlayers, 4096, cols, dtype=torch.bfloat16, device="cuda")
        pairs = []
        for layer in range(nlayers):
            for s in range(8):
                pairs.append((srcs[s][layer][:, 3 * cols:4 * cols], dst[layer][s * 512:(s + 1) * 512]))
        payload = dst.numel() * 2
        for per_sm, tile in ([(4, 65536)] if args.quick else [(2, 65536), (4, 32768), (4, 65536), (4, 131072), (8, 65536)]):
            os.environ["TSB_CTAS_PER_SM"] = str(per_sm)
            os.environ["TSB_TILE_BYTES"] = str(tile)
            med, mn, info = time_plan(pairs, flush=flush)
            rec(f"reshard_{label}", med, mn, payload, {"ctas_per_sm": per_sm, "tile_bytes": tile, "rects": info["num_rects"], "tiles": info["num_tiles"]})
        want = torch.cat([s[:, :, 3 * cols:4 * cols] for s in srcs], dim=1)
        assert torch.equal(dst, want)
        del srcs, dst, want
    os.environ["TSB_CTAS_PER_SM"] = "4"
    os.environ["TSB_TILE_BYTES"] = "65536"

    # ---- 3. fused cast fp32 -> bf16 -----------------------------------------------------------------
    m = torch.randn(512 << 20, device="cuda")  # 2 GiB fp32
    out = torch.empty(m.shape, dtype=torch.bfloat16, device="cuda")
    med, mn, info = time_plan([(m, out)], flush=flush)
    row_bytes = m.numel() * 6
    res["results"].append({"name": "cast_f32_bf16_2GiB_in", "ms_median": round(med, 4), "hbm_GBps": round(row_bytes / med / 1e6, 1)})
    print(json.dumps(res["results"][-1]), flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.add_(1)
    e0.record()
    ref = m.to(torch.bfloat16)
    e1.record()
    e1.synchronize()
    res["results"].append({"name": "torch_to_bf16_2GiB_in", "ms": round(e0.elapsed_time(e1), 4), "hbm_GBps": round(row_bytes / e0.elapsed_time(e1) / 1e6, 1)})
    print(json.dumps(res["results"][-1]), flush=True)
    assert torch.equal(ref, out)

    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
