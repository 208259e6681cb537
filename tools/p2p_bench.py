"""NVLink P2P microbenchmark (2 GPUs, one process): GPU1 pulls from GPU0 with copy_rects.

  python tools/p2p_bench.py [--out gpurun_out/p2p.json]

Reports GB/s per direction for a 4 GiB contiguous pull (BASELINE config #2), narrow-row pulls
(wo / w2 rectangles) and torch's cudaMemcpyPeer for reference; sweeps CTAs/SM and tile size.
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


def time_plan(dev, pairs, iters=6, warmup=2):
    rects, n = build_rects([(StridedMem.from_tensor(s), StridedMem.from_tensor(d)) for s, d in pairs])
    plan = _native.plan_create(dev, rects, n)
    info = _native.plan_info(plan).as_dict()
    times = []
    for i in range(warmup + iters):
        e0 = _native.Event(dev, timing=True).record(None)
        _native.plan_run(plan, None)
        e1 = _native.Event(dev, timing=True).record(None)
        e1.synchronize()
        if i >= warmup:
            times.append(e0.elapsed_ms(e1))
    _native.plan_destroy(plan)
    times.sort()
    return times[len(times) // 2], info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    _native.init()
    assert torch.cuda.device_count() >= 2
    _native.enable_peer_access(1, 0)
    rows = []

    def rec(**kw):
        rows.append(kw)
        print(json.dumps(kw), flush=True)

    n = 4 << 30
    src = torch.empty(n // 2, dtype=torch.bfloat16, device="cuda:0").normal_()
    dst = torch.zeros(n // 2, dtype=torch.bfloat16, device="cuda:1")
    # torch copy (cudaMemcpyPeerAsync)
    with torch.cuda.device(1):
        for i in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src)
            e1.record()
            e1.synchronize()
        rec(name="torch_copy_peer_4GiB", ms=round(e0.elapsed_time(e1), 4), GBps=round(n / e0.elapsed_time(e1) / 1e6, 1))
    # link=0: copy warps pull with LDG.128; link=1: link warp's TMA bulk ring (default)
    for link, stage, stages in ((0, 8192, 6), (1, 8192, 6), (1, 16384, 4), (1, 4096, 8)):
        for per_sm in (2, 3, 4):
            os.environ["TSB_LINK"] = str(link)
            os.environ["TSB_LINK_STAGE_BYTES"] = str(stage)
            os.environ["TSB_LINK_STAGES"] = str(stages)
            os.environ["TSB_CTAS_PER_SM"] = str(per_sm)
            dst.zero_()
            ms, info = time_plan(1, [(src, dst)])
            rec(name="pull_contig_4GiB", link=link, stage_bytes=stage, stages=stages, ctas_per_sm=per_sm, ms=round(ms, 4),
                GBps=round(n / ms / 1e6, 1), grid=info["grid"], block=info["block"])
    os.environ["TSB_LINK_STAGE_BYTES"] = "8192"
    os.environ["TSB_LINK_STAGES"] = "6"
    os.environ["TSB_LINK"] = "1"
    assert torch.equal(src.cpu()[:1 << 20], dst.cpu()[:1 << 20]) and int(dst.view(torch.int16).to(torch.int64).sum()) == int(src.view(torch.int16).to(torch.int64).sum())
    # push (GPU0 writes into GPU1 memory): store-path puts to a remote volume
    os.environ["TSB_CTAS_PER_SM"] = "3"
    os.environ["TSB_TILE_BYTES"] = "65536"
    dst.zero_()
    ms, _ = time_plan(0, [(src, dst)])
    rec(name="push_contig_4GiB", ms=round(ms, 4), GBps=round(n / ms / 1e6, 1))
    del src, dst

    for label, cols_total, cols, nl in (("wo_1KiB_rows", 4096, 512, 32), ("w2_3.5KiB_rows", 14336, 1792, 16)):
        srcs = torch.randn(nl, 8, 512, cols_total, device="cuda:0").to(torch.bfloat16)
        dstt = torch.zeros(nl, 4096, cols, dtype=torch.bfloat16, device="cuda:1")
        pairs = [(srcs[layer, s][:, 3 * cols:4 * cols], dstt[layer][s * 512:(s + 1) * 512]) for layer in range(nl) for s in range(8)]
        payload = dstt.numel() * 2
        for link in (0, 1):
            for per_sm in (2, 3, 4):
                os.environ["TSB_LINK"] = str(link)
                os.environ["TSB_CTAS_PER_SM"] = str(per_sm)
                dstt.zero_()
                ms, info = time_plan(1, pairs)
                rec(name=f"pull_{label}", link=link, ctas_per_sm=per_sm, ms=round(ms, 4), GBps=round(payload / ms / 1e6, 1),
                    rects=info["num_rects"])
        os.environ["TSB_LINK"] = "1"
        want = torch.cat([srcs[:, s, :, 3 * cols:4 * cols] for s in range(8)], dim=1)
        assert torch.equal(dstt.cpu(), want.cpu())
        del srcs, dstt, want
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
