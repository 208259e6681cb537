"""Bidirectional NVLink pulls (2 GPUs, one process): GPU0 pulls a buffer from GPU1 WHILE GPU1 pulls one from
GPU0 -- nothing else touches either HBM.  Separates "the source GPU's HBM is busy" from "the link carries
read requests one way and read responses the other way at the same time".

  python tools/p2p_bidir.py [--out gpurun_out/p2p_bidir.json] [--one-way]   (--one-way: ncu target, single pull)
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


def make_plan(dev, src, dst):
    rects, n = build_rects([(StridedMem.from_tensor(src), StridedMem.from_tensor(dst))])
    return _native.plan_create(dev, rects, n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--one-way", action="store_true")
    ap.add_argument("--gib", type=int, default=4)
    args = ap.parse_args()
    _native.init()
    _native.enable_peer_access(0, 1)
    n = args.gib << 30
    a0 = torch.empty(n // 2, dtype=torch.int16, device="cuda:0").random_()
    a1 = torch.empty(n // 2, dtype=torch.int16, device="cuda:1").random_()
    b0 = torch.zeros_like(a0)  # on GPU0, filled from GPU1
    b1 = torch.zeros_like(a1)  # on GPU1, filled from GPU0
    rows = []
    settings = [("TSB_LINK", "0"), ("TSB_LINK", "1")] if args.one_way else \
        [("TSB_LINK=0", None), ("TSB_LINK=1,TSB_LINK_STAGES=3", None), ("TSB_LINK=1,TSB_LINK_STAGES=6", None),
         ("TSB_LINK=1,TSB_LINK_STAGES=6,TSB_LINK_STAGE_BYTES=8192", None), ("TSB_LINK=1,TSB_LINK_STAGES=4,TSB_LINK_STAGE_BYTES=16384", None),
         ("TSB_LINK=1,TSB_LINK_STAGES=8,TSB_LINK_STAGE_BYTES=2048", None), ("TSB_LINK=1,TSB_LINK_STAGES=3,TSB_CTAS_PER_SM=1", None),
         ("TSB_LINK=0,TSB_CTAS_PER_SM=1", None)]
    for setting, val in settings:
        env = {setting: val} if val is not None else dict(kv.split("=") for kv in setting.split(","))
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        p0 = make_plan(0, a1, b0)
        p1 = make_plan(1, a0, b1)
        res = {"setting": ",".join(f"{k}={v}" for k, v in env.items())}
        for mode in (("one_way",) if args.one_way else ("one_way", "both_ways")):
            times = []
            for it in range(5):
                torch.cuda.synchronize(0)
                torch.cuda.synchronize(1)
                _native.plan_launch(p0, None)
                if mode == "both_ways":
                    _native.plan_launch(p1, None)
                _native.plan_wait(p0)
                ms = _native.plan_elapsed_ms(p0)
                if mode == "both_ways":
                    _native.plan_wait(p1)
                    ms = max(ms, _native.plan_elapsed_ms(p1))
                if it >= 2:
                    times.append(ms)
            times.sort()
            res[mode + "_GBps_per_direction"] = round(n / times[len(times) // 2] / 1e6, 1)
        assert torch.equal(b0, a1.to("cuda:0")) and (args.one_way or torch.equal(b1, a0.to("cuda:1")))
        rows.append(res)
        print(json.dumps(res), flush=True)
        _native.plan_destroy(p0)
        _native.plan_destroy(p1)
        for k, v in saved.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
