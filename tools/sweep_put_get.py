"""BASELINE config #5 (development tool, needs a GPU): put_batch / get_batch bandwidth through the
store path (HBM volumes + copy_rects) as a function of key size, constant total bytes.

  python tools/sweep_put_get.py [--total-gib 8] [--out gpurun_out/put_get.json]

Single process; the volume lives on GPU 0 (local D2D put) and, with 2+ GPUs, a second client on GPU 1
reads over NVLink.  Reports GB/s per key size (1 MiB .. 8 GiB)."""

from __future__ import annotations

import argparse
import asyncio
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import torchstore_b200 as ts  # noqa: E402


async def run(total_bytes: int, out_path: str | None):
    os.environ.pop("RANK", None)
    os.environ["LOCAL_RANK"] = "0"
    await ts.initialize()
    rows = []
    reader_dev = "cuda:1" if torch.cuda.device_count() > 1 else "cuda:0"
    try:
        size = 1 << 20
        while size <= min(total_bytes, 8 << 30):
            n = max(1, total_bytes // size)
            src = {f"k{i}": torch.empty(size // 2, dtype=torch.bfloat16, device="cuda:0").normal_() for i in range(n)}
            dst = {k: torch.empty_like(v, device=reader_dev) for k, v in src.items()}
            await ts.put_batch(src)  # first put allocates arena blocks
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            await ts.put_batch(src)  # overwrite in place
            t_put = time.perf_counter() - t0
            t0 = time.perf_counter()
            await ts.get_batch(dst)
            t_get = time.perf_counter() - t0
            ok = all(torch.equal(dst[k].cpu(), src[k].cpu()) for k in list(src)[:2])
            row = {"key_bytes": size, "keys": n, "put_GBps": round(n * size / t_put / 1e9, 1),
                   "get_GBps": round(n * size / t_get / 1e9, 1), "reader": reader_dev, "ok": ok}
            rows.append(row)
            print(json.dumps(row), flush=True)
            await ts.delete_batch(list(src))
            del src, dst
            size *= 4
    finally:
        await ts.shutdown()
    if out_path:
        with open(out_path, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-gib", type=float, default=8.0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    asyncio.run(run(int(a.total_gib * (1 << 30)), a.out))
