"""PCIe H2D rate of one GPU from a NUMA-local pinned buffer: one copy vs the same bytes split over
several streams / copy engines (development probe for bench.py's e2e leg)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from torchstore_b200 import _native, numa  # noqa: E402

dev = 0
torch.cuda.set_device(dev)
_native.init()
print(json.dumps({"numa": numa.bind_to_gpu_numa(dev)}))
nbytes = 4 << 30
host = numa.pinned_like(nbytes, torch.uint8)
dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
streams = [_native.copy_stream(dev)] + [torch.cuda.Stream().cuda_stream for _ in range(7)]
for nstreams in (1, 2, 3, 4, 8):
    part = nbytes // nstreams // 4096 * 4096
    best = 1e9
    for it in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(nstreams):
            _native.memcpy_async(dev, dst.data_ptr() + i * part, host.data_ptr() + i * part, part, _native.TSB_H2D, streams[i])
        for i in range(nstreams):
            _native.stream_sync(dev, streams[i])
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"streams": nstreams, "GBps": round(part * nstreams / best / 1e9, 2)}), flush=True)
# torch's own pinned allocation for comparison (wherever the driver placed it)
t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
best = 1e9
for it in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dst.copy_(t, non_blocking=True)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
print(json.dumps({"torch_pin_memory_copy": round(nbytes / best / 1e9, 2)}))
