"""TEST INFRASTRUCTURE ONLY -- how much does bench.py's CPU "port" favour the reference?

Runs, in the build container (needs /root/reference), the UNMODIFIED reference
(LocalClient + Controller + InMemoryStore + SharedMemoryTransportBuffer through the in-process
harness) and bench.py's port (oracle/copy_rects_ref.c, all host threads) on the SAME sample:
put_state_dict from N source ranks + get_state_dict(user_state_dict) into N destination ranks of
the first L Llama-3-8B layers, FSDP(N) Shard(0) -> TP(N), CPU tensors.

    python oracle/calibrate_port.py [--n 2] [--layers 1] --out profiles/r1_port_vs_reference_cpu.json
"""

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

import workloads  # noqa: E402
from oracle import ref_harness  # noqa: E402


async def run_reference(n: int, layers: int, iters: int):
    ref_harness.import_reference()
    from torchstore import state_dict_utils
    from torchstore.transport.types import Request, TensorSlice

    layout = workloads.llama_layout(n_layers=layers, with_embeddings=False)
    store = ref_harness.RefStore(n)
    gen = torch.Generator().manual_seed(0)
    fulls = {k: (torch.randn(shape, generator=gen) * 0.02).to(torch.bfloat16) for k, (shape, _) in layout.items()}
    nbytes = sum(v.numel() * 2 for v in fulls.values())

    # the reference's put_state_dict/get_state_dict need DTensors for sharded leaves; feed the same
    # Requests its from_dtensor() would build (types.py:176-196) through the same client entry points
    async def put_rank(rank):
        c = store.client(rank)
        reqs = []
        for name, (shape, _) in layout.items():
            off, shp = workloads.shard_box(shape, n, rank, ("S", 0)) if n > 1 else ((0,) * len(shape), tuple(shape))
            local = fulls[name][tuple(slice(o, o + s) for o, s in zip(off, shp))].contiguous()
            ts = TensorSlice(off, (rank,), tuple(shape), shp, (n,)) if n > 1 else None
            reqs.append(Request(key=f"sd/{name}", tensor_val=local, tensor_slice=ts))
        from torchstore.transport import create_transport_buffer

        ref = c.strategy.select_storage_volume()
        await create_transport_buffer(ref).put_to_storage_volume(reqs)
        await c._controller.notify_put_batch.call([r.meta_only() for r in reqs], ref.volume_id)

    async def get_rank(rank, dests):
        c = store.client(rank)
        reqs = []
        for name, (shape, tp) in layout.items():
            off, shp = workloads.shard_box(shape, n, rank, tp) if n > 1 else ((0,) * len(shape), tuple(shape))
            ts = TensorSlice(off, (rank,), tuple(shape), shp, (n,)) if n > 1 else None
            reqs.append(Request(key=f"sd/{name}", tensor_val=dests[name], tensor_slice=ts))
        results = await c._fetch(reqs)
        for r in reqs:
            c._apply_inplace(results[r.key], dests[r.key.split("/", 1)[1]] if False else r.tensor_val, r)

    dests = []
    for rank in range(n):
        d = {}
        for name, (shape, tp) in layout.items():
            off, shp = workloads.shard_box(shape, n, rank, tp) if n > 1 else ((0,) * len(shape), tuple(shape))
            d[name] = torch.zeros(shp, dtype=torch.bfloat16)
        dests.append(d)
    times = []
    for it in range(iters + 1):
        t0 = time.perf_counter()
        for rank in range(n):
            await put_rank(rank)
        for rank in range(n):
            await get_rank(rank, dests[rank])
        if it > 0:  # first iteration allocates + prefaults the shm segments (cold)
            times.append(time.perf_counter() - t0)
        else:
            cold = time.perf_counter() - t0
    # correctness of what we timed
    for rank in range(n):
        for name, (shape, tp) in layout.items():
            off, shp = workloads.shard_box(shape, n, rank, tp) if n > 1 else ((0,) * len(shape), tuple(shape))
            assert torch.equal(dests[rank][name], fulls[name][tuple(slice(o, o + s) for o, s in zip(off, shp))])
    store.close()
    times.sort()
    return nbytes, times[len(times) // 2], cold


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--layers", type=int, default=1)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    nbytes, warm, cold = asyncio.run(run_reference(a.n, a.layers, a.iters))
    import bench

    port = bench.cpu_reference_run(a.n, steps=a.iters, warmup=1, budget_s=30.0, layers=a.layers)
    res = {
        "host": {"cores": os.cpu_count(), "torch_threads": torch.get_num_threads()},
        "sample": f"{a.layers} llama3-8b layer(s), FSDP({a.n})->TP({a.n}), {nbytes} B bf16, CPU tensors",
        "reference_real": {"warm_GBps": round(nbytes / warm / 1e9, 2), "cold_GBps": round(nbytes / cold / 1e9, 2),
                           "how": "unmodified reference via oracle/ref_harness.py (in-process fake actors, pickle round trips), "
                                  "SharedMemory transport, TORCHSTORE_PIN_SHM irrelevant without CUDA"},
        "port": {"GBps": round(port["value"], 2), "cores": port["cores"]},
        "port_over_reference": round(port["value"] / (nbytes / warm / 1e9), 2),
    }
    print(json.dumps(res, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
