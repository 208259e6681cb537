"""Multi-GPU paths: P2P between two devices of one process, CUDA-IPC between two processes, and a
2-rank SPMD store with direct weight sync.  Skipped on boxes with a single GPU."""

import asyncio
import os
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")


def run(coro):
    return asyncio.run(coro)


@needs2
def test_p2p_pull_between_two_devices_single_process():
    from torchstore_b200.direct_weight_sync import DirectWeightSyncDest, DirectWeightSyncSource

    src = torch.arange(1 << 22, dtype=torch.int32, device="cuda:0").reshape(2048, 2048)
    handles = DirectWeightSyncSource().register({"w": src}, rank=0)
    dst = torch.zeros(2048, 2048, dtype=torch.int32, device="cuda:1")
    sync = DirectWeightSyncDest()
    run(sync.pull({"w": [handles["w"]]}, {"w": dst}))
    assert torch.equal(dst.cpu(), src.cpu())
    info = sync.plan_info()[1]
    assert info["remote_src_bytes"] == src.numel() * 4


@needs2
def test_store_get_across_devices_single_process(monkeypatch):
    import torchstore_b200 as ts

    async def main():
        await ts.initialize(num_storage_volumes=2, strategy=ts.LocalRankStrategy())
        try:
            monkeypatch.delenv("RANK", raising=False)
            monkeypatch.setenv("LOCAL_RANK", "1")
            t = torch.randn(1024, 512, device="cuda:1")
            await ts.put("t", t)  # lands in volume 1 = HBM of GPU 1
            monkeypatch.setenv("LOCAL_RANK", "0")
            dest = torch.zeros(1024, 512, device="cuda:0")
            await ts.get("t", dest)  # GPU 0 pulls over NVLink
            assert torch.equal(dest.cpu(), t.cpu())
        finally:
            await ts.shutdown()

    run(main())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _free_ports(n):
    """n distinct free ports (all sockets held open until every port is chosen)."""
    socks = []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        socks.append(s)
    ports = [s.getsockname()[1] for s in socks]
    for s in socks:
        s.close()
    return ports


def _ipc_worker(rank, world, port, pg_port, outdir, same_gpu):
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(0 if same_gpu else rank), "WORLD_SIZE": str(world),
                       "LOCAL_WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)
    import torch.distributed as dist

    import torchstore_b200 as ts

    dev = torch.device("cuda", 0 if same_gpu else rank)
    torch.cuda.set_device(dev)

    async def main():
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{pg_port}", rank=rank, world_size=world)
        await ts.initialize_spmd(ts.LocalRankStrategy())
        # every rank owns rows [rank*R, (rank+1)*R) of a [world*R, C] weight (FSDP Shard(0)) and
        # wants columns [rank*C/world, ...) of all rows (TP Shard(1))
        R, C = 256, 512
        full = torch.arange(world * R * C, dtype=torch.float32).reshape(world * R, C)
        mine = full[rank * R:(rank + 1) * R].contiguous().to(dev)
        from torchstore_b200.transport.types import TensorSlice

        src_slice = TensorSlice((rank * R, 0), (rank,), (world * R, C), (R, C), (world,))
        from torchstore_b200.direct_weight_sync import DirectWeightSyncDest, DirectWeightSyncSource

        source = DirectWeightSyncSource()
        handles = source.register({"w": mine}, rank=rank, tensor_slices={"w": src_slice})
        await ts.put(f"sync/rank_{rank}", handles)
        dist.barrier()
        all_handles = {"w": []}
        for r in range(world):
            all_handles["w"].append((await ts.get(f"sync/rank_{r}"))["w"])
        cw = C // world
        dest = torch.zeros(world * R, cw, device=dev)
        dslice = TensorSlice((0, rank * cw), (rank,), (world * R, C), (world * R, cw), (world,))
        sync = DirectWeightSyncDest()
        await sync.pull(all_handles, {"w": dest}, {"w": dslice})
        ok1 = torch.equal(dest.cpu(), full[:, rank * cw:(rank + 1) * cw])
        # second round after an in-place update on every source
        dist.barrier()
        mine.mul_(-1.0)
        torch.cuda.synchronize()
        dist.barrier()
        await sync.pull(all_handles, {"w": dest}, {"w": dslice})
        ok2 = torch.equal(dest.cpu(), -full[:, rank * cw:(rank + 1) * cw])
        # store path across processes: tensor put by rank r, read in place by the other rank
        await ts.put(f"t_{rank}", torch.full((128, 64), float(rank + 1), device=dev))
        dist.barrier()
        other = (rank + 1) % world
        got = await ts.get(f"t_{other}", torch.zeros(128, 64, device=dev))
        ok3 = bool((got == float(other + 1)).all().item())
        info = sync.plan_info()[dev.index]
        dist.barrier()
        sync.close()
        await ts.shutdown()
        dist.destroy_process_group()
        np.save(os.path.join(outdir, f"{rank}.npy"), np.array([ok1, ok2, ok3, info["remote_src_bytes"] > 0 or same_gpu]))

    asyncio.run(main())


@pytest.mark.parametrize("same_gpu", [True, False])
def test_two_process_ipc_direct_sync_and_store(same_gpu):
    """Cross-process handles: cudaIpc mapping of torch-allocated params (base handle + offset)."""
    if not same_gpu and _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    port, pg_port = _free_ports(2)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_ipc_worker, args=(world, port, pg_port, d, same_gpu), nprocs=world, join=True)
        for r in range(world):
            assert np.load(os.path.join(d, f"{r}.npy")).all(), (r, np.load(os.path.join(d, f"{r}.npy")))


@needs2
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.int32])
def test_config2_flat_4gib_between_two_gpus(dtype):
    """BASELINE config #2: direct_weight_sync of a 4 GiB flat tensor GPU0 -> GPU1, same placement
    (one exact-match op); bf16 random weights and the int32 arange variant, bit-exact."""
    from torchstore_b200.direct_weight_sync import DirectWeightSyncDest, DirectWeightSyncSource

    nbytes = 4 << 30
    n = nbytes // dtype.itemsize
    if dtype == torch.int32:
        src = torch.arange(n, dtype=torch.int32, device="cuda:0")
    else:
        src = torch.empty(n, dtype=dtype, device="cuda:0").normal_(0, 0.02)
    handles = DirectWeightSyncSource().register({"flat": src}, rank=0)
    dst = torch.zeros(n, dtype=dtype, device="cuda:1")
    sync = DirectWeightSyncDest()
    run(sync.pull({"flat": [handles["flat"]]}, {"flat": dst}))
    assert len(sync._plan) == 1 and sync._plan[0].dest_tensor is None
    a = src.view(torch.int32) if dtype == torch.int32 else src.view(torch.int16)
    b = dst.view(torch.int32) if dtype == torch.int32 else dst.view(torch.int16)
    assert int(a.to(torch.int64).sum()) == int(b.to(torch.int64).sum())  # checksum at full size
    step = n // 64
    for i in range(0, n, step):  # spot-check 64 windows bit for bit
        assert torch.equal(src[i:i + 4096].cpu(), dst[i:i + 4096].cpu())
    gbps = nbytes / sync.last_pull_ms[1] / 1e6
    print(f"config2 {dtype}: {sync.last_pull_ms[1]:.3f} ms, {gbps:.1f} GB/s")
    sync.close()
