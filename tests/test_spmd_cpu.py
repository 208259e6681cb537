"""N>1 control plane on CPU: two processes bootstrap an SPMD store over a TCPStore rendezvous
(gloo-style, no GPU), each hosting its own volume; objects put on one rank are visible on the
other; shutdown is coordinated (reference tests/test_spmd.py:251-374)."""

import asyncio
import json
import os
import socket
import tempfile

import pytest
import torch.multiprocessing as mp

from torchstore_b200.spmd import SPMDEnv


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


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


def _worker(rank, world, port, pg_port, outdir):
    os.environ.update({"RANK": str(rank), "LOCAL_RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    os.environ.pop("TORCHELASTIC_USE_AGENT_STORE", None)
    import torch.distributed as dist

    import torchstore_b200 as ts

    async def main():
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{pg_port}", rank=rank, world_size=world)
        await ts.initialize_spmd(ts.LocalRankStrategy())
        res = {}
        await ts.put(f"from_{rank}", {"rank": rank, "payload": list(range(rank + 3))})
        dist.barrier()
        other = (rank + 1) % world
        res["peer"] = await ts.get(f"from_{other}")
        res["keys"] = sorted(await ts.keys())
        c = await ts.client()
        vm = await c._controller.locate_volumes.call_one([f"from_{rank}"])
        res["my_volume"] = list(vm[f"from_{rank}"].keys())
        # state dict of objects written by rank 0, read by rank 1
        if rank == 0:
            await ts.put_state_dict({"step": 11, "cfg": {"a": 1}}, "sd")
        dist.barrier()
        res["sd"] = await ts.get_state_dict("sd")
        res["exists_missing"] = await ts.exists("nope")
        dist.barrier()
        # host tier across processes (no GPU here): every rank owns rows [rank*R, (rank+1)*R) of a
        # [world*R, C] fp32 weight (FSDP Shard(0)) in ITS volume's POSIX shm segments; each rank then
        # reads its TP Shard(1) columns in place from all volumes, and the whole tensor
        import torch

        from torchstore_b200.transport import create_transport_buffer
        from torchstore_b200.transport.types import Request, TensorSlice

        R, C = 64, 96
        full = torch.arange(world * R * C, dtype=torch.float32).reshape(world * R, C)
        req = Request.from_any("w", full[rank * R:(rank + 1) * R].contiguous(),
                               TensorSlice((rank * R, 0), (rank,), (world * R, C), (R, C), (world,)))
        ref = c.strategy.select_storage_volume()
        await create_transport_buffer(ref).put_to_storage_volume([req])
        await c._controller.notify_put_batch.call([req.meta_only()], ref.volume_id)
        dist.barrier()
        cw = C // world
        dest = torch.zeros(world * R, cw)
        got = await ts.get("w", dest, TensorSlice((0, rank * cw), (rank,), (world * R, C), (world * R, cw), (world,)))
        res["reshard_ok"] = bool(got is dest and torch.equal(dest, full[:, rank * cw:(rank + 1) * cw]))
        res["full_ok"] = bool(torch.equal(await ts.get("w"), full))
        dist.barrier()
        # real DTensors on a CPU mesh (reference tests/test_tensor_slice.py:150-328, :400-506): put a Shard(0)
        # DTensor from every rank, read rectangles that span volumes, the whole tensor, and the same key back
        # under another placement in place; an all-Replicate DTensor is stored as a plain tensor
        from torch.distributed.device_mesh import init_device_mesh
        from torch.distributed.tensor import DTensor, Replicate, Shard, distribute_tensor

        from torchstore_b200.controller import ObjectType

        mesh = init_device_mesh("cpu", (world,))
        cols = 6 if world == 2 else 2 * world  # every rank must own a column under Shard(1) (an empty shard has nothing to fetch)
        original = torch.arange(4 * world * cols, dtype=torch.float32).reshape(4 * world, cols)
        await ts.put("dt", distribute_tensor(original, mesh, [Shard(0)]))
        dist.barrier()
        cross = TensorSlice((2, 1), (), tuple(original.shape), (4, 4), ())     # spans the volume boundary at row 4
        single = TensorSlice((1, 0), (), tuple(original.shape), (2, 3), ())    # inside volume 0
        ok = torch.equal(await ts.get("dt", tensor_slice_spec=cross), original[2:6, 1:5])
        ok = ok and torch.equal(await ts.get("dt", tensor_slice_spec=single), original[1:3, 0:3])
        ok = ok and torch.equal(await ts.get("dt"), original)
        dest_dt = distribute_tensor(torch.zeros_like(original), mesh, [Shard(1)])
        out = await ts.get("dt", dest_dt)
        ok = ok and out is dest_dt and torch.equal(dest_dt.to_local(), original.chunk(world, dim=1)[rank])
        expert = torch.full((16, 8), float(rank))
        await ts.put(f"expert_{rank}.weight", DTensor.from_local(expert, mesh, [Replicate()], run_check=False))
        dist.barrier()
        peer_expert = await ts.get(f"expert_{other}.weight")
        ok = ok and torch.equal(peer_expert, torch.full((16, 8), float(other)))
        info = (await c._controller.locate_volumes.call_one([f"expert_{rank}.weight"]))[f"expert_{rank}.weight"]
        ok = ok and all(v.object_type == ObjectType.TENSOR for v in info.values())
        res["dtensor_ok"] = bool(ok)
        dist.barrier()
        await ts.shutdown()
        # the same job can bring the store up again: nothing of the first incarnation leaks in
        await ts.initialize_spmd(ts.LocalRankStrategy())
        res["second_keys"] = await ts.keys()
        dist.barrier()
        await ts.put(f"again_{rank}", rank)
        dist.barrier()
        res["second_peer"] = await ts.get(f"again_{(rank + 1) % world}")
        dist.barrier()
        await ts.shutdown()
        dist.destroy_process_group()
        with open(os.path.join(outdir, f"{rank}.json"), "w") as f:
            json.dump(res, f)

    asyncio.run(main())


@pytest.mark.parametrize("world", [2, 4])
def test_spmd_store_on_cpu(world):
    port, pg_port = _free_ports(2)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, port, pg_port, d), nprocs=world, join=True)
        out = [json.load(open(os.path.join(d, f"{r}.json"))) for r in range(world)]
    for r in range(world):
        other = (r + 1) % world
        assert out[r]["peer"] == {"rank": other, "payload": list(range(other + 3))}
        assert out[r]["keys"][:world] == [f"from_{i}" for i in range(world)]
        assert out[r]["my_volume"] == [str(r)]
        assert out[r]["sd"] == {"step": 11, "cfg": {"a": 1}}
        assert out[r]["exists_missing"] is False
        assert out[r]["reshard_ok"] and out[r]["full_ok"] and out[r]["dtensor_ok"]
        assert out[r]["second_keys"] == [] and out[r]["second_peer"] == other


def test_spmd_env_parsing(monkeypatch):
    for k, v in {"RANK": "3", "LOCAL_RANK": "1", "WORLD_SIZE": "8", "LOCAL_WORLD_SIZE": "4", "MASTER_ADDR": "h",
                 "MASTER_PORT": "1234"}.items():
        monkeypatch.setenv(k, v)
    env = SPMDEnv.from_env()
    assert (env.rank, env.local_rank, env.world_size, env.num_hosts, env.group_rank) == (3, 1, 8, 2, 0)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "3")
    with pytest.raises(ValueError):
        SPMDEnv.from_env()
    monkeypatch.delenv("RANK")
    with pytest.raises(RuntimeError, match="requires the RANK env var"):
        SPMDEnv.from_env()


def test_spmd_requires_explicit_strategy():
    import torchstore_b200 as ts

    with pytest.raises(RuntimeError, match="explicit HostStrategy or LocalRankStrategy"):
        asyncio.run(ts.initialize_spmd(None))
