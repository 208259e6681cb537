"""World-size-2 gloo test of the all-gather route (the only collective on the path): shape detection
and the gather itself, on CPU tensors.  On GPUs the same calls run over NCCL."""

import os
import socket
import tempfile

import numpy as np
import torch
import torch.multiprocessing as mp

from torchstore_b200.collectives import is_allgather_shaped
from torchstore_b200.transport.types import TensorSlice


def test_shape_detection():
    world = 4
    shapes = {"w": (16, 8), "b": (8,)}

    def ts(name, rank, rows=None):
        g = shapes[name]
        r = g[0] // world if rows is None else rows
        return TensorSlice((rank * r,) + (0,) * (len(g) - 1), (rank,), g, (r,) + g[1:], (world,))

    ok = {n: ts(n, 2) for n in shapes}
    assert is_allgather_shaped(ok, shapes, rank=2, world=4)
    assert not is_allgather_shaped(ok, shapes, rank=1, world=4)           # not my shard
    assert not is_allgather_shaped(ok, {"w": (16, 8)}, rank=2, world=4)     # different key sets
    assert not is_allgather_shaped(ok, {"w": (16, 8), "b": (2,)}, 2, 4)    # dest is not the full tensor
    col = dict(ok, w=TensorSlice((0, 4), (2,), (16, 8), (16, 2), (4,)))     # column shard
    assert not is_allgather_shaped(col, shapes, 2, 4)
    uneven = {"w": TensorSlice((12,), (2,), (18,), (6,), (3,))}
    assert is_allgather_shaped(uneven, {"w": (18,)}, 2, 3)
    assert not is_allgather_shaped({"w": TensorSlice((14,), (2,), (19,), (5,), (3,))}, {"w": (19,)}, 2, 3)


def _worker(rank, world, port, outdir):
    import torch.distributed as dist

    from torchstore_b200.collectives import all_gather_state_dict

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    full = {"w": torch.arange(8 * 6, dtype=torch.float32).reshape(8, 6), "b": torch.arange(10, dtype=torch.int64)}
    shards = {k: v.chunk(world, dim=0)[rank].contiguous() for k, v in full.items()}
    dests = {k: torch.zeros_like(v) for k, v in full.items()}
    n = all_gather_state_dict(shards, dests)
    ok = n == 2 and all(torch.equal(dests[k], full[k]) for k in full)
    try:
        all_gather_state_dict({"w": shards["w"]}, {"w": torch.zeros(3, 3)})
        ok = False
    except ValueError:
        pass
    dist.destroy_process_group()
    np.save(os.path.join(outdir, f"{rank}.npy"), np.array([ok]))


def test_all_gather_state_dict_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d), nprocs=2, join=True)
        assert all(np.load(os.path.join(d, f"{r}.npy")).all() for r in range(2))
