"""torchrun-style bootstrap without Monarch (reference torchstore/spmd.py:43-362).

Every rank calls ``await initialize(strategy)``.  Each rank starts one ``ActorServer`` thread and
hosts its own storage volume (LocalRankStrategy: volume id == rank, tensors in that rank's GPU);
rank 0 also hosts the controller.  Handles are exchanged through a ``TCPStore`` rendezvous on
MASTER_ADDR:MASTER_PORT (the store torchrun's agent already serves, when there is one).
"""

from __future__ import annotations

import logging
import os
import pickle
import socket
from dataclasses import dataclass
from datetime import timedelta
from typing import Any

import torch
from torch.distributed import TCPStore

import torchstore_b200.api as _api
from torchstore_b200 import rpc
from torchstore_b200 import strategy as strategy_mod
from torchstore_b200.controller import Controller
from torchstore_b200.storage_volume import StorageVolume
from torchstore_b200.strategy import HostStrategy, LocalRankStrategy, TorchStoreStrategy

logger = logging.getLogger(__name__)


def _spmd_key(store_name: str, suffix: str, generation: int | None = None) -> str:
    gen = "" if generation is None else f"g{generation}/"
    return f"torchstore/spmd/{store_name}/{gen}{suffix}"


@dataclass(frozen=True)
class SPMDEnv:
    rank: int
    local_rank: int
    world_size: int
    local_world_size: int
    master_addr: str
    master_port: int

    @property
    def num_hosts(self) -> int:
        return self.world_size // self.local_world_size

    @property
    def group_rank(self) -> int:
        return self.rank // self.local_world_size

    @staticmethod
    def _parse(name: str, default: str | None = None) -> str:
        value = os.environ.get(name, default)
        if value is None:
            raise RuntimeError(f"SPMD TorchStore initialization requires the {name} env var")
        return value

    @classmethod
    def from_env(cls, *, master_addr: str | None = None, master_port: int | None = None) -> "SPMDEnv":
        rank = int(cls._parse("RANK"))
        local_rank = int(cls._parse("LOCAL_RANK"))
        world_size = int(cls._parse("WORLD_SIZE"))
        local_world_size = int(cls._parse("LOCAL_WORLD_SIZE", str(world_size)))
        if world_size % local_world_size != 0:
            raise ValueError(f"world_size ({world_size}) must be divisible by local_world_size ({local_world_size})")
        return cls(rank=rank, local_rank=local_rank, world_size=world_size, local_world_size=local_world_size,
                   master_addr=cls._parse("MASTER_ADDR", master_addr),
                   master_port=int(cls._parse("MASTER_PORT", None if master_port is None else str(master_port))))


class _SPMDSession:
    """Per-process resources of an SPMD store; ``rendezvous`` is reusable by callers."""

    def __init__(self, *, rendezvous, controller, store_name: str, is_primary: bool, env: SPMDEnv,
                 owned_actors: list[str], generation: int = 1) -> None:
        self.generation = generation
        self.rendezvous = rendezvous
        self.controller = controller
        self.is_primary = is_primary
        self.env = env
        self._store_name = store_name
        self._owned = owned_actors

    async def shutdown(self) -> None:
        """Coordinated teardown: every rank checks in, rank 0 tears the controller down and
        publishes the outcome, the others wait for it.  Safe to call twice."""
        if _api._spmd_state_map.pop(self._store_name, None) is None:
            return
        name = self._store_name
        gen = self.generation
        err: Exception | None = None
        try:
            self.rendezvous.add(_spmd_key(name, "shutdown_arrivals", gen), 1)
            if self.is_primary:
                status = "ok"
                try:
                    # wait until every rank stopped issuing requests
                    import time

                    deadline = time.time() + 120
                    while self.rendezvous.add(_spmd_key(name, "shutdown_arrivals", gen), 0) < self.env.world_size:
                        if time.time() > deadline:
                            raise RuntimeError("Timed out waiting for all ranks to reach shutdown")
                        time.sleep(0.002)
                    await self.controller.teardown.call()
                except Exception as e:  # noqa: BLE001
                    err, status = e, repr(e)
                self.rendezvous.set(_spmd_key(name, "shutdown", gen), status)
                # the primary may own the rendezvous server: keep it alive until every other rank
                # has read the outcome (otherwise their pending get dies with "connection reset")
                import time

                deadline = time.time() + 30
                while self.rendezvous.add(_spmd_key(name, "shutdown_acks", gen), 0) < self.env.world_size - 1:
                    if time.time() > deadline:
                        logger.warning("TorchStore shutdown: not every rank acknowledged within 30 s")
                        break
                    time.sleep(0.002)
            else:
                try:
                    status = self.rendezvous.get(_spmd_key(name, "shutdown", gen)).decode()
                    self.rendezvous.add(_spmd_key(name, "shutdown_acks", gen), 1)
                except Exception as e:
                    raise RuntimeError("Timed out waiting for TorchStore shutdown") from e
                if status != "ok":
                    raise RuntimeError(f"TorchStore SPMD shutdown failure - '{name}': {status}")
        finally:
            from torchstore_b200 import state_dict_utils

            cl = _api._local_clent_map.get(name)
            if cl is not None:
                state_dict_utils.reset_direct_cache(cl)
                cl.close_sessions()
                cl.strategy.transport_context.clear()
            _api.reset_client(name)
            for actor in self._owned:
                rpc.unregister_actor(actor)
        if err is not None:
            raise err


def _validate_strategy(strategy) -> HostStrategy | LocalRankStrategy:
    if isinstance(strategy, (HostStrategy, LocalRankStrategy)):
        return strategy
    raise RuntimeError("SPMD mode requires an explicit HostStrategy or LocalRankStrategy")


_rendezvous_cache: dict[tuple, Any] = {}


def _open_rendezvous(env: SPMDEnv, timeout: timedelta):
    """TCPStore on MASTER_ADDR:MASTER_PORT, created once per process and kept for its lifetime: a
    store that is shut down and re-initialised in the same job must not race a dying server against
    a new one on the same port."""
    key = (env.master_addr, env.master_port, env.rank, env.world_size)
    store = _rendezvous_cache.get(key)
    if store is None:
        # under torchrun the elastic agent already serves a TCPStore on MASTER_PORT: join it as a client
        agent_store = os.environ.get("TORCHELASTIC_USE_AGENT_STORE", "False") == "True"
        store = TCPStore(env.master_addr, env.master_port, env.world_size,
                         is_master=(env.rank == 0 and not agent_store), timeout=timeout, wait_for_workers=False)
        _rendezvous_cache[key] = store
    return store


async def initialize(strategy: TorchStoreStrategy | None = None, store_name: str = _api.DEFAULT_TORCHSTORE_NAME, *,
                     env: SPMDEnv | None = None, rendezvous_timeout: timedelta = timedelta(seconds=120),
                     transport: str = "ipc", monarch_port: int = 26600, rendezvous=None) -> None:
    """Initialize TorchStore from RANK / LOCAL_RANK / WORLD_SIZE / LOCAL_WORLD_SIZE / MASTER_ADDR /
    MASTER_PORT.  One volume per rank (LocalRankStrategy) or per host (HostStrategy).

    ``transport`` / ``monarch_port`` are accepted for signature compatibility (the control plane is
    localhost sockets on one box).  ``rendezvous`` may pass an existing c10d store."""
    import time

    t_begin = time.perf_counter()

    def trace(what: str) -> None:
        if os.environ.get("TSB_TRACE_INIT"):
            import sys

            print(f"[spmd r{os.environ.get('RANK', '?')} +{time.perf_counter() - t_begin:6.2f}s] {what}", file=sys.stderr, flush=True)

    strategy = _validate_strategy(strategy)
    if store_name in _api._spmd_state_map:
        raise RuntimeError(f"TorchStore '{store_name}' is already initialized")
    if env is None:
        env = SPMDEnv.from_env()
    os.environ.setdefault("HOSTNAME", socket.gethostname())
    if rendezvous is None:
        rendezvous = _open_rendezvous(env, rendezvous_timeout)

    # k-th initialisation of this store name in this job: every rank bumps its own counter, so all
    # ranks derive the same generation and never read keys of an earlier incarnation
    gen = int(rendezvous.add(_spmd_key(store_name, f"generation/{env.rank}"), 1))
    trace("rendezvous open")
    rpc.ActorServer.instance()  # start this process' actor server (idempotent)
    trace("actor server up")
    owned: list[str] = []
    hosts_volume = isinstance(strategy, LocalRankStrategy) or env.local_rank == 0
    if hosts_volume:
        strategy_mod._spawn_rank[0] = env.rank if isinstance(strategy, LocalRankStrategy) else env.group_rank
        device = env.local_rank if torch.cuda.is_available() else None
        vol = StorageVolume(id_func=strategy.get_volume_id, device=device)
        strategy_mod._spawn_rank[0] = 0
        vol_name = f"{store_name}/g{gen}/volume/{env.rank}"
        owned.append(vol_name)
        vol_ref = rpc.register_actor(vol_name, vol)
        rendezvous.set(_spmd_key(store_name, f"volume/{env.rank}", gen), pickle.dumps(vol_ref))
        trace("volume registered")
    else:
        rendezvous.set(_spmd_key(store_name, f"volume/{env.rank}", gen), pickle.dumps(None))

    controller_key = _spmd_key(store_name, "controller", gen)
    if env.rank == 0:
        members = []
        for r in range(env.world_size):
            ref = pickle.loads(rendezvous.get(_spmd_key(store_name, f"volume/{r}", gen)))
            if ref is not None:
                members.append(({"gpus": r}, ref))
        name = f"{store_name}/g{gen}/controller"
        owned.append(name)
        trace("all volume refs collected")
        controller = rpc.register_actor(name, Controller())
        await controller.init.call(strategy=strategy, num_storage_volumes=len(members),
                                   storage_volumes=rpc.ActorMesh(members))
        trace("controller initialised")
        rendezvous.set(controller_key, pickle.dumps(controller))
    else:
        controller = pickle.loads(rendezvous.get(controller_key))
        trace("controller ref received")
    _api._spmd_state_map[store_name] = _SPMDSession(rendezvous=rendezvous, controller=controller, store_name=store_name,
                                                   is_primary=(env.rank == 0), env=env, owned_actors=owned, generation=gen)


__all__ = ["SPMDEnv", "initialize"]
