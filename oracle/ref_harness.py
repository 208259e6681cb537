"""TEST INFRASTRUCTURE ONLY -- runs the *unmodified* reference (/root/reference) in-process.

This module exists so that ``oracle/gen_golden.py`` can execute the real torchstore code
in the build container (where Monarch, pygtrie and portpicker are not installed and there is
no network) and record its outputs as golden fixtures under ``tests/golden/``.

It is never imported by the product package, by ``bench.py`` or by any ``-m gpu`` test:
``/root/reference`` does not exist on the GPU box.

What it does (recipe from SURVEY.md Appendix B):
  * registers in-memory stub modules for ``monarch`` (Actor / endpoint / this_host /
    current_rank / get_or_spawn_controller / ProcMesh), ``pygtrie`` and ``portpicker``;
  * provides ``Shim``: a fake actor handle whose ``<endpoint>.call / call_one`` pickle
    round-trips the arguments and the result, reproducing the "client half / volume half
    of the same TransportBuffer" behaviour of real Monarch RPC;
  * provides ``make_store(num_volumes)``: Controller + StorageVolumes + per-rank LocalClient
    over the reference's SharedMemory transport.
"""

from __future__ import annotations

import collections
import os
import pickle
import sys
import types

REFERENCE_ROOT = os.environ.get("TORCHSTORE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "torchstore"))


# --------------------------------------------------------------------------------------
# stub modules
# --------------------------------------------------------------------------------------


class _StringTrie(collections.OrderedDict):
    """Minimal pygtrie.StringTrie stand-in: ordered dict + prefix iteration."""

    def __init__(self, *a, separator=".", **kw):
        super().__init__()
        self._sep = separator

    def iterkeys(self, prefix=None):
        if prefix is None:
            yield from list(self.keys())
            return
        hit = False
        for k in list(self.keys()):
            if k == prefix or k.startswith(prefix + self._sep):
                hit = True
                yield k
        if not hit:
            raise KeyError(prefix)


_current_rank = {"rank": 0}


def install_stubs() -> None:
    if "monarch" in sys.modules and getattr(sys.modules["monarch"], "_tsb200_stub", False):
        return

    monarch = types.ModuleType("monarch")
    monarch._tsb200_stub = True
    actor = types.ModuleType("monarch.actor")

    class Actor:  # noqa: D401 - stub
        pass

    def endpoint(fn):
        return fn

    class _Rank:
        @property
        def rank(self):
            return _current_rank["rank"]

    def current_rank():
        return _Rank()

    def this_host():
        raise RuntimeError("monarch stub: this_host() is not available in the oracle harness")

    async def get_or_spawn_controller(name, cls, *a, **kw):
        raise RuntimeError("monarch stub: get_or_spawn_controller is not available")

    class ProcMesh:
        pass

    actor.Actor = Actor
    actor.endpoint = endpoint
    actor.current_rank = current_rank
    actor.this_host = this_host
    actor.get_or_spawn_controller = get_or_spawn_controller
    actor.ProcMesh = ProcMesh
    monarch.actor = actor

    src = types.ModuleType("monarch._src")
    src_actor = types.ModuleType("monarch._src.actor")
    actor_mesh = types.ModuleType("monarch._src.actor.actor_mesh")
    actor_mesh._context = None
    src_actor.actor_mesh = actor_mesh
    src.actor = src_actor
    monarch._src = src

    pygtrie = types.ModuleType("pygtrie")
    pygtrie.StringTrie = _StringTrie
    pygtrie.Trie = _StringTrie

    portpicker = types.ModuleType("portpicker")

    def pick_unused_port():
        import socket

        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        return port

    portpicker.pick_unused_port = pick_unused_port

    sys.modules.update(
        {
            "monarch": monarch,
            "monarch.actor": actor,
            "monarch._src": src,
            "monarch._src.actor": src_actor,
            "monarch._src.actor.actor_mesh": actor_mesh,
            "pygtrie": pygtrie,
            "portpicker": portpicker,
        }
    )


def import_reference():
    """Import the reference package (read-only, unmodified). Returns the module."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    install_stubs()
    os.environ.setdefault("HYPERACTOR_CODEC_MAX_FRAME_LENGTH", "1")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import torchstore  # noqa: F401

    return torchstore


# --------------------------------------------------------------------------------------
# fake actor handles
# --------------------------------------------------------------------------------------


class ActorError(Exception):
    """Stand-in for monarch's ActorError: the remote exception arrives wrapped, with its type name
    in the message (the reference relies on this: client.py:485-496 greps for "KeyError")."""


class _Endpoint:
    def __init__(self, obj, name, pickle_roundtrip=True):
        self._obj = obj
        self._name = name
        self._rt = pickle_roundtrip

    async def _invoke(self, *args, **kwargs):
        if self._rt:
            args, kwargs = pickle.loads(pickle.dumps((args, kwargs)))
        try:
            out = await getattr(self._obj, self._name)(*args, **kwargs)
        except Exception as e:  # noqa: BLE001 - mimic the RPC boundary
            raise ActorError(f"{type(e).__name__}: {e}") from e
        if self._rt:
            out = pickle.loads(pickle.dumps(out))
        return out

    async def call_one(self, *args, **kwargs):
        return await self._invoke(*args, **kwargs)

    async def call(self, *args, **kwargs):
        return await self._invoke(*args, **kwargs)


class Shim:
    """``Shim(obj).<endpoint>.call_one(*a)`` == pickled ``await obj.<endpoint>(*a)``."""

    def __init__(self, obj, pickle_roundtrip=True):
        self._obj = obj
        self._rt = pickle_roundtrip

    def __getattr__(self, name):
        return _Endpoint(self._obj, name, self._rt)

    def __reduce__(self):
        # A pickled actor handle arrives as "some handle"; the volume half never uses it.
        return (type(None), ())


class RefStore:
    """Controller + N StorageVolumes + per-rank clients, all in this process."""

    def __init__(self, num_volumes: int, transport_name: str = "SharedMemory"):
        ts = import_reference()
        from torchstore.client import LocalClient
        from torchstore.controller import Controller
        from torchstore.storage_volume import StorageVolume
        from torchstore.strategy import LocalRankStrategy, StorageVolumeRef
        from torchstore.transport import TransportType
        from torchstore.utils import get_local_hostname

        self.ts = ts
        ttype = getattr(TransportType, transport_name)
        volumes = {}
        for r in range(num_volumes):
            volumes[str(r)] = StorageVolume(id_func=lambda r=r: str(r))
        self.volumes = volumes

        class _Strategy(LocalRankStrategy):
            def get_storage_volume(self_inner, volume_id):
                return StorageVolumeRef(
                    Shim(volumes[volume_id]),
                    volume_id,
                    self_inner.transport_context,
                    self_inner.default_transport_type,
                    volume_hostname=get_local_hostname(),
                )

        self._strategy_cls = _Strategy
        self._ttype = ttype
        self.controller = Controller()
        self.controller.is_initialized = True
        self._controller_strategy = _Strategy(ttype)
        self._controller_strategy.volume_id_to_coord = {v: {"gpus": int(v)} for v in volumes}
        self.controller.strategy = self._controller_strategy
        self._clients = {}
        self.LocalClient = LocalClient

    def client(self, rank: int):
        """A LocalClient that behaves as the process with LOCAL_RANK=rank."""
        if rank not in self._clients:
            strat = self._strategy_cls(self._ttype)
            strat.volume_id_to_coord = {v: {"gpus": int(v)} for v in self.volumes}
            self._clients[rank] = self.LocalClient(Shim(self.controller), strat)
        os.environ.pop("RANK", None)
        os.environ["LOCAL_RANK"] = str(rank)
        return self._clients[rank]

    def close(self):
        for c in self._clients.values():
            c.strategy.transport_context.clear()
        for v in self.volumes.values():
            v.store.reset()
