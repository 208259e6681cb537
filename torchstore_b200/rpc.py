"""Minimal actor plumbing for a single NVSwitch box (the reference runs on Monarch, which is a
cluster-scale Rust actor system; on one box the control plane only carries small pickled
metadata, so a thread + a localhost socket per process is enough).

The handle API mirrors what the reference calls on Monarch handles, so client / transport /
controller code reads the same (reference client.py:47,86,247; strategy.py:107,137):

    ref.<endpoint>.call_one(*args)      one actor, returns its result
    ref.<endpoint>.call(*args)          single ref: same as call_one; mesh: list of (coord, result)
    mesh.slice(**coord)                 one member of an actor mesh

* ``LocalActorRef``  -- the actor object lives in this process: direct await (optionally with a
  pickle round-trip, which is what makes a TransportBuffer arrive as "the volume half").
* ``RemoteActorRef`` -- the actor lives in another process of the box: length-prefixed pickles over
  a ``multiprocessing.connection`` socket served by that process's ``ActorServer`` thread.
  Refs pickle to (address, authkey, name) and resolve back to a LocalActorRef inside the owner.

Errors cross the boundary as ``ActorError`` whose message starts with the original exception's
type name -- the reference relies on that for ``exists()`` (client.py:485-496).
"""

from __future__ import annotations

import asyncio
import os
import pickle
import threading
from multiprocessing.connection import Client, Listener
from typing import Any

PICKLE_LOCAL = os.environ.get("TORCHSTORE_B200_PICKLE_LOCAL_RPC", "1") == "1"


class ActorError(RuntimeError):
    """An exception raised inside an endpoint, as seen by the caller."""


class Actor:
    """Marker base class (monarch.actor.Actor stand-in)."""


def endpoint(fn):
    """Marks a coroutine method as remotely callable (monarch.actor.endpoint stand-in)."""
    fn._tsb_endpoint = True
    return fn


# process-local registry: actor name -> (object, mailbox lock)
_registry: dict[str, tuple[Any, threading.Lock]] = {}
_registry_lock = threading.Lock()


def register_actor(name: str, obj: Any) -> "LocalActorRef":
    with _registry_lock:
        _registry[name] = (obj, threading.Lock())
    return LocalActorRef(name)


def unregister_actor(name: str) -> None:
    with _registry_lock:
        _registry.pop(name, None)


def _lookup(name: str):
    with _registry_lock:
        ent = _registry.get(name)
    if ent is None:
        raise ActorError(f"KeyError: no actor named {name!r} in pid {os.getpid()}")
    return ent


async def _invoke(name: str, method: str, args, kwargs):
    obj, lock = _lookup(name)
    fn = getattr(obj, method)
    if not getattr(fn, "_tsb_endpoint", False):
        raise ActorError(f"AttributeError: {type(obj).__name__}.{method} is not an endpoint")
    # One endpoint at a time per actor (Monarch's mailbox semantics), for callers on ANY thread or
    # event loop: a plain non-reentrant lock, acquired by yielding to the loop instead of blocking it,
    # so a coroutine of the same loop that holds the mailbox can finish and other actors keep being
    # served meanwhile.  (An actor must not call its own endpoints through a ref.)
    spins = 0
    while not lock.acquire(blocking=False):
        spins += 1
        await asyncio.sleep(0 if spins < 200 else 0.0005)
    try:
        return await fn(*args, **kwargs)
    finally:
        lock.release()


class _LocalEndpoint:
    def __init__(self, name: str, method: str):
        self._name, self._method = name, method

    async def call_one(self, *args, **kwargs):
        if PICKLE_LOCAL:
            args, kwargs = pickle.loads(pickle.dumps((args, kwargs)))
        try:
            out = await _invoke(self._name, self._method, args, kwargs)
        except ActorError:
            raise
        except Exception as e:  # noqa: BLE001 -- the RPC boundary
            raise ActorError(f"{type(e).__name__}: {e}") from e
        if PICKLE_LOCAL:
            out = pickle.loads(pickle.dumps(out))
        return out

    call = call_one


class LocalActorRef:
    def __init__(self, name: str):
        self._name = name

    def __getattr__(self, method: str):
        if method.startswith("_"):
            raise AttributeError(method)
        return _LocalEndpoint(self._name, method)

    def __reduce__(self):
        server = ActorServer.instance(create=False)
        if server is None:
            # single-process store: the handle only ever comes back to this process
            return (LocalActorRef, (self._name,))
        return (_resolve_ref, (server.address, server.authkey, self._name, os.getpid()))


def _resolve_ref(address, authkey, name, pid):
    """Unpickle a handle: inside the owning process it becomes a direct (in-thread) reference."""
    if pid == os.getpid():
        return LocalActorRef(name)
    server = ActorServer.instance(create=False)
    if server is not None and tuple(server.address) == tuple(address) and server.authkey == authkey:
        return LocalActorRef(name)  # a handle that travelled through another process and came home
    return RemoteActorRef(address, authkey, name)


# ---------------------------------------------------------------------------------------------
# remote side
# ---------------------------------------------------------------------------------------------
class _Connections:
    """Per-process cache of client connections, one per server address (thread-safe)."""

    def __init__(self):
        self._conns: dict[tuple, tuple[Any, threading.Lock]] = {}
        self._lock = threading.Lock()

    def get(self, address, authkey):
        key = (tuple(address), authkey)
        with self._lock:
            ent = self._conns.get(key)
            if ent is None:
                ent = (Client(tuple(address), authkey=authkey), threading.Lock())
                self._conns[key] = ent
        return ent

    def close(self):
        with self._lock:
            for conn, _ in self._conns.values():
                try:
                    conn.close()
                except Exception:
                    pass
            self._conns.clear()


_connections = _Connections()


class _RemoteEndpoint:
    def __init__(self, ref: "RemoteActorRef", method: str):
        self._ref, self._method = ref, method

    async def call_one(self, *args, **kwargs):
        conn, lock = _connections.get(self._ref._address, self._ref._authkey)
        payload = pickle.dumps((self._ref._name, self._method, args, kwargs), protocol=pickle.HIGHEST_PROTOCOL)

        def roundtrip():
            with lock:  # one request/reply in flight per connection
                conn.send_bytes(payload)
                return conn.recv_bytes()

        # blocking socket I/O runs on a worker thread: the caller's loop stays live, so
        # asyncio.gather over several volumes really is concurrent (reference client.py:329-331)
        ok, out = pickle.loads(await asyncio.get_running_loop().run_in_executor(None, roundtrip))
        if not ok:
            raise ActorError(out)
        return out

    call = call_one


class RemoteActorRef:
    def __init__(self, address, authkey: bytes, name: str):
        self._address, self._authkey, self._name = tuple(address), authkey, name

    def __getattr__(self, method: str):
        if method.startswith("_"):
            raise AttributeError(method)
        return _RemoteEndpoint(self, method)

    def __reduce__(self):
        return (_resolve_ref, (self._address, self._authkey, self._name, -1))


class ActorServer:
    """One per process: accepts connections on 127.0.0.1 and runs endpoint coroutines on its own
    event loop thread."""

    _instance: "ActorServer | None" = None
    _instance_lock = threading.Lock()

    def __init__(self, host: str = "127.0.0.1"):
        self.authkey = os.urandom(16)
        # backlog: every rank of the box may connect at the same instant (multiprocessing's default
        # of 1 overflows the accept queue at 8 ranks and a dropped handshake hangs the client forever)
        self._listener = Listener((host, 0), backlog=512, authkey=self.authkey)
        self.address = self._listener.address
        self._loop = asyncio.new_event_loop()
        self._closed = False
        self._threads: list[threading.Thread] = []
        self._loop_thread = threading.Thread(target=self._run_loop, name="tsb200-actor-loop", daemon=True)
        self._loop_thread.start()
        self._accept_thread = threading.Thread(target=self._accept, name="tsb200-actor-accept", daemon=True)
        self._accept_thread.start()

    @classmethod
    def instance(cls, create: bool = True) -> "ActorServer | None":
        with cls._instance_lock:
            if cls._instance is None and create:
                cls._instance = ActorServer()
            return cls._instance

    def _run_loop(self):
        asyncio.set_event_loop(self._loop)
        self._loop.run_forever()

    def _accept(self):
        while not self._closed:
            try:
                conn = self._listener.accept()
            except Exception:
                if self._closed:
                    return
                continue
            t = threading.Thread(target=self._serve, args=(conn,), name="tsb200-actor-conn", daemon=True)
            t.start()
            self._threads.append(t)

    def _serve(self, conn):
        while not self._closed:
            try:
                raw = conn.recv_bytes()
            except (EOFError, OSError):
                return
            try:
                name, method, args, kwargs = pickle.loads(raw)
                fut = asyncio.run_coroutine_threadsafe(_invoke(name, method, args, kwargs), self._loop)
                reply = (True, fut.result())
            except ActorError as e:
                reply = (False, str(e))
            except Exception as e:  # noqa: BLE001
                reply = (False, f"{type(e).__name__}: {e}")
            try:
                conn.send_bytes(pickle.dumps(reply, protocol=pickle.HIGHEST_PROTOCOL))
            except Exception as e:  # result not picklable, or peer gone
                try:
                    conn.send_bytes(pickle.dumps((False, f"{type(e).__name__}: {e}")))
                except Exception:
                    return

    def close(self):
        self._closed = True
        try:
            self._listener.close()
        except Exception:
            pass
        self._loop.call_soon_threadsafe(self._loop.stop)
        with ActorServer._instance_lock:
            if ActorServer._instance is self:
                ActorServer._instance = None


# ---------------------------------------------------------------------------------------------
# actor meshes (the StorageVolume fleet)
# ---------------------------------------------------------------------------------------------
class _MeshEndpoint:
    def __init__(self, members, method: str):
        self._members, self._method = members, method

    async def call(self, *args, **kwargs):
        out = []
        for coord, ref in self._members:
            out.append((coord, await getattr(ref, self._method).call_one(*args, **kwargs)))
        return out

    async def call_one(self, *args, **kwargs):
        if len(self._members) != 1:
            raise ActorError("ValueError: call_one on a mesh with more than one actor")
        return await getattr(self._members[0][1], self._method).call_one(*args, **kwargs)


class ActorMesh:
    """An ordered set of actor refs addressed by a coordinate dict, e.g. {"gpus": 3}."""

    def __init__(self, members: list[tuple[dict, Any]]):
        self._members = list(members)

    def slice(self, **coord):
        for c, ref in self._members:
            if c == coord:
                return ref
        raise KeyError(f"no actor at {coord}")

    def __len__(self):
        return len(self._members)

    def __getattr__(self, method: str):
        if method.startswith("_"):
            raise AttributeError(method)
        return _MeshEndpoint(self._members, method)

    def __reduce__(self):
        return (ActorMesh, (self._members,))
