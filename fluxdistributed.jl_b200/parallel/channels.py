"""Channels: ``Channel(n)`` and ``RemoteChannel(() -> Channel(1), pid)`` semantics
(reference E4; ``bin/driver.jl:22-23``; ``src/sync.jl:41-43,75,145,148``).

* ``put`` blocks while the channel is full (capacity-1 back-pressure),
* ``take`` blocks until a value is present,
* ``isready`` is a non-blocking probe,
* ``None`` is a legal message — it is the abort signal of ``syncgrads``
  (``sync.jl:49-53``).

Three transports behind one interface:

``Channel``        in-process (threads) — task mode and unit tests
``RemoteChannel``  ``multiprocessing`` queue, usable from any process of one node;
                   payloads are serialised with ``torch.save`` (the analogue of Julia's
                   ``Serialization`` over sockets)
``StoreChannel``   a ``torch.distributed`` TCPStore-backed channel that works across
                   nodes ("multi-node training is handled by the same design",
                   ``README.md:7``)

On the B200 fast path the channels carry *descriptors* (step number, NaN flag,
arena epoch) rather than gradient bytes — the bytes move inside the fused kernel
over NVLink.
"""
from __future__ import annotations

import io
import pickle
import queue
import threading
import time
from typing import Any, Iterable, Optional, Sequence

import torch


class ChannelClosed(Exception):
    pass


class _Boxed:
    """Wrapper so that ``None`` can travel through queues that use None internally."""

    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v


def _dumps(obj) -> bytes:
    buf = io.BytesIO()
    torch.save(obj, buf)
    return buf.getvalue()


def _loads(b: bytes):
    return torch.load(io.BytesIO(b), map_location="cpu", weights_only=False)


class Channel:
    """In-process bounded channel."""

    def __init__(self, capacity: int = 1):
        self._q: "queue.Queue[_Boxed]" = queue.Queue(maxsize=max(capacity, 1))
        self._closed = False

    def put(self, v, timeout: Optional[float] = None) -> None:
        if self._closed:
            raise ChannelClosed()
        self._q.put(_Boxed(v), timeout=timeout)

    def take(self, timeout: Optional[float] = None):
        try:
            return self._q.get(timeout=timeout).v
        except queue.Empty:
            raise TimeoutError("channel take timed out") from None

    def isready(self) -> bool:
        return not self._q.empty()

    def fetch(self):
        """Peek without removing (Julia ``fetch``); only safe with one consumer."""
        with self._q.mutex:
            if not self._q.queue:
                raise queue.Empty
            return self._q.queue[0].v

    def close(self):
        self._closed = True


class RemoteChannel:
    """Cross-process capacity-``n`` channel living on this node.

    Create it in the parent *before* spawning workers and pass it as a process
    argument (the owner pid of the reference's ``RemoteChannel(f, pid)`` decides
    where the buffer lives; with OS pipes that distinction disappears)."""

    def __init__(self, capacity: int = 1, ctx=None, owner: int = 0):
        import torch.multiprocessing as mp

        self._ctx = ctx or mp.get_context("spawn")
        self._q = self._ctx.Queue(maxsize=max(capacity, 1))
        self.owner = owner

    def put(self, v, timeout: Optional[float] = None) -> None:
        self._q.put(_dumps(v), timeout=timeout)

    def take(self, timeout: Optional[float] = None):
        try:
            return _loads(self._q.get(timeout=timeout))
        except queue.Empty:
            raise TimeoutError("channel take timed out") from None

    def isready(self) -> bool:
        return not self._q.empty()

    def close(self):
        self._q.close()


class StoreChannel:
    """Capacity-1 single-producer/single-consumer channel over a c10d Store."""

    def __init__(self, store, name: str, poll_s: float = 0.0005):
        self.store, self.name, self.poll_s = store, name, poll_s
        self._put_seq = 0
        self._take_seq = 0

    def _count(self, key: str) -> int:
        return int(self.store.add(f"{self.name}/{key}", 0))

    def put(self, v, timeout: Optional[float] = None) -> None:
        t0 = time.monotonic()
        k = self._put_seq + 1
        while self._count("taken") < k - 1:  # capacity-1 back-pressure
            if timeout is not None and time.monotonic() - t0 > timeout:
                raise TimeoutError("channel put timed out")
            time.sleep(self.poll_s)
        self.store.set(f"{self.name}/msg/{k}", _dumps(v))
        self.store.add(f"{self.name}/put", 1)
        self._put_seq = k

    def take(self, timeout: Optional[float] = None):
        t0 = time.monotonic()
        k = self._take_seq + 1
        while self._count("put") < k:
            if timeout is not None and time.monotonic() - t0 > timeout:
                raise TimeoutError("channel take timed out")
            time.sleep(self.poll_s)
        data = self.store.get(f"{self.name}/msg/{k}")
        try:
            self.store.delete_key(f"{self.name}/msg/{k}")
        except Exception:
            pass
        self.store.add(f"{self.name}/taken", 1)
        self._take_seq = k
        return _loads(data)

    def isready(self) -> bool:
        return self._count("put") > self._count("taken")

    def close(self):
        pass


def wait_all_ready(chans: Sequence, poll_s: float = 0.0005, timeout: Optional[float] = None,
                   stop: Optional[threading.Event] = None) -> bool:
    """Block until every channel has a message.  Replaces the reference's hot
    busy-spin ``all(isready, ip) || continue`` (``sync.jl:41``, quirk Q3) with a
    sleeping poll; returns False on timeout / stop."""
    t0 = time.monotonic()
    while True:
        if all(c.isready() for c in chans):
            return True
        if stop is not None and stop.is_set():
            return False
        if timeout is not None and time.monotonic() - t0 > timeout:
            return False
        time.sleep(poll_s)
