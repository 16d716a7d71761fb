"""Prefetching loaders.

``DataLoader(f, (args...,), buffersize=k)`` reproduces the Flux-fork loader the
reference depends on (used at ``src/ddp_tasks.jl:278-283``): ``f`` is
mapped over ``zip(*args)`` on a background thread and up to ``buffersize`` results
are buffered.  ``DataLoader(f, data, batchsize=b)`` is the second form
(``src/sync.jl:136``): iterate mini-batches of an in-memory dataset, applying ``f``
(e.g. an upload) to each.

``StagedLoader`` is the B200 replacement for the reference's pageable
``gpu(shard)`` upload: batches are produced into a ring of *pinned*
host buffers, copied with ``cudaMemcpyAsync`` on a dedicated side stream and
handed to the compute stream through an event, so the H2D copy of step i+1
overlaps the compute of step i.  uint8 payloads are normalised on the GPU.
"""
from __future__ import annotations

import queue
import threading
from typing import Any, Callable, Iterable, Iterator, Optional, Sequence, Tuple

import torch

_SENTINEL = object()


class DataLoader:
    def __init__(self, f: Callable, data, buffersize: int = 5, batchsize: Optional[int] = None,
                 epochs: int = 1, shuffle: bool = False):
        self.f = f
        self.buffersize = max(int(buffersize), 1)
        self.batchsize = batchsize
        self.epochs = epochs
        self.shuffle = shuffle
        self.data = data
        self._q: Optional[queue.Queue] = None
        self._thread: Optional[threading.Thread] = None
        self._stop = threading.Event()

    # ---- item source
    def _items(self) -> Iterator[Tuple]:
        if self.batchsize is None:
            args = self.data if isinstance(self.data, tuple) else (self.data,)
            yield from zip(*args)
            return
        xs = self.data if isinstance(self.data, (tuple, list)) else (self.data,)
        n = xs[0].shape[0]
        for _ in range(self.epochs):
            perm = torch.randperm(n) if self.shuffle else torch.arange(n)
            for s in range(0, n - self.batchsize + 1, self.batchsize):
                idx = perm[s:s + self.batchsize]
                b = tuple(x[idx] for x in xs)
                yield (b if len(b) > 1 else b[0],)

    def __len__(self):
        if self.batchsize is None:
            args = self.data if isinstance(self.data, tuple) else (self.data,)
            return min(len(a) for a in args)
        xs = self.data if isinstance(self.data, (tuple, list)) else (self.data,)
        return (xs[0].shape[0] // self.batchsize) * self.epochs

    def _producer(self):
        assert self._q is not None
        try:
            for item in self._items():
                if self._stop.is_set():
                    break
                self._q.put(self.f(*item))
        except BaseException as e:  # surface loader errors at the consumer (errormonitor)
            self._q.put(e)
        finally:
            self._q.put(_SENTINEL)

    def __iter__(self):
        self._stop.clear()
        self._q = queue.Queue(maxsize=self.buffersize)
        self._thread = threading.Thread(target=self._producer, daemon=True, name="fluxdist-loader")
        self._thread.start()
        return self

    def __next__(self):
        assert self._q is not None, "call iter() first"
        item = self._q.get()
        if item is _SENTINEL:
            raise StopIteration
        if isinstance(item, BaseException):
            raise item
        return item

    def close(self):
        self._stop.set()
        if self._q is not None:
            try:
                while True:
                    self._q.get_nowait()
            except queue.Empty:
                pass


class StagedLoader:
    """Pinned ring + side-stream H2D with event hand-off.

    ``produce(slot_x, slot_y)`` fills pinned host tensors in place (called on a
    background thread).  Iterating yields device tensors that are safe to use on
    the *current* stream of ``device``.
    """

    def __init__(self, produce: Callable[[torch.Tensor, torch.Tensor], None], steps: int, device,
                 x_shape: Sequence[int], y_shape: Sequence[int], x_dtype=torch.uint8, y_dtype=torch.int64,
                 depth: int = 3, transform: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        self.produce, self.steps, self.depth = produce, steps, max(depth, 2)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.transform = transform
        pin = self.cuda
        self.hx = [torch.empty(tuple(x_shape), dtype=x_dtype, pin_memory=pin) for _ in range(self.depth)]
        self.hy = [torch.empty(tuple(y_shape), dtype=y_dtype, pin_memory=pin) for _ in range(self.depth)]
        self.dx = [torch.empty(tuple(x_shape), dtype=x_dtype, device=self.device) for _ in range(self.depth)]
        self.dy = [torch.empty(tuple(y_shape), dtype=y_dtype, device=self.device) for _ in range(self.depth)]
        self.h2d_bytes = self.hx[0].numel() * self.hx[0].element_size() + self.hy[0].numel() * self.hy[0].element_size()
        if self.cuda:
            self.stream = torch.cuda.Stream(device=self.device)
            self.copied = [torch.cuda.Event() for _ in range(self.depth)]
            self.consumed = [torch.cuda.Event() for _ in range(self.depth)]
        self._free: "queue.Queue[int]" = queue.Queue()
        self._ready: "queue.Queue[Any]" = queue.Queue()
        self._thread: Optional[threading.Thread] = None

    def _producer(self):
        try:
            for _ in range(self.steps):
                s = self._free.get()
                if s is None:
                    return
                self.produce(self.hx[s], self.hy[s])
                self._ready.put(s)
        except BaseException as e:
            self._ready.put(e)
        self._ready.put(_SENTINEL)

    def __iter__(self):
        for s in range(self.depth):
            self._free.put(s)
        self._thread = threading.Thread(target=self._producer, daemon=True, name="fluxdist-staged")
        self._thread.start()
        self._last: Optional[int] = None
        return self

    def __len__(self):
        return self.steps

    def __next__(self):
        if self._last is not None:  # slot handed out last time is consumed by work queued so far
            if self.cuda:
                self.consumed[self._last].record(torch.cuda.current_stream(self.device))
            self._free.put(self._last)
            self._last = None
        s = self._ready.get()
        if s is _SENTINEL:
            raise StopIteration
        if isinstance(s, BaseException):
            raise s
        if self.cuda:
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(self.consumed[s])  # device buffer no longer read by compute
                self.dx[s].copy_(self.hx[s], non_blocking=True)
                self.dy[s].copy_(self.hy[s], non_blocking=True)
                self.copied[s].record(self.stream)
            torch.cuda.current_stream(self.device).wait_event(self.copied[s])
        else:
            self.dx[s].copy_(self.hx[s])
            self.dy[s].copy_(self.hy[s])
        self._last = s
        x = self.dx[s]
        if self.transform is not None:
            x = self.transform(x)
        return x, self.dy[s]

    def close(self):
        self._free.put(None)
