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


class _CudaView:
    """Zero-copy torch view of raw device memory through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


_TYPESTR = {torch.uint8: "|u1", torch.int8: "|i1", torch.int32: "<i4", torch.int64: "<i8", torch.float32: "<f4",
            torch.float16: "<f2", torch.float64: "<f8", torch.int16: "<i2", torch.bool: "|b1"}


def _host_view(ptr: int, shape, dtype: torch.dtype) -> torch.Tensor:
    import ctypes

    n = 1
    for d in shape:
        n *= int(d)
    nbytes = n * torch.empty(0, dtype=dtype).element_size()
    buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(int(ptr))
    t = torch.frombuffer(buf, dtype=torch.uint8, count=nbytes)
    return t.view(dtype).view(tuple(shape))


def _dev_view(ptr: int, shape, dtype: torch.dtype, device) -> torch.Tensor:
    if dtype == torch.bfloat16:  # no typestr for bf16: view as int16 and reinterpret
        return torch.as_tensor(_CudaView(ptr, shape, "<i2"), device=device).view(torch.bfloat16)
    return torch.as_tensor(_CudaView(ptr, shape, _TYPESTR[dtype]), device=device)


class StagedLoader:
    """Pinned ring + copy-stream H2D with event hand-off.

    ``produce(slot_x, slot_y)`` fills pinned host tensors in place (called on a background thread); iterating yields
    device tensors that are safe to use on the *current* stream of ``device``.  On CUDA with the extension the ring is
    the native one (``csrc/staging.cpp::PinnedRing``: cudaHostAlloc slots, copies issued up to ``prefetch`` batches
    ahead, host slot reuse gated on the copy event); otherwise a pure-PyTorch ring with the same protocol (CPU: plain
    copies).  ``StagedLoader.from_batches`` adapts a ``batch_fn(arg) -> (x, y)`` producer (what ``prepare_training``
    and the process-mode workers use).
    """

    def __init__(self, produce: Callable[[torch.Tensor, torch.Tensor], None], steps: int, device,
                 x_shape: Sequence[int], y_shape: Sequence[int], x_dtype=torch.uint8, y_dtype=torch.int64,
                 depth: int = 4, transform: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 prefetch: int = 2, native: Optional[bool] = None):
        self.produce, self.steps, self.depth = produce, steps, max(depth, 3)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.transform = transform
        self.prefetch = max(1, min(prefetch, self.depth - 1))
        self.x_shape, self.y_shape, self.x_dtype, self.y_dtype = tuple(x_shape), tuple(y_shape), x_dtype, y_dtype
        xb = _numel(x_shape) * torch.empty(0, dtype=x_dtype).element_size()
        yb = _numel(y_shape) * torch.empty(0, dtype=y_dtype).element_size()
        self.h2d_bytes = xb + yb
        self.ring = None
        if native is None:
            native = self.cuda and _native_ring_ok()
        if native:
            from .. import ops

            idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
            self.ring = ops.extension().PinnedRing(idx, xb, yb, self.depth, self.prefetch)
            self.hx = [_host_view(self.ring.host_x(s), x_shape, x_dtype) for s in range(self.depth)]
            self.hy = [_host_view(self.ring.host_y(s), y_shape, y_dtype) for s in range(self.depth)]
            dev = torch.device("cuda", idx)
            self.dx = [_dev_view(self.ring.dev_x(s), x_shape, x_dtype, dev) for s in range(self.depth)]
            self.dy = [_dev_view(self.ring.dev_y(s), y_shape, y_dtype, dev) for s in range(self.depth)]
        else:
            pin = self.cuda
            self.hx = [torch.empty(self.x_shape, dtype=x_dtype, pin_memory=pin) for _ in range(self.depth)]
            self.hy = [torch.empty(self.y_shape, dtype=y_dtype, pin_memory=pin) for _ in range(self.depth)]
            self.dx = [torch.empty(self.x_shape, dtype=x_dtype, device=self.device) for _ in range(self.depth)]
            self.dy = [torch.empty(self.y_shape, dtype=y_dtype, device=self.device) for _ in range(self.depth)]
            if self.cuda:
                self.stream = torch.cuda.Stream(device=self.device)
                self.copied = [torch.cuda.Event() for _ in range(self.depth)]
                self.consumed = [torch.cuda.Event() for _ in range(self.depth)]
        self._free: "queue.Queue[Any]" = queue.Queue()
        self._ready: "queue.Queue[Any]" = queue.Queue()
        self._inflight: list = []
        self._thread: Optional[threading.Thread] = None
        self._last: Optional[int] = None
        self._first: Optional[Tuple] = None

    # ---- adapter: batch_fn(arg) -> (x, y) CPU tensors, shapes taken from the first batch
    @classmethod
    def from_batches(cls, batch_fn: Callable, args: Sequence, device, depth: int = 4,
                     transform: Optional[Callable] = None, prefetch: int = 2) -> "StagedLoader":
        args = list(args)
        if not args:
            raise ValueError("StagedLoader.from_batches needs at least one batch")
        first = batch_fn(args[0])
        x0, y0 = first
        it = iter(args[1:])
        state = {"first": (x0, y0)}

        held = {}  # pinned slot -> signature of the source tensors it currently holds

        def sig(t):
            return (t.data_ptr(), tuple(t.shape), t.dtype, t._version)

        def produce(sx, sy):
            if state["first"] is not None:
                x, y = state["first"]
                state["first"] = None
            else:
                x, y = batch_fn(next(it))
            # a source that keeps handing out the same (unmodified) host tensors — cached / synthetic batches cycled with
            # the period of the ring — is not copied into a slot that already holds it; the H2D copy still happens
            k = (sig(x), sig(y))
            if held.get(sx.data_ptr()) != k:
                sx.copy_(x)
                sy.copy_(y)
                held[sx.data_ptr()] = k

        return cls(produce, len(args), device, tuple(x0.shape), tuple(y0.shape), x0.dtype, y0.dtype, depth=depth,
                   transform=transform, prefetch=prefetch)

    # ---- producer thread
    def _producer(self):
        try:
            for _ in range(self.steps):
                if self.ring is not None:
                    s = self.ring.acquire()
                    if s < 0:
                        return
                else:
                    s = self._free.get()
                    if s is None:
                        return
                self.produce(self.hx[s], self.hy[s])
                if self.ring is not None:
                    self.ring.commit(s)
                self._ready.put(s)
        except BaseException as e:  # surface loader errors at the consumer
            self._ready.put(e)
        self._ready.put(_SENTINEL)

    def __iter__(self):
        if self.ring is None:
            for s in range(self.depth):
                self._free.put(s)
        self._served = 0
        self._thread = threading.Thread(target=self._producer, daemon=True, name="fluxdist-staged")
        self._thread.start()
        self._last = None
        return self

    def __len__(self):
        return self.steps

    def _issue_copy(self, s: int) -> None:
        if self.cuda:
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(self.consumed[s])  # device buffer no longer read by compute
                self.dx[s].copy_(self.hx[s], non_blocking=True)
                self.dy[s].copy_(self.hy[s], non_blocking=True)
                self.copied[s].record(self.stream)
        else:
            self.dx[s].copy_(self.hx[s])
            self.dy[s].copy_(self.hy[s])
        self._inflight.append(s)

    def _top_up(self, block: bool) -> bool:
        """Issue copies for ready slots (python ring).  Returns False when the producer is exhausted."""
        while len(self._inflight) < self.prefetch:
            try:
                s = self._ready.get(block=block)
            except queue.Empty:
                return True
            if s is _SENTINEL:
                self._ready.put(_SENTINEL)
                return False
            if isinstance(s, BaseException):
                raise s
            self._issue_copy(s)
            block = False
        return True

    def __next__(self):
        if self._served >= self.steps:
            raise StopIteration
        if self.ring is not None:
            # errors of the producer thread surface here; the native ring hands out slots in commit order
            while True:
                try:
                    head = self._ready.get_nowait()
                except queue.Empty:
                    break
                if isinstance(head, BaseException):
                    raise head
            stream = torch.cuda.current_stream(self.device).cuda_stream
            s, _, _ = self.ring.next(stream)
        else:
            if self._last is not None:  # slot handed out last time is consumed by work queued so far
                if self.cuda:
                    self.consumed[self._last].record(torch.cuda.current_stream(self.device))
                    # the pinned buffers may be refilled only once the copy engine has read them (round-1 race)
                    self.copied[self._last].synchronize()
                self._free.put(self._last)
                self._last = None
            self._top_up(False)
            if not self._inflight:
                self._top_up(True)
            if not self._inflight:
                raise StopIteration
            s = self._inflight.pop(0)
            if self.cuda:
                torch.cuda.current_stream(self.device).wait_event(self.copied[s])
            self._last = s
            self._top_up(False)  # the copy of the NEXT batch is queued before this batch's compute
        self._served += 1
        x = self.dx[s]
        if self.transform is not None:
            x = self.transform(x)
        return x, self.dy[s]

    def close(self):
        if self.ring is not None:
            self.ring.stop()
        else:
            self._free.put(None)


def _numel(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


def _native_ring_ok() -> bool:
    try:
        from .. import ops

        return ops.native_available() and hasattr(ops.extension(), "PinnedRing")
    except Exception:
        return False
