"""Pluggable structured logging (reference §5.5: Julia ``Logging`` records, the
``with_logger(lg) do train(...) end`` idiom of ``README.md:72-92`` and the Wandb
adapter ``src/loggers/wandb.jl:1``).

A logger is anything with ``info(name, **kv)``.  ``with_logger(lg)`` swaps the
process-wide current logger for a block, like Julia's ``with_logger``.
"""
from __future__ import annotations

import contextlib
import json
import sys
import threading
import time
from typing import Any, Dict, List, Optional


class Logger:
    def info(self, name: str, **kv: Any) -> None:  # pragma: no cover - interface
        raise NotImplementedError

    def close(self) -> None:
        pass


class ConsoleLogger(Logger):
    """Default: print ``[ Info: name  k=v ...`` records like Julia's ConsoleLogger."""

    def __init__(self, stream=None):
        self.stream = stream

    def info(self, name, **kv):
        s = "  ".join(f"{k}={v}" for k, v in kv.items())
        print(f"[ Info: {name}  {s}".rstrip(), file=self.stream or sys.stdout, flush=True)


class NullLogger(Logger):
    def info(self, name, **kv):
        pass


class MemoryLogger(Logger):
    """Keeps records in a list (tests; also the Dict-as-config trick of wandb.jl:1)."""

    def __init__(self):
        self.records: List[Dict[str, Any]] = []
        self._lock = threading.Lock()

    def info(self, name, **kv):
        with self._lock:
            self.records.append({"name": name, "t": time.time(), **kv})

    def named(self, name):
        return [r for r in self.records if r["name"] == name]


class JSONLLogger(Logger):
    """One JSON object per record, append-only; survives crashes (flush per line)."""

    def __init__(self, path: str):
        self._f = open(path, "a", buffering=1)
        self._lock = threading.Lock()

    def info(self, name, **kv):
        rec = {"name": name, "t": time.time()}
        for k, v in kv.items():
            try:
                json.dumps(v)
                rec[k] = v
            except TypeError:
                rec[k] = repr(v)
        with self._lock:
            self._f.write(json.dumps(rec) + "\n")

    def close(self):
        self._f.close()


class WandbLogger(Logger):
    """Adapter with the constructor of the README recipe
    (``WandbLogger(project=, name=, config=, step_increment=)``, ``README.md:83-87``).
    ``wandb`` is optional; without it the config lives in a plain dict and the
    records go to a fallback logger — the role of ``loggers/wandb.jl:1``."""

    def __init__(self, project: str = "DDP", name: Optional[str] = None, config: Optional[dict] = None,
                 step_increment: int = 1, fallback: Optional[Logger] = None):
        self.config = dict(config or {})
        self.step = 0
        self.step_increment = step_increment
        self._fallback = fallback or ConsoleLogger()
        try:  # pragma: no cover - wandb is not installed in the build image
            import wandb  # type: ignore

            self._run = wandb.init(project=project, name=name, config=self.config)
        except Exception:
            self._run = None

    def get_config(self, key, default=None):
        """``Wandb.get_config(lg, str) = get!(lg, str, nothing)``."""
        return self.config.setdefault(key, default)

    def update_config(self, d: dict, allow_val_change: bool = True):
        self.config.update(d)
        if self._run is not None:  # pragma: no cover
            self._run.config.update(d, allow_val_change=allow_val_change)

    def info(self, name, **kv):
        self.step += self.step_increment
        if self._run is not None:  # pragma: no cover
            self._run.log({f"{name}/{k}": v for k, v in kv.items()}, step=self.step)
        else:
            self._fallback.info(name, **kv)

    def close(self):
        if self._run is not None:  # pragma: no cover
            self._run.finish()


_current: Logger = ConsoleLogger()
_lock = threading.Lock()


def get_logger() -> Logger:
    return _current


def set_logger(lg: Logger) -> Logger:
    global _current
    with _lock:
        old, _current = _current, lg
    return old


@contextlib.contextmanager
def with_logger(lg: Logger):
    """``with_logger(lg) do ... end``."""
    old = set_logger(lg)
    try:
        yield lg
    finally:
        set_logger(old)
