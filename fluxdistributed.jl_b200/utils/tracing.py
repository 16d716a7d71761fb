"""Tracing / profiling helpers (the reference has none; the metric
requires device timing as max over ranks).

* ``nvtx_range`` — NVTX ranges per phase (load / fwd / bwd / allreduce+update).
* ``DeviceTimer`` — CUDA-event timing on the launching stream.
* ``PhaseTimes`` — accumulates per-phase device/host time for the step breakdown.
* ``ClockSampler`` — samples ``nvidia-smi`` clocks/throttle reasons during a timed
  region (bench contract: the ``clocks`` key).
"""
from __future__ import annotations

import contextlib
import statistics
import subprocess
import threading
import time
from collections import defaultdict
from typing import Dict, List, Optional

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class DeviceTimer:
    """CUDA events on the current stream; falls back to wall-clock on CPU."""

    def __init__(self, device=None):
        self.cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
        self.device = device
        if self.cuda:
            self._a = torch.cuda.Event(enable_timing=True)
            self._b = torch.cuda.Event(enable_timing=True)
        self._t0 = 0.0
        self._t1 = 0.0

    def start(self):
        if self.cuda:
            self._a.record()
        else:
            self._t0 = time.perf_counter()

    def stop(self):
        if self.cuda:
            self._b.record()
        else:
            self._t1 = time.perf_counter()

    def elapsed_ms(self) -> float:
        if self.cuda:
            self._b.synchronize()
            return self._a.elapsed_time(self._b)
        return (self._t1 - self._t0) * 1e3


class PhaseTimes:
    def __init__(self):
        self.ms: Dict[str, List[float]] = defaultdict(list)

    @contextlib.contextmanager
    def phase(self, name: str, device=None):
        t = DeviceTimer(device)
        with nvtx_range(name):
            t.start()
            yield
            t.stop()
        self.ms[name].append(t.elapsed_ms())

    def summary(self) -> Dict[str, float]:
        return {k: statistics.fmean(v) for k, v in self.ms.items() if v}


class ClockSampler:
    """Background SM-clock / throttle-reason sampler for the timed region of a benchmark (B200_PROFILING.md clocks
    line).  Samples NVML in-process every ``period_ms`` (a query costs microseconds, so even a 100 ms region gets
    several samples); falls back to a looping ``nvidia-smi`` subprocess when NVML is not importable."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0, period_ms: int = 10):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self._proc: Optional[subprocess.Popen] = None
        self._lines: List[str] = []
        self._thread: Optional[threading.Thread] = None
        self._stop = threading.Event()
        self._sm: List[float] = []
        self._smax: List[float] = []
        self._reasons: set = set()
        self._nvml = None
        self._recording = threading.Event()

    # ---------------------------------------------------------------- NVML path
    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        try:  # CUDA_VISIBLE_DEVICES-proof: resolve the CUDA device to its NVML handle by UUID
            import torch

            uuid = str(torch.cuda.get_device_properties(self.gpu_index).uuid)
            uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)

    def _sample_nvml(self):
        if not self._recording.is_set():
            return
        nv, h = self._nvml
        self._sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
        self._smax.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
        for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                          ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                          ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
            if mask & int(bit):
                self._reasons.add(name)

    def resume(self):
        """Begin recording (cheap: sets a flag).  Pair with ``start(paused=True)`` so the NVML/nvidia-smi start-up cost
        stays outside a timed region."""
        self._sm.clear(); self._smax.clear(); self._reasons.clear(); self._lines.clear()
        self._recording.set()
        return self

    def start(self, paused: bool = False):
        """Start sampling.  ``paused=True`` does all the expensive set-up (NVML init, device lookup, sampler thread)
        but records nothing until :meth:`resume` — bench.py calls it BEFORE the pre-timing barrier so no rank enters
        the timed region late (round-1 defect: the set-up ran between the barrier and the first event record)."""
        self._recording.set()
        try:
            self._nvml = self._nvml_handle()
            self._sample_nvml()
            self._sm.clear(); self._smax.clear(); self._reasons.clear()  # the probe above may predate the load
        except Exception:
            self._nvml = None
        if self._nvml is not None:
            def loop():
                while not self._stop.is_set():
                    try:
                        self._sample_nvml()
                    except Exception:
                        pass
                    self._stop.wait(self.period_ms / 1000.0)

            if paused:
                self._recording.clear()
            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            return self
        try:
            self._proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-i", str(self.gpu_index), "-lms", str(max(self.period_ms, 50))],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._proc = None
            return self

        def pump():
            assert self._proc is not None and self._proc.stdout is not None
            for line in self._proc.stdout:
                if self._recording.is_set():
                    self._lines.append(line.strip())

        if paused:
            self._recording.clear()
        self._thread = threading.Thread(target=pump, daemon=True)
        self._thread.start()
        return self

    def stop(self) -> Dict:
        self._recording.set()
        self._stop.set()
        if self._nvml is not None:
            if self._thread is not None:
                self._thread.join(timeout=1.0)
            try:
                self._sample_nvml()  # one last sample: a very short region still reports what the clock was
            except Exception:
                pass
        if self._proc is not None:
            self._proc.terminate()  # exact PID we started
            try:
                self._proc.wait(timeout=2)
            except Exception:
                self._proc.kill()
        sm, smax, reasons = list(self._sm), list(self._smax), set(self._reasons)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self._lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                "samples": len(sm)}
