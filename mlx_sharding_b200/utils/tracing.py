"""Tracing hooks (SURVEY §5.1 — the reference has none beyond a tokens/sec print).

* ``nvtx_range(name)``: NVTX ranges around stage / layer execution when ``MLXB200_NVTX=1`` (visible in ncu / nsys
  timelines and in ``torch.profiler`` traces); a no-op otherwise so the hot path pays nothing.
* ``StageTimer``: CUDA-event timing of every stage forward (device time, not wall clock) with a cheap
  ring of recent samples; ``busy_fraction()`` feeds the ``/metrics`` endpoint.
"""
from __future__ import annotations

import contextlib
import os
import time
from collections import deque

import torch

_NVTX = os.environ.get("MLXB200_NVTX", "0") == "1"


@contextlib.contextmanager
def nvtx_range(name: str):
    if _NVTX and torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class StageTimer:
    def __init__(self, enabled: bool = False, keep: int = 256):
        self.enabled = enabled and torch.cuda.is_available()
        self.samples = deque(maxlen=keep)   # (wall_start, device_ms)
        self._pending = deque()
        self.t0 = time.perf_counter()

    @contextlib.contextmanager
    def measure(self):
        if not self.enabled:
            yield
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        self._pending.append((time.perf_counter(), e0, e1))
        while self._pending and self._pending[0][2].query():
            t, a, b = self._pending.popleft()
            self.samples.append((t, a.elapsed_time(b)))

    def busy_fraction(self) -> float:
        if not self.samples:
            return 0.0
        span = max(time.perf_counter() - self.samples[0][0], 1e-6)
        return min(1.0, sum(ms for _, ms in self.samples) / 1e3 / span)

    def mean_ms(self) -> float:
        return sum(ms for _, ms in self.samples) / len(self.samples) if self.samples else 0.0
