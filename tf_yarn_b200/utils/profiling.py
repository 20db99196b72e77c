"""Device timing and NVTX helpers.

The reference has no profiler hooks (only wall-clock ``catchtime`` prints and
``NCCL_DEBUG=INFO``).  Here every multi-GPU number is timed on the device with CUDA events and
reduced with max over ranks, and the hot regions can be annotated for Nsight.
"""
from __future__ import annotations

import contextlib
import time
from typing import Callable, Optional

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range around a region (no-op without CUDA)."""
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class DeviceTimer:
    """CUDA-event timer on a stream: ``with DeviceTimer() as t: ...; t.ms``."""

    def __init__(self, stream: Optional["torch.cuda.Stream"] = None):
        self.stream = stream
        self.ms: Optional[float] = None

    def __enter__(self):
        self._s = torch.cuda.Event(enable_timing=True)
        self._e = torch.cuda.Event(enable_timing=True)
        self._s.record(self.stream or torch.cuda.current_stream())
        return self

    def __exit__(self, *exc):
        self._e.record(self.stream or torch.cuda.current_stream())
        self._e.synchronize()
        self.ms = self._s.elapsed_time(self._e)
        return False


def time_kernel(fn: Callable[[], None], iters: int = 50, warmup: int = 5, flush_l2: bool = True) -> float:
    """Mean device time of ``fn`` in microseconds; optionally evicts L2 (writes 256 MiB) between calls."""
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if flush_l2 else None
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(iters):
        if scratch is not None:
            scratch.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        total += s.elapsed_time(e)
    return total / iters * 1e3


def max_over_ranks(value: float) -> float:
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


@contextlib.contextmanager
def catchtime(label: str = ""):
    """Wall-clock timer for host-side set-up steps (reference: tf_yarn/_task_commons.py:117-125)."""
    t0 = time.perf_counter()
    yield
    print(f"{label or 'step'}: {time.perf_counter() - t0:.3f} s")
