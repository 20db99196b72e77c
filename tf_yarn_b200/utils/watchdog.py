"""Step watchdog: failure detection for the training loops.

A data-parallel step on this runtime ends in a kernel that spins on flags written by the peer GPUs
(``ops/csrc/tfy_common.cuh``: ``tfy_grid_entry`` / ``tfy_grid_exit``).  If a peer rank dies (Python exception, OOM,
killed container) the surviving ranks' kernels wait forever and their host threads block in a CUDA synchronisation
that no Python exception can interrupt -- the application would hang instead of failing.  The reference inherits
the equivalent protection from its dependencies (NCCL's watchdog thread under torch DDP, gRPC deadlines under the
TF parameter server; reference: tf_yarn/pytorch/tasks/worker.py:101, tf_yarn/tensorflow/tasks/tf_task_common.py:46-50).

Here the loops call :meth:`StepWatchdog.beat` once per step; a daemon thread checks the heartbeat and, after
``timeout_secs`` without one, dumps every thread's stack and ends the process with :data:`EXIT_CODE`, which the
launcher reports as a failed task (``run_on_yarn`` then raises ``RunFailed`` or retries, ``nb_retries``).

``TFY_STEP_TIMEOUT_SECS`` overrides the timeout (0 disables); the default is :data:`DEFAULT_DISTRIBUTED_SECS` for
multi-rank GPU training and 0 otherwise.
"""
from __future__ import annotations

import faulthandler
import logging
import os
import sys
import threading
import time
from typing import Callable, Optional

logger = logging.getLogger(__name__)

EXIT_CODE = 75                      # EX_TEMPFAIL: a retry may succeed
DEFAULT_DISTRIBUTED_SECS = 1800.0


def default_timeout(distributed: bool) -> float:
    raw = os.environ.get("TFY_STEP_TIMEOUT_SECS")
    if raw is not None and raw.strip() != "":
        return max(0.0, float(raw))
    return DEFAULT_DISTRIBUTED_SECS if distributed else 0.0


def _default_action(what: str, idle_secs: float) -> None:
    logger.critical("%s made no progress for %.0f s: a peer rank is probably gone; stacks follow, exiting with %d",
                    what, idle_secs, EXIT_CODE)
    try:
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        sys.stderr.flush()
    finally:
        os._exit(EXIT_CODE)         # the main thread may sit in an uninterruptible CUDA call


class StepWatchdog:
    """``with StepWatchdog(timeout, "Model.fit") as wd: ... wd.beat()`` -- inert when ``timeout_secs`` is 0."""

    def __init__(self, timeout_secs: float, what: str = "training loop",
                 action: Optional[Callable[[str, float], None]] = None):
        self.timeout = float(timeout_secs)
        self.what = what
        self.action = action or _default_action
        self._last = time.monotonic()
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.fired = False

    def beat(self) -> None:
        self._last = time.monotonic()

    def start(self) -> "StepWatchdog":
        if self.timeout > 0 and self._thread is None:
            self.beat()
            self._thread = threading.Thread(target=self._run, name="tfy-step-watchdog", daemon=True)
            self._thread.start()
        return self

    def close(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
            self._thread = None

    def _run(self) -> None:
        period = min(5.0, max(0.01, self.timeout / 4.0))
        while not self._stop.wait(period):
            idle = time.monotonic() - self._last
            if idle > self.timeout:
                self.fired = True
                self.action(self.what, idle)
                return

    def __enter__(self) -> "StepWatchdog":
        return self.start()

    def __exit__(self, *exc) -> None:
        self.close()
