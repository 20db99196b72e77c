"""Session-run hook protocol (``tf.estimator.SessionRunHook`` shaped) and the built-in hooks.

``before_run`` may return ``SessionRunArgs(fetches)``; the fetch the framework supports is the
global step (``get_global_step()`` sentinel), delivered as ``run_values.results`` in ``after_run``
-- the only pattern the reference uses (reference: tf_yarn/tensorflow/metrics.py:56-68).
"""
from __future__ import annotations

import logging
import time
from typing import Any, NamedTuple, Optional

logger = logging.getLogger(__name__)

GLOBAL_STEP = "global_step"


def get_global_step() -> str:
    """Sentinel fetch meaning "the global step after this run call"."""
    return GLOBAL_STEP


class SessionRunArgs(NamedTuple):
    fetches: Any = None
    feed_dict: Any = None
    options: Any = None


class SessionRunValues(NamedTuple):
    results: Any = None
    options: Any = None
    run_metadata: Any = None


class SessionRunContext:
    def __init__(self, estimator=None, step: int = 0):
        self.estimator = estimator
        self.step = step
        self.stop_requested = False
        self.original_args = None
        self.session = None

    def request_stop(self) -> None:
        self.stop_requested = True


class SessionRunHook:
    def begin(self): ...
    def after_create_session(self, session=None, coord=None): ...
    def before_run(self, run_context): return None
    def after_run(self, run_context, run_values): ...
    def end(self, session=None): ...


class StopAtStepHook(SessionRunHook):
    def __init__(self, num_steps: Optional[int] = None, last_step: Optional[int] = None):
        self.num_steps, self.last_step = num_steps, last_step

    def begin(self):
        self._start = None

    def before_run(self, run_context):
        return SessionRunArgs(get_global_step())

    def after_run(self, run_context, run_values):
        step = run_values.results
        if self._start is None:
            self._start = step - 1
        last = self.last_step if self.last_step is not None else self._start + self.num_steps
        if step >= last:
            run_context.request_stop()


class StepCounterHook(SessionRunHook):
    """Every N steps / seconds call ``_log_and_record(elapsed_steps, elapsed_time, global_step)``."""

    def __init__(self, every_n_steps: Optional[int] = 100, every_n_secs: Optional[float] = None, output_dir=None,
                 summary_writer=None):
        if (every_n_steps is None) == (every_n_secs is None):
            raise ValueError("exactly one of every_n_steps and every_n_secs should be provided.")
        self._every_steps, self._every_secs = every_n_steps, every_n_secs
        self._summary_writer = summary_writer
        self._output_dir = output_dir
        self._last_step: Optional[int] = None
        self._last_time: Optional[float] = None
        self.last_steps_per_sec: Optional[float] = None

    def before_run(self, run_context):
        return SessionRunArgs(get_global_step())

    def after_run(self, run_context, run_values):
        step = run_values.results
        now = time.time()
        if self._last_step is None:
            self._last_step, self._last_time = step, now
            return
        due = (step - self._last_step >= self._every_steps) if self._every_steps is not None \
            else (now - self._last_time >= self._every_secs)
        if due:
            self._log_and_record(step - self._last_step, now - self._last_time, step)
            self._last_step, self._last_time = step, now

    def _log_and_record(self, elapsed_steps: int, elapsed_time: float, global_step: int) -> None:
        sps = elapsed_steps / max(elapsed_time, 1e-9)
        self.last_steps_per_sec = sps
        if self._summary_writer is not None:
            self._summary_writer.add_scalar("global_step/sec", sps, global_step)
        logger.info("global_step/sec: %g", sps)


class LoggingTensorHook(SessionRunHook):
    """``tf.estimator.LoggingTensorHook(tensors, every_n_iter=None, every_n_secs=None, at_end=False, formatter=None)``:
    there are no graph tensors here, so the names in ``tensors`` are only echoed; what is logged is the training loss
    (the one quantity every train step produces) every N steps / seconds."""

    def __init__(self, tensors=None, every_n_iter: Optional[int] = None, every_n_secs: Optional[float] = None,
                 at_end: bool = False, formatter=None):
        if isinstance(tensors, int) and every_n_iter is None:        # LoggingTensorHook(100): the pre-round-2 call shape
            tensors, every_n_iter = None, tensors
        if every_n_iter is None and every_n_secs is None and not at_end:
            every_n_iter = 100
        self.tensors = list(tensors) if tensors is not None else []
        self.every_n_iter, self.every_n_secs, self.at_end, self.formatter = every_n_iter, every_n_secs, at_end, formatter
        self._last_time = time.time()
        self._estimator = None

    def before_run(self, run_context):
        return SessionRunArgs(get_global_step())

    def _emit(self, step, estimator) -> None:
        loss = getattr(estimator, "_last_loss_t", None)
        loss = float(loss) if loss is not None else estimator.last_loss
        values = {"step": step, "loss": loss}
        logger.info(self.formatter(values) if self.formatter else f"step = {step}, loss = {loss}")

    def after_run(self, run_context, run_values):
        step = run_values.results
        self._estimator = run_context.estimator
        if run_context.estimator is None:
            return
        due = (self.every_n_iter is not None and step % self.every_n_iter == 0) or \
              (self.every_n_secs is not None and time.time() - self._last_time >= self.every_n_secs)
        if due:
            self._last_time = time.time()
            self._emit(step, run_context.estimator)

    def end(self, session=None):
        if self.at_end and self._estimator is not None:
            self._emit(self._estimator.get_global_step(), self._estimator)


class NanTensorHook(SessionRunHook):
    """``tf.estimator.NanTensorHook(loss_tensor, fail_on_nan_loss=True)``: stop (or raise) when the loss is NaN / inf.
    Reads the loss every step (one 4-byte device-to-host copy)."""

    def __init__(self, loss_tensor=None, fail_on_nan_loss: bool = True):
        self.fail_on_nan_loss = fail_on_nan_loss

    def before_run(self, run_context):
        return SessionRunArgs(get_global_step())

    def after_run(self, run_context, run_values):
        est = run_context.estimator
        loss = getattr(est, "_last_loss_t", None) if est is not None else None
        if loss is None:
            return
        value = float(loss)
        if value != value or value in (float("inf"), float("-inf")):
            if self.fail_on_nan_loss:
                raise RuntimeError(f"NaN loss during training (step {run_values.results})")
            logger.warning("NaN loss at step %s: stopping", run_values.results)
            run_context.request_stop()
