"""Exporters run by the evaluator after each evaluation (reference:
tf_yarn/tensorflow/tasks/evaluator_task.py:93-96,118-121: ``exporter.name`` and
``exporter.export(estimator, export_path, checkpoint_path, eval_result, is_the_final_export)``)."""
from __future__ import annotations

import os
import shutil
from typing import Callable, Optional


class Exporter:
    def __init__(self, name: str):
        self._name = name

    @property
    def name(self) -> str:
        return self._name

    def export(self, estimator, export_path, checkpoint_path, eval_result, is_the_final_export):
        raise NotImplementedError


class LatestExporter(Exporter):
    """Export after every evaluation, keeping the ``exports_to_keep`` most recent."""

    def __init__(self, name: str, serving_input_receiver_fn=None, exports_to_keep: Optional[int] = 5, **_ignored):
        super().__init__(name)
        self.exports_to_keep = exports_to_keep

    def export(self, estimator, export_path, checkpoint_path, eval_result, is_the_final_export):
        out = estimator.export_saved_model(export_path, checkpoint_path=checkpoint_path)
        if self.exports_to_keep:
            subdirs = sorted(d for d in os.listdir(export_path) if d.isdigit())
            for old in subdirs[:-self.exports_to_keep]:
                shutil.rmtree(os.path.join(export_path, old), ignore_errors=True)
        return out


class FinalExporter(Exporter):
    """Export only after the final evaluation."""

    def __init__(self, name: str, serving_input_receiver_fn=None, **_ignored):
        super().__init__(name)

    def export(self, estimator, export_path, checkpoint_path, eval_result, is_the_final_export):
        if not is_the_final_export:
            return None
        return estimator.export_saved_model(export_path, checkpoint_path=checkpoint_path)


class BestExporter(Exporter):
    """Export when ``compare_fn(best, current)`` says the new evaluation is better (default: lower loss)."""

    def __init__(self, name: str = "best_exporter", serving_input_receiver_fn=None,
                 compare_fn: Optional[Callable] = None, **_ignored):
        super().__init__(name)
        self._best = None
        self._compare = compare_fn or (lambda best, cur: cur["loss"] < best["loss"])

    def export(self, estimator, export_path, checkpoint_path, eval_result, is_the_final_export):
        if eval_result is None:
            return None
        if self._best is None or self._compare(self._best, eval_result):
            self._best = dict(eval_result)
            return estimator.export_saved_model(export_path, checkpoint_path=checkpoint_path)
        return None
