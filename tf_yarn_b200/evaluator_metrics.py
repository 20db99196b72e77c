"""Client-side poller of the evaluator's liveness statistics.

The evaluator task publishes four numbers to the KV store after every
evaluation step; the client logs them when they change and lie inside the
user's thresholds (reference: tf_yarn/evaluator_metrics.py:12-70; scenarios in
reference tests/test_evaluator_metrics.py:14-55).
"""
from __future__ import annotations

import logging
import sys
import warnings
from typing import Dict, List, Optional, Tuple

from tf_yarn_b200 import mlflow
from tf_yarn_b200.topologies import ContainerTask

MONITORED_METRICS = {
    "awake_time_ratio": "Awake/idle ratio",
    "eval_step_mean_duration": "Eval step mean duration (in sec)",
    "last_training_step": "Training step of last checkpoint",
    "nb_eval_steps": "Number of evaluation steps done",
}

logger = logging.getLogger(__name__)


class EvaluatorMetricsLogger:
    def __init__(self, evaluator_list: List[ContainerTask], app,
                 log_thresholds: Optional[Dict[str, Tuple[float, float]]] = None, n_try: int = 0):
        self.evaluator_list = list(evaluator_list)
        self.app = app
        self.n_try = n_try
        self.last_metrics = {
            ev.to_container_key(): {metric: None for metric in MONITORED_METRICS} for ev in self.evaluator_list
        }
        self.log_thresholds: Dict[str, Tuple[float, float]] = {}
        if log_thresholds:
            for key, (lo, hi) in log_thresholds.items():
                if key in MONITORED_METRICS:
                    self.log_thresholds[key] = (lo if lo else 0, hi if hi else sys.float_info.max)
            unknown = set(log_thresholds) - set(MONITORED_METRICS)
            if unknown:
                warnings.warn(f"The following evaluation metrics are not monitored: {sorted(unknown)}")

    def log(self) -> None:
        for evaluator in self.evaluator_list:
            key = evaluator.to_container_key()
            fresh = []
            for metric, label in MONITORED_METRICS.items():
                raw = self.app.kv.get(f"{key.to_kv_str()}/{metric}", None)
                if not raw:
                    continue
                stat = float(raw.decode() if isinstance(raw, (bytes, bytearray)) else raw)
                if stat == self.last_metrics[key][metric]:
                    continue
                bounds = self.log_thresholds.get(metric)
                if bounds is None or bounds[0] <= stat <= bounds[1]:
                    fresh.append(f"{label}: {stat}")
                # state and MLflow are updated whether or not the value was inside the bounds
                self.last_metrics[key][metric] = stat
                mlflow.log_metric(mlflow.format_key(f"{key.to_kv_str()}_{metric}_{self.n_try}"), stat)
            if fresh:
                logger.info(f"Statistics for {key.to_kv_str()}: {' '.join(fresh)}")
