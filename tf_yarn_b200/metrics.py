"""Result of a run and one-shot KV loggers (reference: tf_yarn/metrics.py:19-59)."""
from __future__ import annotations

import logging
from datetime import timedelta
from typing import Dict, List, NamedTuple, Optional, Tuple

from tf_yarn_b200 import mlflow
from tf_yarn_b200.topologies import ContainerKey

logger = logging.getLogger(__name__)


class Metrics(NamedTuple):
    """Durations aggregated by the client from the tasks' lifecycle events."""
    total_training_duration: Optional[timedelta]
    total_eval_duration: Optional[timedelta]
    container_duration: Dict[ContainerKey, Optional[timedelta]]
    train_eval_time_per_node: Dict[ContainerKey, Optional[timedelta]]

    def as_text(self, n_try: int = 0) -> str:
        lines = []
        for name, value in self._asdict().items():
            if isinstance(value, dict):
                for key, dur in value.items():
                    if dur:
                        lines.append(f"{mlflow.format_key(f'{name}_{key}_{n_try}')}: {dur.total_seconds()} secs")
            elif value:
                lines.append(f"{mlflow.format_key(f'{name}_{n_try}')}: {value.total_seconds()} secs")
        return "".join(line + "\n" for line in lines)

    def log_mlflow(self, n_try: int) -> None:
        mlflow.save_text_to_mlflow(self.as_text(n_try), "tf_yarn_duration_stats")


class OneShotMetricsLogger:
    """Log (and tag in MLflow) each ``(kv_key, label)`` once, the first time the key exists."""

    def __init__(self, app, events: List[Tuple[str, str]], n_try: int = 0):
        self.app = app
        self.events = list(events)
        self.n_try = n_try

    def log(self) -> None:
        self.events = [ev for ev in self.events if not self._log_one(*ev)]

    def _log_one(self, key: str, label: str) -> bool:
        value = self.app.kv.get(key, None)
        if not value:
            return False
        value = value.decode() if isinstance(value, (bytes, bytearray)) else value
        logger.info("%s %s", label, value)
        mlflow.set_tag(f"{mlflow.format_key(key)}_{self.n_try}", value)
        return True
