"""RunConfig / session config / TF_CONFIG parsing for the Estimator facade.

The task programs read and mutate these attributes exactly like the reference
does on ``tf.estimator.RunConfig`` (reference: tf_yarn/tensorflow/tasks/
_independent_workers_task.py:27, gloo_allred_task.py:59-68,
tf_yarn/tensorflow/metrics.py:119-121, _tensorboard_task.py:39).
"""
from __future__ import annotations

import copy
import json
import os
from typing import Dict, List, Optional


class SessionConfig:
    """Stand-in for ``tf.compat.v1.ConfigProto``: only ``device_filters`` matters to the launcher."""

    def __init__(self, device_filters: Optional[List[str]] = None, **kwargs):
        self.device_filters = list(device_filters or [])
        self.extra = kwargs

    def __repr__(self):
        return f"SessionConfig(device_filters={self.device_filters})"


ConfigProto = SessionConfig

_DEFAULT = object()


class RunConfig:
    def __init__(self, model_dir: Optional[str] = None, tf_random_seed: Optional[int] = None,
                 save_summary_steps: Optional[int] = 100, save_checkpoints_steps=_DEFAULT,
                 save_checkpoints_secs=_DEFAULT, session_config: Optional[SessionConfig] = None,
                 keep_checkpoint_max: int = 5, log_step_count_steps: Optional[int] = 100,
                 compute_dtype: str = "bfloat16"):
        if save_checkpoints_steps is _DEFAULT and save_checkpoints_secs is _DEFAULT:
            save_checkpoints_steps, save_checkpoints_secs = None, 600
        elif save_checkpoints_steps is _DEFAULT:
            save_checkpoints_steps = None
        elif save_checkpoints_secs is _DEFAULT:
            save_checkpoints_secs = None
        self.model_dir = model_dir
        self.tf_random_seed = tf_random_seed
        self.save_summary_steps = save_summary_steps
        self.save_checkpoints_steps = save_checkpoints_steps
        self.save_checkpoints_secs = save_checkpoints_secs
        self.session_config = session_config
        self.keep_checkpoint_max = keep_checkpoint_max
        self.log_step_count_steps = log_step_count_steps
        self.compute_dtype = compute_dtype

    def replace(self, **kwargs) -> "RunConfig":
        new = copy.copy(self)
        for k, v in kwargs.items():
            if not hasattr(new, k):
                raise ValueError(f"RunConfig has no property {k!r}")
            setattr(new, k, v)
        return new

    # cluster view (from TF_CONFIG) -------------------------------------------------
    @property
    def cluster(self) -> "ClusterInfo":
        return ClusterInfo.from_env()

    @property
    def task_type(self) -> str:
        return self.cluster.task_type

    @property
    def task_id(self) -> int:
        return self.cluster.task_id

    @property
    def is_chief(self) -> bool:
        return self.cluster.is_chief

    @property
    def num_ps_replicas(self) -> int:
        return len(self.cluster.spec.get("ps", []))

    @property
    def num_worker_replicas(self) -> int:
        return len(self.cluster.spec.get("worker", [])) + len(self.cluster.spec.get("chief", []))


class ClusterInfo:
    """Parsed ``TF_CONFIG`` (``{"cluster": {...}, "task": {"type", "index"}, "environment"}``)."""

    def __init__(self, spec: Dict[str, List[str]], task_type: str, task_id: int, environment: str = ""):
        self.spec, self.task_type, self.task_id, self.environment = spec, task_type, task_id, environment

    @classmethod
    def from_env(cls) -> "ClusterInfo":
        raw = os.environ.get("TF_CONFIG")
        if not raw:
            return cls({}, "chief", 0)
        cfg = json.loads(raw)
        task = cfg.get("task", {})
        return cls(cfg.get("cluster", {}), task.get("type", "chief"), int(task.get("index", 0)),
                   cfg.get("environment", ""))

    @property
    def is_chief(self) -> bool:
        return self.task_type == "chief" or (not self.spec and self.task_type != "evaluator")

    @property
    def distributed(self) -> bool:
        return bool(self.spec)

    @property
    def has_ps(self) -> bool:
        return bool(self.spec.get("ps"))

    def trainers(self) -> List[str]:
        """Task keys of the roles that run training steps, chief first."""
        out = [f"chief:{i}" for i in range(len(self.spec.get("chief", [])))]
        out += [f"worker:{i}" for i in range(len(self.spec.get("worker", [])))]
        return out
