"""Optional experiment tracking (MLflow when present, a local file tracker otherwise).

Every call is a no-op returning a default when tracking is disabled, and
tracking failures never break a run (reference: tf_yarn/mlflow.py:20-144).

Tracking is enabled when ``TF_YARN_USE_MLFLOW`` is not ``"False"`` and either

* the real ``mlflow`` package is importable and a tracking URI is configured, or
* ``TFY_TRACKING_DIR`` points at a directory: the built-in :class:`LocalTracker`
  then records params / metrics / tags / artifacts as plain files, which keeps
  the MLflow-style workflow usable on an air-gapped box.
"""
from __future__ import annotations

import functools
import json
import logging
import os
import shutil
import tempfile
import time
import uuid
import warnings
from typing import Any, Dict, List, Optional

logger = logging.getLogger(__name__)

_use_mlflow: Optional[bool] = None
_backend: Any = None


class LocalTracker:
    """Minimal file-system tracker with the subset of the MLflow fluent API the framework uses."""

    def __init__(self, root: str, run_id: Optional[str] = None):
        self.root = root
        self.run_id = run_id or os.environ.get("MLFLOW_RUN_ID") or uuid.uuid4().hex
        self.run_dir = os.path.join(root, self.run_id)
        for sub in ("metrics", "params", "tags", "artifacts"):
            os.makedirs(os.path.join(self.run_dir, sub), exist_ok=True)

    def active_run_id(self) -> str:
        return self.run_id

    def get_tracking_uri(self) -> str:
        return "file://" + os.path.abspath(self.root)

    def _write(self, kind: str, key: str, value: Any) -> None:
        with open(os.path.join(self.run_dir, kind, format_key(key)), "w") as f:
            f.write(str(value))

    def set_tag(self, key: str, value: Any) -> None:
        self._write("tags", key, value)

    def log_param(self, key: str, value: Any) -> None:
        self._write("params", key, value)

    def log_metric(self, key: str, value: float, step: Optional[int] = None) -> None:
        with open(os.path.join(self.run_dir, "metrics", format_key(key)), "a") as f:
            f.write(json.dumps({"ts": time.time(), "value": float(value), "step": step}) + "\n")

    def log_artifact(self, local_path: str, artifact_path: Optional[str] = None) -> None:
        dst = os.path.join(self.run_dir, "artifacts", artifact_path or "")
        os.makedirs(dst, exist_ok=True)
        shutil.copy(local_path, dst)

    def log_artifacts(self, local_dir: str, artifact_path: Optional[str] = None) -> None:
        dst = os.path.join(self.run_dir, "artifacts", artifact_path or "")
        shutil.copytree(local_dir, dst, dirs_exist_ok=True)

    # read side (tests, examples)
    def metric_history(self, key: str) -> List[Dict[str, Any]]:
        path = os.path.join(self.run_dir, "metrics", format_key(key))
        if not os.path.exists(path):
            return []
        with open(path) as f:
            return [json.loads(line) for line in f if line.strip()]

    def list(self, kind: str) -> List[str]:
        return sorted(os.listdir(os.path.join(self.run_dir, kind)))


class _RealMlflow:
    def __init__(self, module):
        self.m = module

    def active_run_id(self) -> str:
        run = self.m.active_run()
        if run is None:
            run = self.m.start_run()
        return run.info.run_id

    def get_tracking_uri(self) -> str:
        return self.m.get_tracking_uri()

    def __getattr__(self, name):
        return getattr(self.m, name)


def reset() -> None:
    """Forget the cached detection (tests; after changing the environment)."""
    global _use_mlflow, _backend
    _use_mlflow, _backend = None, None


def _detect() -> bool:
    global _backend
    if os.environ.get("TF_YARN_USE_MLFLOW", "") == "False":
        return False
    local_dir = os.environ.get("TFY_TRACKING_DIR")
    if local_dir:
        _backend = LocalTracker(local_dir)
        return True
    try:
        import mlflow as real  # noqa: F401
    except ImportError:
        return False
    if not (os.environ.get("MLFLOW_TRACKING_URI") or real.get_tracking_uri()):
        warnings.warn("MLflow is installed but no tracking URI is set; tracking disabled")
        return False
    _backend = _RealMlflow(real)
    return True


def use_mlflow() -> bool:
    global _use_mlflow
    if _use_mlflow is None:
        try:
            _use_mlflow = _detect()
        except Exception:  # noqa: BLE001
            logger.exception("experiment tracking detection failed; disabled")
            _use_mlflow = False
    return _use_mlflow


def backend():
    return _backend if use_mlflow() else None


def optional_mlflow(return_default=None):
    """Decorator: run only when tracking is enabled; swallow tracking-side failures."""
    def decorator(f):
        @functools.wraps(f)
        def wrapper(*args, **kwargs):
            if not use_mlflow():
                return return_default
            try:
                return f(*args, **kwargs)
            except (ConnectionError, OSError, RuntimeError) as exc:
                logger.error("experiment tracking call %s failed: %s", f.__name__, exc)
                return return_default
        return wrapper
    return decorator


@optional_mlflow(return_default="")
def active_run_id() -> str:
    return _backend.active_run_id()


@optional_mlflow(return_default="")
def get_tracking_uri() -> str:
    return _backend.get_tracking_uri()


@optional_mlflow()
def set_tag(key: str, value: Any) -> None:
    _backend.set_tag(key, value)


@optional_mlflow()
def set_tags(tags: Dict[str, Any]) -> None:
    for k, v in tags.items():
        _backend.set_tag(k, v)


@optional_mlflow()
def log_param(key: str, value: Any) -> None:
    _backend.log_param(key, value)


@optional_mlflow()
def log_params(params: Dict[str, Any]) -> None:
    for k, v in params.items():
        _backend.log_param(k, v)


@optional_mlflow()
def log_metric(key: str, value: float, step: Optional[int] = None) -> None:
    _backend.log_metric(key, value, step)


@optional_mlflow()
def log_metrics(metrics: Dict[str, float], step: Optional[int] = None) -> None:
    for k, v in metrics.items():
        _backend.log_metric(k, v, step)


@optional_mlflow()
def log_artifact(local_path: str, artifact_path: Optional[str] = None) -> None:
    _backend.log_artifact(local_path, artifact_path)


@optional_mlflow()
def log_artifacts(local_dir: str, artifact_path: Optional[str] = None) -> None:
    _backend.log_artifacts(local_dir, artifact_path)


def format_key(key: str) -> str:
    """Tracking keys may not contain ``:`` or ``/``."""
    return key.replace(":", "_").replace("/", "_") if key else ""


@optional_mlflow()
def save_text_to_mlflow(content: str, filename: str) -> None:
    if not content:
        return
    logger.info("save file %s to experiment tracking", filename)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, filename)
        with open(path, "w") as f:
            f.write(content)
        _backend.log_artifact(path)


def task_env() -> Dict[str, str]:
    """Environment that lets task processes log into the client's run."""
    if not use_mlflow():
        return {}
    env = {"MLFLOW_RUN_ID": active_run_id(), "MLFLOW_TRACKING_URI": get_tracking_uri(), "GIT_PYTHON_REFRESH": "quiet"}
    if os.environ.get("TFY_TRACKING_DIR"):
        env["TFY_TRACKING_DIR"] = os.environ["TFY_TRACKING_DIR"]
    return env
