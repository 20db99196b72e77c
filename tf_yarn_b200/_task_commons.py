"""Helpers shared by every task program (what runs inside a task process).

Semantics follow the reference (reference: tf_yarn/_task_commons.py:19-125)
with its latent defects fixed: ranks are collision-free for mixed role types
 and the task identity comes from ``TFY_TASK_KEY`` (the legacy
``SKEIN_CONTAINER_ID="worker_0"`` form is still understood).
"""
from __future__ import annotations

import json
import logging
import logging.config
import os
import time
from contextlib import contextmanager
from typing import Any, List, Optional, Tuple

import cloudpickle

from tf_yarn_b200 import constants, event
from tf_yarn_b200._internal import iter_tasks, reserve_sock_addr
from tf_yarn_b200.topologies import ContainerKey, ContainerTask

_logger = logging.getLogger(__name__)


def setup_logging() -> None:
    here = os.path.dirname(__file__)
    logging.config.fileConfig(os.path.join(here, "default.log.conf"), disable_existing_loggers=False)
    exit_cleanly_on_sigterm()


def exit_cleanly_on_sigterm() -> None:
    """Turn the launcher's SIGTERM (application finished, failed or killed: launcher/local.py ``_kill_all``) into a
    normal interpreter exit, so ``atexit`` hooks and ``finally`` blocks run: the parameter-server shards unlink
    their /dev/shm files, KV connections close.  (YARN gives the reference's containers the same SIGTERM-then-
    SIGKILL sequence.)  Only the main thread of the main interpreter can install it; the exit status stays 143."""
    import signal
    import threading
    if threading.current_thread() is not threading.main_thread():
        return

    def _term(signum, frame):
        raise SystemExit(128 + signum)

    try:
        signal.signal(signal.SIGTERM, _term)
    except (ValueError, OSError):          # embedded interpreter / exotic platform: keep the default action
        pass


class TaskClient:
    """What a task sees of the application: its KV store (skein ``ApplicationClient`` stand-in)."""

    def __init__(self, kv=None):
        if kv is None:
            from tf_yarn_b200.kv import KVClient
            kv = KVClient()
        self.kv = kv

    @classmethod
    def from_current(cls) -> "TaskClient":
        return cls()


def get_task_key() -> ContainerKey:
    """Identity of the current task: ``ContainerKey("worker", 0)``."""
    raw = os.environ.get("TFY_TASK_KEY")
    if raw:
        key = ContainerKey.from_kv_str(raw)
        if key is not None:
            return key
    raw = os.environ.get("SKEIN_CONTAINER_ID")
    if raw:
        task_type, task_id = raw.rsplit("_", 1)
        return ContainerKey(task_type, int(task_id))
    raise RuntimeError("not running inside a tf_yarn_b200 task (TFY_TASK_KEY is not set)")


def get_task() -> str:
    return get_task_key().to_kv_str()


def n_try() -> int:
    return int(os.environ.get("TF_YARN_N_TRY", os.environ.get("TFY_N_TRY", "0")))


def is_worker(task_type: Optional[str] = None) -> bool:
    return (task_type or get_task_key().type) == "worker"


def is_evaluator(task_type: Optional[str] = None) -> bool:
    return (task_type or get_task_key().type) == "evaluator"


def is_chief(task_type: Optional[str] = None) -> bool:
    return (task_type or get_task_key().type) == "chief"


def _setup_container_logs(client) -> None:
    task = get_task()
    event.broadcast_container_start_time(client, task)
    event.logs_event(client, task, os.environ.get("TFY_LOG_FILE", ""))


def _get_cluster_tasks(client) -> List[ContainerTask]:
    """Tasks that form the training cluster (evaluator / tensorboard excluded by the launcher)."""
    raw = client.kv.wait(constants.KV_CLUSTER_INSTANCES)
    raw = raw.decode() if isinstance(raw, (bytes, bytearray)) else raw
    return list(iter_tasks(json.loads(raw)))


def _compute_world_size(cluster_tasks: List[ContainerTask]) -> int:
    return sum(task.nb_proc for task in cluster_tasks)


def _get_nb_workers(task_id: int, cluster_tasks: List[ContainerTask]) -> int:
    """Processes per instance for the task with index ``task_id``."""
    return [task.nb_proc for task in cluster_tasks if task.id == task_id][0]


def get_pickled_experiment(client) -> bytes:
    return client.kv.wait(constants.KV_EXPERIMENT_FN)


def _get_experiment(client) -> Any:
    """Unpickle the experiment function and CALL it; failures are reported as start+stop(exc)."""
    try:
        experiment = cloudpickle.loads(get_pickled_experiment(client))()
    except Exception as e:
        task = get_task()
        event.start_event(client, task)
        event.stop_event(client, task, e)
        raise
    return experiment


def get_pickled_fn(client):
    """The raw pickled callable (``tf_yarn_b200.distributed`` ships ``fn(local_rank)`` itself)."""
    return cloudpickle.loads(get_pickled_experiment(client))


# ---------------------------------------------------------------------------
# ranks
# ---------------------------------------------------------------------------
# fixed role order gives every process of the training cluster a unique global rank even
# when chief and workers coexist (the reference's task_id*n+local_rank collides there)
_ROLE_ORDER = {"chief": 0, "worker": 1, "ps": 2}


def rank_table(cluster_tasks: List[ContainerTask], roles: Optional[Tuple[str, ...]] = None):
    """``{(type, id, local_rank): global_rank}`` over the given roles (default: all cluster roles)."""
    ordered = sorted((t for t in cluster_tasks if roles is None or t.type in roles),
                     key=lambda t: (_ROLE_ORDER.get(t.type, 99), t.type, t.id))
    table = {}
    rank = 0
    for t in ordered:
        for local in range(t.nb_proc):
            table[(t.type, t.id, local)] = rank
            rank += 1
    return table


def compute_rank(task_id: int, local_rank: int, n_workers_per_executor: int) -> int:
    """Reference formula (valid for single-role topologies); prefer :func:`rank_table`."""
    return task_id * n_workers_per_executor + local_rank


def choose_master(client, rank: int, key_prefix: str = "") -> Tuple[str, int]:
    """Rank 0 picks a free port and publishes MASTER_ADDR/MASTER_PORT; the others wait for it."""
    ka, kp = f"{key_prefix}MASTER_ADDR", f"{key_prefix}MASTER_PORT"
    if rank == 0:
        with reserve_sock_addr() as (host, port):
            event.broadcast(client, ka, host)
            event.broadcast(client, kp, str(port))
            return host, port
    return event.wait(client, ka), int(event.wait(client, kp))


@contextmanager
def catchtime(label: str = ""):
    start = time.perf_counter()
    yield
    _logger.info("%s took %.3f s", label or "step", time.perf_counter() - start)
