"""Pieces shared by the Estimator task programs (reference: tf_yarn/tensorflow/tasks/tf_task_common.py:21-118)."""
from __future__ import annotations

import logging
import re
import sys
from typing import List, Optional, Tuple, Union

from tf_yarn_b200 import event
from tf_yarn_b200._internal import MonitoredThread
from tf_yarn_b200._task_commons import TaskClient, _get_cluster_tasks, _setup_container_logs, get_task
from tf_yarn_b200.tensorflow import Experiment, KerasExperiment, cluster
from tf_yarn_b200.topologies import ContainerTask

_logger = logging.getLogger(__name__)


def _prepare_container(host_port: Tuple[str, int]):
    """Connect to the KV store, publish logs/start time, run the INIT barrier (socket stays reserved)."""
    client = TaskClient.from_current()
    _setup_container_logs(client)
    cluster_tasks = _get_cluster_tasks(client)
    cluster_spec = cluster.start_cluster(host_port, client, cluster_tasks)
    return client, cluster_spec, cluster_tasks


def _log_sys_info() -> None:
    import torch
    _logger.info("Python %s", sys.version)
    _logger.info("torch %s (cuda: %s)", torch.__version__, torch.cuda.is_available())


def _gen_monitored_train_and_evaluate(client):
    task = get_task()

    def train_and_evaluate(estimator, train_spec, eval_spec):
        from tf_yarn_b200.estimator import train_and_evaluate as tae
        event.broadcast_train_eval_start_timer(client, task)
        tae(estimator, train_spec, eval_spec)
        event.broadcast_train_eval_stop_timer(client, task)

    return train_and_evaluate


def _execute_dispatched_function(client, experiment: Union[Experiment, KerasExperiment]) -> MonitoredThread:
    task = get_task()
    _logger.info("Starting execution %s", task)
    if isinstance(experiment, Experiment):
        thread = MonitoredThread(name=task, target=_gen_monitored_train_and_evaluate(client),
                                 args=tuple(experiment), daemon=True)
    elif isinstance(experiment, KerasExperiment):
        raise ValueError("KerasExperiment using parameter strategy is unsupported")
    else:
        raise ValueError("experiment must be an Experiment or a KerasExperiment")
    thread.start()
    event.start_event(client, task)
    return thread


def _shutdown_container(client, cluster_tasks: List[ContainerTask], session_config,
                        thread: Optional[MonitoredThread]) -> None:
    """Publish ``stop`` (with the training exception, if any), then wait for ``stop`` of every
    connected task -- the STOP barrier that lets never-ending ps tasks leave -- and re-raise."""
    exception = thread.exception if isinstance(thread, MonitoredThread) else None
    task = get_task()
    event.stop_event(client, task, exception)
    _wait_for_connected_tasks(client, cluster_tasks, getattr(session_config, "device_filters", []) or [])
    event.broadcast_container_stop_time(client, task)
    if exception is not None:
        raise exception from None


def _wait_for_connected_tasks(client, all_tasks: List[ContainerTask], device_filters, message: str = "stop"):
    for task in all_tasks:
        if _matches_device_filters(task, device_filters):
            event.wait(client, f"{task.to_container_key().to_kv_str()}/{message}")


def _matches_device_filters(task: ContainerTask, device_filters: List[str]) -> bool:
    """``/job:ps`` matches every ps task, ``/job:worker/task:42`` one task; no filters match everything."""
    for device_filter in device_filters:
        found = re.findall(r"^/job:([a-z]+)(?:/task:(\d+))?$", device_filter.replace("master", "chief"))
        if not found:
            continue
        filter_type, filter_id = found[0]
        if filter_type == task.type and (not filter_id or filter_id == str(task.id)):
            return True
    return not device_filters
