"""Synchronous all-reduce task (the reference's Horovod-gloo task, on NVLink).

(reference: tf_yarn/tensorflow/tasks/gloo_allred_task.py:31-142, used as
``custom_task_module="tf_yarn.tensorflow.tasks.gloo_allred_task"``)

The chief assigns ranks (chief = 0, then workers by id), reserves the rendezvous
address and publishes both through the KV store; every trainer exports the
``HOROVOD_*`` variables, calls ``hvd.init()`` and runs the experiment:

* KerasExperiment: ``model.fit(**train_params)`` with ``x = input_data_fn()``,
  ``y = target_data_fn()``; non-chief ranks lose their ``ModelCheckpoint`` callbacks.
* Experiment (Estimator): ``estimator.train(input_fn, hooks, max_steps)``; non-chief
  ranks do not write checkpoints / summaries.

The evaluator role runs the evaluator loop.  Unlike the reference (whose KV keys are
built from a NamedTuple repr and keyed by IP address -- SURVEY.md §3.3) rank info is
keyed by task: ``<type>:<id>/rank_info``.
"""
from __future__ import annotations

import logging
import os
from typing import List, Optional

from tf_yarn_b200 import _internal, _task_commons, event
from tf_yarn_b200._task_commons import TaskClient, get_task, get_task_key, setup_logging
from tf_yarn_b200.tensorflow import Experiment, KerasExperiment
from tf_yarn_b200.tensorflow.tasks.evaluator_task import evaluator_fn
from tf_yarn_b200.topologies import ContainerTask

logger = logging.getLogger(__name__)

N_PROCESS_PER_WORKER = 1


def get_net_if() -> str:
    """Interface the collectives use: NVLink needs none; gloo (CPU) rides on loopback on one box."""
    return "lo"


def _trainers(cluster_tasks: List[ContainerTask]) -> List[ContainerTask]:
    chief = [t for t in cluster_tasks if t.type == "chief"]
    workers = sorted((t for t in cluster_tasks if t.type == "worker"), key=lambda t: t.id)
    return chief + workers


def _driver_fn(client, cluster_tasks: List[ContainerTask]) -> None:
    """Chief only: rank assignment + rendezvous address."""
    trainers = _trainers(cluster_tasks)
    size = len(trainers)
    for rank, t in enumerate(trainers):
        # "rank,size,local_rank,local_size,cross_rank,cross_size": one box => local == global
        info = f"{rank},{size},{rank},{size},0,1"
        event.broadcast(client, f"{t.to_container_key().to_kv_str()}/rank_info", info)
    with _internal.reserve_sock_addr() as (host, port):
        event.broadcast(client, "chief:0/sock_addr", f"{host}:{port}")


def _setup_hvd_env(client) -> None:
    task = get_task()
    event.broadcast(client, f"{task}/addr", _internal.local_hostname())
    rank, size, local_rank, local_size, cross_rank, cross_size = \
        event.wait(client, f"{task}/rank_info").split(",")
    addr, port = event.wait(client, "chief:0/sock_addr").split(":")
    os.environ.update({
        "HOROVOD_GLOO_RENDEZVOUS_ADDR": addr, "HOROVOD_GLOO_RENDEZVOUS_PORT": port,
        "HOROVOD_CONTROLLER": "nvlink", "HOROVOD_CPU_OPERATIONS": "gloo", "HOROVOD_GLOO_IFACE": get_net_if(),
        "HOROVOD_RANK": rank, "HOROVOD_SIZE": size, "HOROVOD_LOCAL_RANK": local_rank,
        "HOROVOD_LOCAL_SIZE": local_size, "HOROVOD_CROSS_RANK": cross_rank, "HOROVOD_CROSS_SIZE": cross_size,
        # rendezvous of the NVLink communicator goes through the launcher's KV store
        "TFY_RANK": rank, "TFY_WORLD_SIZE": size,
    })


def _worker_fn(client) -> None:
    from tf_yarn_b200 import hvd
    from tf_yarn_b200.keras.callbacks import ModelCheckpoint
    task = get_task()
    _setup_hvd_env(client)
    hvd.init()
    experiment = _task_commons._get_experiment(client)
    event.start_event(client, task)
    event.broadcast_train_eval_start_timer(client, task)
    if isinstance(experiment, Experiment):
        if not _task_commons.is_chief():
            # only the chief writes checkpoints and summaries
            experiment.estimator._model_dir = "."
            experiment.estimator._config = experiment.estimator._config.replace(
                model_dir=None, save_summary_steps=None, save_checkpoints_steps=None, save_checkpoints_secs=None,
                log_step_count_steps=None)
        logger.info("start training..")
        experiment.estimator.train(experiment.train_spec.input_fn, hooks=list(experiment.train_spec.hooks),
                                   max_steps=experiment.train_spec.max_steps)
    elif isinstance(experiment, KerasExperiment):
        train_params = dict(experiment.train_params)
        if not _task_commons.is_chief():
            train_params["callbacks"] = [cb for cb in train_params.get("callbacks", [])
                                         if not isinstance(cb, ModelCheckpoint)]
        if experiment.input_data_fn:
            train_params["x"] = experiment.input_data_fn()
        if experiment.target_data_fn:
            train_params["y"] = experiment.target_data_fn()
        logger.info("start training..")
        experiment.model.fit(**train_params)
    else:
        raise ValueError("experiment must be an Experiment or a KerasExperiment")
    event.broadcast_train_eval_stop_timer(client, task)
    hvd.shutdown()


def main() -> None:
    setup_logging()
    client = TaskClient.from_current()
    task_key = get_task_key()
    task = task_key.to_kv_str()
    event.init_event(client, task, "127.0.0.1:0")
    _task_commons._setup_container_logs(client)
    cluster_tasks = _task_commons._get_cluster_tasks(client)
    error: Optional[BaseException] = None
    try:
        if task_key.type == "chief":
            _driver_fn(client, cluster_tasks)
        if task_key.type in ("chief", "worker"):
            _worker_fn(client)
        elif task_key.type == "evaluator":
            event.start_event(client, task)
            event.broadcast_train_eval_start_timer(client, task)
            evaluator_fn(client)
            event.broadcast_train_eval_stop_timer(client, task)
        else:
            logger.info("%s: nothing to do", task)
    except Exception as exc:  # noqa: BLE001
        error = exc
    event.stop_event(client, task, error)
    event.broadcast_container_stop_time(client, task)
    if error is not None:
        raise error


if __name__ == "__main__":
    main()
