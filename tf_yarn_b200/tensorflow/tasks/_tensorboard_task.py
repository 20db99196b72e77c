"""TensorBoard side task (reference: tf_yarn/tensorflow/tasks/_tensorboard_task.py:26-62)."""
from __future__ import annotations

import logging
import os

from tf_yarn_b200 import _internal, _task_commons, event, tensorboard
from tf_yarn_b200._task_commons import TaskClient, get_task
from tf_yarn_b200.tensorflow.tasks.tf_task_common import _wait_for_connected_tasks

_logger = logging.getLogger(__name__)


def _model_dir(client) -> str:
    if os.environ.get("TB_MODEL_DIR"):
        return os.environ["TB_MODEL_DIR"]
    experiment = _task_commons._get_experiment(client)
    if hasattr(experiment, "estimator"):
        return experiment.estimator.config.model_dir
    if hasattr(experiment, "model_dir"):
        return experiment.model_dir
    if getattr(experiment, "tensorboard_hdfs_dir", None):
        return experiment.tensorboard_hdfs_dir
    raise ValueError("cannot find the model directory: set TaskSpec(tb_model_dir=...)")


def main() -> None:
    _task_commons.setup_logging()
    client = TaskClient.from_current()
    task = get_task()
    event.init_event(client, task, "127.0.0.1:0")
    _task_commons._setup_container_logs(client)
    cluster_tasks = _task_commons._get_cluster_tasks(client)
    model_dir = _model_dir(client)
    os.makedirs(model_dir, exist_ok=True)
    thread = _internal.MonitoredThread(name=task, target=tensorboard.start_tf_board, args=(client, model_dir),
                                       daemon=True)
    thread.start()
    # run until every training task has stopped, then linger so the last events can be read
    _wait_for_connected_tasks(client, cluster_tasks, [])
    timeout = tensorboard.get_termination_timeout()
    _logger.info("all tasks stopped; tensorboard stays up %d more seconds", timeout)
    thread.join(timeout)
    import time
    if timeout > 0 and not thread.is_alive():
        time.sleep(timeout)
    event.stop_event(client, task, thread.exception)
    event.broadcast_container_stop_time(client, task)


if __name__ == "__main__":
    main()
