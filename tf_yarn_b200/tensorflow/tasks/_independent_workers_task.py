"""Default task program: asynchronous parameter-server training of an Estimator experiment.

(reference: tf_yarn/tensorflow/tasks/_independent_workers_task.py:17-43)

reserve a port -> init barrier -> TF_CONFIG -> build the experiment -> run
``train_and_evaluate`` on a monitored thread -> (ps: do not join) -> stop barrier.
"""
from __future__ import annotations

import logging

from tf_yarn_b200 import _internal, _task_commons
from tf_yarn_b200._task_commons import get_task_key
from tf_yarn_b200.tensorflow import KerasExperiment, cluster
from tf_yarn_b200.tensorflow.tasks import tf_task_common

_logger = logging.getLogger(__name__)


def main() -> None:
    _task_commons.setup_logging()
    tf_task_common._log_sys_info()
    task_type, _task_id = get_task_key()
    with _internal.reserve_sock_addr() as host_port:
        client, cluster_spec, cluster_tasks = tf_task_common._prepare_container(host_port)
        cluster.setup_tf_config(cluster_spec)
        experiment = _task_commons._get_experiment(client)
        if isinstance(experiment, KerasExperiment):
            # the reference raises here without telling anybody (the task shows up as KILLED);
            # publish the reason so the client's report carries it
            err = ValueError("KerasExperiment using parameter strategy is unsupported")
            from tf_yarn_b200 import event
            event.start_event(client, _task_commons.get_task())
            event.stop_event(client, _task_commons.get_task(), err)
            raise err
        session_config = experiment.config.session_config
    cluster.start_tf_server(cluster_spec, session_config)
    thread = tf_task_common._execute_dispatched_function(client, experiment)
    # "ps" tasks do not terminate by themselves: they leave through the stop barrier
    if task_type != "ps":
        thread.join()
        _logger.info("%s:%d %s", task_type, _task_id, thread.state)
    tf_task_common._shutdown_container(client, cluster_tasks, session_config, thread)


if __name__ == "__main__":
    main()
