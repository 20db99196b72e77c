"""``run_on_yarn`` for arbitrary ``fn(local_rank)`` jobs (reference: tf_yarn/distributed/client.py:9-20)."""
from typing import Callable, Dict

from tf_yarn_b200 import client
from tf_yarn_b200.topologies import TaskSpec

TASK_MODULE = "tf_yarn_b200.distributed.task"


def run_on_yarn(experiment_fn: Callable[[int], None], task_specs: Dict[str, TaskSpec], **kwargs):
    """Run ``experiment_fn(local_rank)`` in every process of every task instance.

    Inside the function, :func:`tf_yarn_b200.distributed.task.get_task` returns the
    global rank, world size and the master address chosen through the KV store.
    """
    kwargs.setdefault("custom_task_module", TASK_MODULE)
    return client.run_on_yarn(experiment_fn, task_specs, **kwargs)
