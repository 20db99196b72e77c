"""``run_on_yarn`` for PytorchExperiment (reference: tf_yarn/pytorch/client.py:12-23)."""
from typing import Callable, Dict

from tf_yarn_b200 import client
from tf_yarn_b200.pytorch.experiment import PytorchExperiment
from tf_yarn_b200.topologies import TaskSpec

TASK_MODULE = "tf_yarn_b200.pytorch.tasks.worker"


def run_on_yarn(experiment_fn: Callable[[], PytorchExperiment], task_specs: Dict[str, TaskSpec], **kwargs):
    kwargs.setdefault("custom_task_module", TASK_MODULE)
    return client.run_on_yarn(experiment_fn, task_specs, **kwargs)
