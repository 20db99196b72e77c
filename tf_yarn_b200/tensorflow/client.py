"""``run_on_yarn`` for Estimator / Keras experiments (reference: tf_yarn/tensorflow/client.py:17-31)."""
from typing import Callable, Dict, Union

from tf_yarn_b200 import client
from tf_yarn_b200.tensorflow.experiment import Experiment
from tf_yarn_b200.tensorflow.keras_experiment import KerasExperiment
from tf_yarn_b200.topologies import TaskSpec, single_server_topology

ExperimentFn = Callable[[], Union[Experiment, KerasExperiment]]


def _wrap_with_monitors(experiment_fn: ExperimentFn) -> ExperimentFn:
    def _new_experiment_fn():
        from tf_yarn_b200.tensorflow import metrics
        return metrics._add_monitor_to_experiment(experiment_fn())
    return _new_experiment_fn


def run_on_yarn(experiment_fn: ExperimentFn, task_specs: Dict[str, TaskSpec] = None, *args, **kwargs):
    """Run an ``Experiment`` / ``KerasExperiment``; monitoring hooks are injected inside the tasks."""
    if task_specs is None:
        task_specs = single_server_topology()
    return client.run_on_yarn(_wrap_with_monitors(experiment_fn), task_specs, *args, **kwargs)
