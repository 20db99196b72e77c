"""Monitoring hooks injected into Estimator experiments (reference: tf_yarn/tensorflow/metrics.py:18-142)."""
from __future__ import annotations

import logging
import time
from typing import Dict, List, Union

from tf_yarn_b200 import mlflow
from tf_yarn_b200._task_commons import TaskClient, get_task, is_chief, n_try
from tf_yarn_b200.estimator import SessionRunArgs, SessionRunHook, StepCounterHook, get_global_step
from tf_yarn_b200.estimator import summary as summary_lib
from tf_yarn_b200.event import broadcast
from tf_yarn_b200.tensorflow import experiment, keras_experiment

logger = logging.getLogger(__name__)


class StepPerSecondHook(StepCounterHook):
    """Chief-only: log ``steps_per_sec_<n_try>`` to the experiment tracker."""

    def __init__(self, every_n_steps=100, every_n_secs=None, output_dir=None, summary_writer=None):
        super().__init__(every_n_steps=every_n_steps, every_n_secs=every_n_secs, output_dir=output_dir,
                         summary_writer=summary_writer)

    def _log_and_record(self, elapsed_steps: int, elapsed_time: float, global_step: int):
        steps_per_sec = elapsed_steps / max(elapsed_time, 1e-9)
        self.last_steps_per_sec = steps_per_sec
        if is_chief():
            mlflow.log_metric(f"steps_per_sec_{n_try()}", steps_per_sec, step=global_step)


class EvalMonitorHook(SessionRunHook):
    """Publishes evaluator liveness statistics to the KV store after every evaluation step.

    Usage: ``EvalSpec(..., hooks=[EvalMonitorHook()])`` (added automatically by ``run_on_yarn``).
    """

    def __init__(self, client=None):
        self._client = client
        self.task = None
        self.step_counter = 0
        self.eval_start_time = 0.0
        self.eval_step_dur_accu = 0.0
        self.start_time = time.time()

    @property
    def client(self):
        if self._client is None:
            self._client = TaskClient.from_current()
        return self._client

    def before_run(self, run_context):
        self.eval_start_time = time.time()
        return SessionRunArgs(get_global_step())

    def after_run(self, _run_context, run_values):
        self.step_counter += 1
        cur_time = time.time()
        self.eval_step_dur_accu += cur_time - self.eval_start_time
        self.broadcast("eval_step_mean_duration", str(self.eval_step_dur_accu / self.step_counter))
        self.broadcast("awake_time_ratio", str(self.eval_step_dur_accu / max(cur_time - self.start_time, 1e-9)))
        self.broadcast("nb_eval_steps", str(self.step_counter))
        self.broadcast("last_training_step", str(run_values.results))

    def broadcast(self, key: str, value: str):
        if self.task is None:
            self.task = get_task()
        broadcast(self.client, f"{self.task}/{key}", value)

    def __getstate__(self):           # the experiment is cloudpickled on the client: drop live handles
        d = dict(self.__dict__)
        d["_client"] = None
        return d


def get_all_metrics(model_path: str) -> Dict[str, List]:
    """Scalars of the event files under ``model_path`` as ``{'step', 'name', 'value'}`` lists."""
    return summary_lib.read_scalars(model_path)


def _hook_name_already_exists(hook, hooks) -> bool:
    return any(type(h).__name__ == type(hook).__name__ for h in hooks)


def _add_monitor_to_experiment(my_experiment: Union[experiment.Experiment, keras_experiment.KerasExperiment]):
    if isinstance(my_experiment, experiment.Experiment):
        logger.info("configured training hooks: %s", my_experiment.train_spec.hooks)
        training_hooks = list(my_experiment.train_spec.hooks)
        if my_experiment.config.log_step_count_steps is not None:
            hook = StepPerSecondHook(every_n_steps=my_experiment.config.log_step_count_steps)
            if not _hook_name_already_exists(hook, training_hooks):
                training_hooks.append(hook)
            else:
                logger.warning("do not add StepPerSecondHook as there is already one configured")
        train_spec = my_experiment.train_spec._replace(hooks=training_hooks)
        eval_spec = my_experiment.eval_spec._replace(hooks=(EvalMonitorHook(), *my_experiment.eval_spec.hooks))
        my_experiment = my_experiment._replace(eval_spec=eval_spec, train_spec=train_spec)
    elif isinstance(my_experiment, keras_experiment.KerasExperiment):
        logger.debug("KerasExperiment: steps/sec is reported by the train engine, no hook injected")
    else:
        raise ValueError("experiment must be an Experiment or a KerasExperiment")
    return my_experiment
