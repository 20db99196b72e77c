"""Estimator / Keras front-end (reference: tf_yarn/tensorflow/__init__.py:1-15).

TensorFlow itself is not needed: ``Experiment`` wraps a :mod:`tf_yarn_b200.estimator`
Estimator, ``KerasExperiment`` a :mod:`tf_yarn_b200.keras` model.
"""
from tf_yarn_b200.client import RunFailed, get_safe_experiment_fn
from tf_yarn_b200.metrics import Metrics
from tf_yarn_b200.tensorflow.experiment import Experiment
from tf_yarn_b200.tensorflow.keras_experiment import KerasExperiment
from tf_yarn_b200.tensorflow.client import run_on_yarn
from tf_yarn_b200.topologies import NodeLabel, TaskSpec, ps_strategy_topology, single_server_topology

__all__ = ["Experiment", "KerasExperiment", "run_on_yarn", "RunFailed", "Metrics", "TaskSpec", "NodeLabel",
           "single_server_topology", "ps_strategy_topology", "get_safe_experiment_fn"]
