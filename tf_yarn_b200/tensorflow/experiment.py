"""Estimator experiment descriptor (reference: tf_yarn/tensorflow/experiment.py:6-13)."""
from typing import NamedTuple

from tf_yarn_b200.estimator import Estimator, EvalSpec, RunConfig, TrainSpec


class Experiment(NamedTuple):
    estimator: Estimator
    train_spec: TrainSpec
    eval_spec: EvalSpec

    @property
    def config(self) -> RunConfig:
        return self.estimator.config
