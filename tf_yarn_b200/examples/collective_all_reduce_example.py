"""Estimator trained with the synchronous all-reduce task (reference:
tf_yarn/examples/collective_all_reduce_example.py:58-94)."""
import logging
import os
import tempfile
from datetime import datetime

import torch

from tf_yarn_b200 import estimator as est
from tf_yarn_b200 import hvd, keras
from tf_yarn_b200.examples import winequality
from tf_yarn_b200.tensorflow import Experiment, NodeLabel, TaskSpec, run_on_yarn

logging.basicConfig(level="INFO")

WINE_QUALITY_FILE = os.path.join(tempfile.gettempdir(), "tf_yarn_b200_test", "winequality-red.csv")
MODEL_DIR = os.path.join(tempfile.gettempdir(), "tf_yarn_b200_test", f"hvd_est_{int(datetime.now().timestamp())}")
LABEL = NodeLabel.GPU if torch.cuda.is_available() else NodeLabel.CPU


def experiment_fn() -> Experiment:
    def train_input_fn():
        return winequality.get_dataset(WINE_QUALITY_FILE, split="train").shuffle(1000).batch(128).repeat()

    def eval_input_fn():
        return winequality.get_dataset(WINE_QUALITY_FILE, split="test").shuffle(1000).batch(128)

    estimator = est.LinearClassifier(
        feature_columns=winequality.get_feature_columns(), model_dir=MODEL_DIR,
        n_classes=winequality.get_n_classes(),
        optimizer=lambda: hvd.DistributedOptimizer(keras.optimizers.Adam()),
        config=est.RunConfig(save_checkpoints_steps=5))
    return Experiment(estimator,
                      est.TrainSpec(train_input_fn, max_steps=10, hooks=[hvd.BroadcastGlobalVariablesHook(0)]),
                      est.EvalSpec(eval_input_fn, steps=10, start_delay_secs=0, throttle_secs=2))


def main():
    winequality.ensure_dataset(WINE_QUALITY_FILE)
    return run_on_yarn(
        experiment_fn,
        task_specs={
            "chief": TaskSpec(memory="2 GiB", vcores=4, label=LABEL),
            "worker": TaskSpec(memory="2 GiB", vcores=4, instances=1, label=LABEL),
            "evaluator": TaskSpec(memory="2 GiB", vcores=1),
            "tensorboard": TaskSpec(memory="2 GiB", vcores=1, tb_model_dir=MODEL_DIR,
                                    tb_termination_timeout_seconds=5),
        },
        custom_task_module="tf_yarn_b200.tensorflow.tasks.gloo_allred_task")


if __name__ == "__main__":
    print(main())
