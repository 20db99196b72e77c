"""LinearClassifier on the wine-quality data: chief + evaluator + tensorboard.

(reference: tf_yarn/examples/linear_classifier_example.py)
"""
import logging
import os
import tempfile
from datetime import datetime

from tf_yarn_b200 import estimator as est
from tf_yarn_b200.examples import winequality
from tf_yarn_b200.tensorflow import Experiment, TaskSpec, run_on_yarn

logging.basicConfig(level="INFO")

WINE_QUALITY_FILE = os.path.join(tempfile.gettempdir(), "tf_yarn_b200_test", "winequality-red.csv")
MODEL_DIR = os.path.join(tempfile.gettempdir(), "tf_yarn_b200_test", f"linear_{int(datetime.now().timestamp())}")


def experiment_fn() -> Experiment:
    def train_input_fn():
        return winequality.get_dataset(WINE_QUALITY_FILE, split="train").shuffle(1000).batch(128).repeat()

    def eval_input_fn():
        return winequality.get_dataset(WINE_QUALITY_FILE, split="test").shuffle(1000).batch(128)

    estimator = est.LinearClassifier(feature_columns=winequality.get_feature_columns(), model_dir=MODEL_DIR,
                                     n_classes=winequality.get_n_classes(),
                                     config=est.RunConfig(save_checkpoints_steps=50))
    return Experiment(estimator, est.TrainSpec(train_input_fn, max_steps=100),
                      est.EvalSpec(eval_input_fn, steps=10, start_delay_secs=0, throttle_secs=5))


def main():
    winequality.ensure_dataset(WINE_QUALITY_FILE)
    return run_on_yarn(experiment_fn, task_specs={
        "chief": TaskSpec(memory="2 GiB", vcores=4),
        "evaluator": TaskSpec(memory="2 GiB", vcores=1),
        "tensorboard": TaskSpec(memory="2 GiB", vcores=1, tb_model_dir=MODEL_DIR, tb_termination_timeout_seconds=5),
    })


if __name__ == "__main__":
    print(main())
