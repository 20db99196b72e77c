"""Experiment tracking: run metrics, evaluator statistics, durations and task logs end up in the tracker.

(reference: tf_yarn/examples/mlflow_example.py:61-119, docs/MLflow.md)

With the real ``mlflow`` package installed and ``MLFLOW_TRACKING_URI`` set, MLflow is used.  On an
air-gapped box set ``TFY_TRACKING_DIR`` (done below) to use the built-in file tracker.
"""
import logging
import os
import tempfile
from datetime import datetime

from tf_yarn_b200 import estimator as est
from tf_yarn_b200 import mlflow
from tf_yarn_b200.examples import winequality
from tf_yarn_b200.tensorflow import Experiment, TaskSpec, run_on_yarn

logging.basicConfig(level="INFO")

ROOT = os.path.join(tempfile.gettempdir(), "tf_yarn_b200_test")
WINE_QUALITY_FILE = os.path.join(ROOT, "winequality-red.csv")
MODEL_DIR = os.path.join(ROOT, f"mlflow_{int(datetime.now().timestamp())}")


def experiment_fn() -> Experiment:
    def train_input_fn():
        return winequality.get_dataset(WINE_QUALITY_FILE, split="train").shuffle(1000).batch(128).repeat()

    def eval_input_fn():
        return winequality.get_dataset(WINE_QUALITY_FILE, split="test").shuffle(1000).batch(128)

    estimator = est.LinearClassifier(winequality.get_feature_columns(), model_dir=MODEL_DIR,
                                     n_classes=winequality.get_n_classes(),
                                     config=est.RunConfig(save_checkpoints_steps=25, log_step_count_steps=10))
    return Experiment(estimator, est.TrainSpec(train_input_fn, max_steps=100),
                      est.EvalSpec(eval_input_fn, steps=10, start_delay_secs=0, throttle_secs=1))


def main():
    os.environ.setdefault("TFY_TRACKING_DIR", os.path.join(ROOT, "tracking"))
    mlflow.reset()
    winequality.ensure_dataset(WINE_QUALITY_FILE)
    run_on_yarn(experiment_fn, task_specs={"chief": TaskSpec(memory="2 GiB", vcores=4),
                                           "evaluator": TaskSpec(memory="2 GiB", vcores=1)})
    tracker = mlflow.backend()
    metrics = tracker.list("metrics") if hasattr(tracker, "list") else []
    print("tracked metrics:", metrics)
    assert any(m.startswith("steps_per_sec") for m in metrics), metrics
    assert any("evaluator" in m for m in metrics), metrics
    return metrics


if __name__ == "__main__":
    main()
