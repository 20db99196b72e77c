"""The smallest possible experiment: a parameter-free estimator, one chief, a shared client.

(reference: tf_yarn/examples/id_estimator_example.py:14-59)
"""
import logging

import torch

from tf_yarn_b200 import data
from tf_yarn_b200 import estimator as est
from tf_yarn_b200.launcher.local import LocalClient
from tf_yarn_b200.tensorflow import Experiment, TaskSpec, run_on_yarn

logging.basicConfig(level="INFO")


def model_fn(features, labels, mode):
    # identity: no network, no variables; a training step only advances the global step
    return est.EstimatorSpec(mode=mode, network=None, predictions=lambda x: {"x": x["x"]})


def experiment_fn() -> Experiment:
    def input_fn():
        return data.Dataset.from_tensor_slices(({"x": torch.tensor([[1.0], [2.0]])}, torch.tensor([0, 1]))).batch(2)
    estimator = est.Estimator(model_fn=model_fn)
    return Experiment(estimator, est.TrainSpec(input_fn, max_steps=1),
                      est.EvalSpec(input_fn, steps=1, start_delay_secs=0, throttle_secs=0))


def main():
    # several runs can share one client, as the reference shares a skein.Client
    with LocalClient() as client:
        return run_on_yarn(experiment_fn, task_specs={"chief": TaskSpec(memory="1 GiB", vcores=1)},
                           skein_client=client)


if __name__ == "__main__":
    print(main())
