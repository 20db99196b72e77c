"""PyTorch front-end (reference: tf_yarn/pytorch/__init__.py:11-25)."""
from tf_yarn_b200.client import RunFailed, get_safe_experiment_fn
from tf_yarn_b200.metrics import Metrics
from tf_yarn_b200.pytorch.client import run_on_yarn
from tf_yarn_b200.pytorch.experiment import DataLoaderArgs, DistributedDataParallelArgs, PytorchExperiment
from tf_yarn_b200.topologies import NodeLabel, TaskSpec

__all__ = ["PytorchExperiment", "DataLoaderArgs", "DistributedDataParallelArgs", "run_on_yarn", "RunFailed",
           "Metrics", "TaskSpec", "NodeLabel", "get_safe_experiment_fn"]
