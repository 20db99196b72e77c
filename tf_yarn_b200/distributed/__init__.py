"""Framework-agnostic distributed tasks: the launcher supplies rank / world / master only."""
from tf_yarn_b200.distributed.client import run_on_yarn  # noqa: F401
from tf_yarn_b200.distributed.task import TaskParameters, get_task  # noqa: F401
