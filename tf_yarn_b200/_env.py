"""Command line of a task program (reference: tf_yarn/_env.py:6-24)."""
from __future__ import annotations

import sys
from typing import Optional

from tf_yarn_b200.topologies import ALL_TASK_TYPES

DEFAULT_TASK_MODULE = "tf_yarn_b200.tensorflow.tasks._independent_workers_task"
TENSORBOARD_TASK_MODULE = "tf_yarn_b200.tensorflow.tasks._tensorboard_task"


def gen_task_module(task_type: str, custom_task_module: Optional[str] = None) -> str:
    if task_type == "tensorboard":
        return TENSORBOARD_TASK_MODULE
    if task_type in ALL_TASK_TYPES:
        return custom_task_module or DEFAULT_TASK_MODULE
    raise ValueError(f"task_type: {task_type} is not handled")


def gen_task_cmd(interpreter: Optional[str], task_type: str, custom_task_module: Optional[str] = None) -> str:
    """``"<python> -m <module> "`` for the given role."""
    return f"{interpreter or sys.executable} -m {gen_task_module(task_type, custom_task_module)} "
