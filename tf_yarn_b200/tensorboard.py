"""TensorBoard side service (reference: tf_yarn/tensorboard.py:15-58).

Runs inside the ``tensorboard`` task: launches TensorBoard on a reserved port
over the model directory and advertises ``<task>/url`` through the KV store.
"""
from __future__ import annotations

import logging
import os
import shlex
from typing import Iterable, Optional

from tf_yarn_b200 import _internal, event
from tf_yarn_b200._task_commons import get_task
from tf_yarn_b200.topologies import ContainerTask

_logger = logging.getLogger(__name__)

DEFAULT_TERMINATION_TIMEOUT_SECONDS = 30
URL_EVENT_LABEL = "Tensorboard listening on"


def get_termination_timeout() -> int:
    timeout = os.environ.get("TB_TERMINATION_TIMEOUT_SECONDS")
    return int(timeout) if timeout is not None else DEFAULT_TERMINATION_TIMEOUT_SECONDS


def start_tf_board(client, tf_board_model_dir: str) -> Optional[str]:
    """Launch TensorBoard in-process (its own server thread); returns the URL or None."""
    task = get_task()
    try:
        from tensorboard import program
        if hasattr(program, "setup_environment"):       # removed from recent TensorBoard releases
            program.setup_environment()
        board = program.TensorBoard()
        with _internal.reserve_sock_addr() as (host, port):
            url = f"http://{host}:{port}"
            argv = ["tensorboard", f"--logdir={tf_board_model_dir}", f"--port={port}", f"--host={host}"]
            extra = os.getenv("TB_EXTRA_ARGS", "")
            if extra:
                argv += shlex.split(extra)
            board.configure(argv)
        board.launch()
        event.start_event(client, task)
        event.url_event(client, task, url)
        return url
    except Exception as exc:  # noqa: BLE001
        _logger.error("Cannot start tensorboard: %s", exc, exc_info=True)
        event.stop_event(client, task, exc)
        return None


def url_event_name(tasks: Iterable[ContainerTask]) -> Optional[str]:
    """``"tensorboard:0/url"`` iff the job has exactly one tensorboard task."""
    boards = [t for t in tasks if t.type == "tensorboard"]
    if len(boards) == 1:
        return boards[0].to_container_key().to_kv_str() + "/url"
    return None
