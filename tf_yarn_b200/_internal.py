"""Small concurrency / networking utilities shared by launcher and tasks.

Behavioural parity with the reference (reference: tf_yarn/_internal.py:22-96).
"""
from __future__ import annotations

import os
import socket
import threading
from contextlib import contextmanager
from typing import Iterable, Iterator, List, Optional, Tuple

from tf_yarn_b200.topologies import ContainerTask


class MonitoredThread(threading.Thread):
    """Thread that remembers the exception that ended ``target``."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._exc: Optional[BaseException] = None

    @property
    def exception(self) -> Optional[BaseException]:
        return self._exc

    @property
    def state(self) -> str:
        if self.is_alive():
            return "RUNNING"
        return "FAILED" if self._exc is not None else "SUCCEEDED"

    def run(self) -> None:
        try:
            super().run()
        except Exception as exc:  # noqa: BLE001 - that is the point
            self._exc = exc


def get_so_reuseport() -> Optional[int]:
    return getattr(socket, "SO_REUSEPORT", None)


def local_hostname() -> str:
    """The address peers on the same box use; one-box jobs always rendezvous on loopback."""
    return os.environ.get("TFY_HOST", "127.0.0.1")


@contextmanager
def reserve_sock_addr() -> Iterator[Tuple[str, int]]:
    """Bind port 0 with SO_REUSEPORT and keep it bound while the context is open.

    The holder hands the port to a server started right after leaving the
    context, which minimises the window in which another process may grab it.
    """
    so_reuseport = get_so_reuseport()
    if so_reuseport is None:
        raise RuntimeError("SO_REUSEPORT is not supported by the operating system")
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
        sock.setsockopt(socket.SOL_SOCKET, so_reuseport, 1)
        sock.bind(("", 0))
        _, port = sock.getsockname()
        yield local_hostname(), port


def iter_tasks(tasks: List[Tuple[str, int, int]]) -> Iterable[ContainerTask]:
    """Expand ``[(type, n_instances, nb_proc), ...]`` into one ContainerTask per instance."""
    for task_type, n_instances, nb_proc in tasks:
        for task_id in range(n_instances):
            yield ContainerTask(task_type, task_id, nb_proc)


def xset_environ(**kwargs: str) -> None:
    """Set environment variables, refusing to overwrite any that already exist."""
    for key, value in kwargs.items():
        if key in os.environ:
            raise RuntimeError(f"{key} already set in os.environ: {value}")
    os.environ.update(kwargs)
