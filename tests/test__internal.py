import os
import socket

import pytest

from tf_yarn_b200._internal import MonitoredThread, iter_tasks, reserve_sock_addr, xset_environ
from tf_yarn_b200.topologies import ContainerTask


def test_monitored_thread_captures_exception():
    def fail():
        raise RuntimeError("boom")
    t = MonitoredThread(target=fail)
    t.start()
    t.join()
    assert isinstance(t.exception, RuntimeError) and t.state == "FAILED"


def test_monitored_thread_success():
    t = MonitoredThread(target=lambda: None)
    t.start()
    t.join()
    assert t.exception is None and t.state == "SUCCEEDED"


def test_reserve_sock_addr_holds_the_port():
    with reserve_sock_addr() as (host, port):
        assert host and port > 0
        # the port is taken while the context is open (a plain bind without SO_REUSEPORT must fail)
        with socket.socket() as s:
            with pytest.raises(OSError):
                s.bind(("", port))
    with socket.socket() as s:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(("", port))


def test_xset_environ(monkeypatch):
    monkeypatch.delenv("TFY_TEST_FOO", raising=False)
    xset_environ(TFY_TEST_FOO="bar")
    assert os.environ["TFY_TEST_FOO"] == "bar"
    with pytest.raises(RuntimeError):
        xset_environ(TFY_TEST_FOO="baz")
    assert os.environ["TFY_TEST_FOO"] == "bar"
    os.environ.pop("TFY_TEST_FOO", None)


def test_iter_tasks():
    tasks = list(iter_tasks([("chief", 1, 1), ("worker", 2, 4)]))
    assert tasks == [ContainerTask("chief", 0, 1), ContainerTask("worker", 0, 4), ContainerTask("worker", 1, 4)]
