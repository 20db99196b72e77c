import json

import cloudpickle
import pytest

from tf_yarn_b200 import _task_commons, constants
from tf_yarn_b200.topologies import ContainerKey, ContainerTask

from fakes import FakeClient


def test_get_task_key(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "worker:3")
    assert _task_commons.get_task_key() == ContainerKey("worker", 3)
    monkeypatch.delenv("TFY_TASK_KEY")
    monkeypatch.setenv("SKEIN_CONTAINER_ID", "chief_0")
    assert _task_commons.get_task_key() == ContainerKey("chief", 0)
    monkeypatch.delenv("SKEIN_CONTAINER_ID")
    with pytest.raises(RuntimeError):
        _task_commons.get_task_key()


def test_get_experiment(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "chief:0")
    client = FakeClient({constants.KV_EXPERIMENT_FN: cloudpickle.dumps(lambda: {"answer": 42})})
    assert _task_commons._get_experiment(client) == {"answer": 42}
    assert "chief:0/stop" not in client.kv


def test_get_experiment_failure_posts_start_and_stop(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "chief:0")

    def broken():
        raise ValueError("cannot build the experiment")
    client = FakeClient({constants.KV_EXPERIMENT_FN: cloudpickle.dumps(broken)})
    with pytest.raises(ValueError):
        _task_commons._get_experiment(client)
    assert client.kv["chief:0/start"] == b""
    assert b"cannot build the experiment" in client.kv["chief:0/stop"]


def test_cluster_tasks_and_world_size():
    client = FakeClient({constants.KV_CLUSTER_INSTANCES: json.dumps([["chief", 1, 1], ["worker", 2, 4]])})
    tasks = _task_commons._get_cluster_tasks(client)
    assert tasks == [ContainerTask("chief", 0, 1), ContainerTask("worker", 0, 4), ContainerTask("worker", 1, 4)]
    assert _task_commons._compute_world_size(tasks) == 9


def test_choose_master_rank0_publishes_and_others_wait():
    client = FakeClient()
    host, port = _task_commons.choose_master(client, 0)
    assert client.kv["MASTER_ADDR"].decode() == host and int(client.kv["MASTER_PORT"]) == port
    assert _task_commons.choose_master(client, 3) == (host, port)


def test_rank_table_is_collision_free_for_mixed_roles():
    tasks = [ContainerTask("worker", 0, 2), ContainerTask("chief", 0, 1), ContainerTask("worker", 1, 2),
             ContainerTask("ps", 0, 1)]
    table = _task_commons.rank_table(tasks, roles=("chief", "worker"))
    assert table[("chief", 0, 0)] == 0
    assert sorted(table.values()) == list(range(5))
    assert table[("worker", 1, 1)] == 4
    # the reference formula collides for chief:0 / worker:0 (both rank 0)
    assert _task_commons.compute_rank(0, 0, 1) == _task_commons.compute_rank(0, 0, 1)


def test_roles(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "evaluator:0")
    assert _task_commons.is_evaluator() and not _task_commons.is_chief() and not _task_commons.is_worker()
    assert _task_commons.is_chief("chief") and _task_commons.is_worker("worker")
    monkeypatch.setenv("TF_YARN_N_TRY", "2")
    assert _task_commons.n_try() == 2


def test_sigterm_runs_atexit_hooks(tmp_path):
    """The launcher stops tasks with SIGTERM: task programs turn it into a normal exit so atexit hooks run (the
    parameter-server shards unlink their /dev/shm files there) and the status stays 143."""
    import signal
    import subprocess
    import sys
    import time
    marker = tmp_path / "cleaned"
    ready = tmp_path / "ready"
    code = (
        "import atexit, time, pathlib\n"
        "from tf_yarn_b200 import _task_commons\n"
        "_task_commons.setup_logging()\n"
        f"atexit.register(lambda: pathlib.Path({str(marker)!r}).write_text('x'))\n"
        f"pathlib.Path({str(ready)!r}).write_text('x')\n"
        "time.sleep(60)\n")
    p = subprocess.Popen([sys.executable, "-c", code])
    deadline = time.time() + 60
    while not ready.exists() and time.time() < deadline:
        time.sleep(0.05)
    assert ready.exists()
    p.send_signal(signal.SIGTERM)
    assert p.wait(30) == 128 + signal.SIGTERM
    assert marker.exists()
