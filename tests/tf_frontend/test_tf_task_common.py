from unittest import mock

import pytest

from tf_yarn_b200.tensorflow import Experiment, KerasExperiment
from tf_yarn_b200.tensorflow.tasks import tf_task_common
from tf_yarn_b200.topologies import ContainerTask

from fakes import FakeClient


@pytest.mark.parametrize("task,filters,expected", [
    (ContainerTask("worker", 1, 1), [], True),
    (ContainerTask("worker", 1, 1), ["/job:ps"], False),
    (ContainerTask("ps", 0, 1), ["/job:ps"], True),
    (ContainerTask("worker", 1, 1), ["/job:ps", "/job:worker/task:1"], True),
    (ContainerTask("worker", 2, 1), ["/job:ps", "/job:worker/task:1"], False),
    (ContainerTask("chief", 0, 1), ["/job:master"], True),
])
def test_matches_device_filters(task, filters, expected):
    assert tf_task_common._matches_device_filters(task, filters) == expected


def test_wait_for_connected_tasks_honours_filters():
    client = FakeClient({"ps:0/stop": "", "worker:0/stop": ""})
    tasks = [ContainerTask("ps", 0, 1), ContainerTask("worker", 0, 1), ContainerTask("worker", 1, 1)]
    # worker:1 never stops, but it does not match the filters
    tf_task_common._wait_for_connected_tasks(client, tasks, ["/job:ps", "/job:worker/task:0"])


def test_execute_dispatched_function_runs_thread_and_posts_start(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "worker:0")
    client = FakeClient()
    calls = []
    with mock.patch("tf_yarn_b200.estimator.train_and_evaluate", side_effect=lambda *a: calls.append(a)):
        thread = tf_task_common._execute_dispatched_function(client, Experiment("est", "train", "eval"))
        thread.join()
    assert calls == [("est", "train", "eval")]
    assert client.kv["worker:0/start"] == b""
    assert "worker:0/train_eval_start_time" in client.kv and "worker:0/train_eval_stop_time" in client.kv
    assert thread.state == "SUCCEEDED"


def test_keras_experiment_rejected_by_ps_strategy(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "worker:0")
    with pytest.raises(ValueError, match="KerasExperiment using parameter strategy is unsupported"):
        tf_task_common._execute_dispatched_function(FakeClient(), KerasExperiment(None, "d", {}))
    with pytest.raises(ValueError):
        tf_task_common._execute_dispatched_function(FakeClient(), object())


def test_shutdown_container_posts_stop_waits_and_reraises(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "worker:0")
    client = FakeClient({"chief:0/stop": ""})
    tasks = [ContainerTask("chief", 0, 1), ContainerTask("worker", 0, 1)]
    from tf_yarn_b200._internal import MonitoredThread

    def boom():
        raise RuntimeError("training failed")
    thread = MonitoredThread(target=boom)
    thread.start()
    thread.join()
    with pytest.raises(RuntimeError, match="training failed"):
        tf_task_common._shutdown_container(client, tasks, None, thread)
    assert b"training failed" in client.kv["worker:0/stop"]
    assert "worker:0/container_stop_time" in client.kv


def test_allreduce_task_assigns_ranks_and_exports_the_horovod_environment(monkeypatch):
    """Chief = rank 0, workers follow by id; every trainer turns its rank_info into the HOROVOD_* variables the
    reference's Horovod-gloo task exports (reference: tf_yarn/tensorflow/tasks/gloo_allred_task.py:36-54,94-123)."""
    import os
    from fakes import FakeClient
    from tf_yarn_b200.tensorflow.tasks import allred_task
    from tf_yarn_b200.topologies import ContainerTask
    tasks = [ContainerTask("worker", 2, 1), ContainerTask("chief", 0, 1), ContainerTask("worker", 0, 1),
             ContainerTask("worker", 1, 1)]
    client = FakeClient()
    allred_task._driver_fn(client, tasks)
    kv = client.kv
    assert kv["chief:0/rank_info"] == b"0,4,0,4,0,1" and kv["worker:0/rank_info"] == b"1,4,1,4,0,1"
    assert kv["worker:2/rank_info"] == b"3,4,3,4,0,1"
    host, port = kv["chief:0/sock_addr"].decode().split(":")
    assert int(port) > 0
    monkeypatch.setenv("TFY_TASK_KEY", "worker:1")
    for k in ("HOROVOD_RANK", "HOROVOD_SIZE", "TFY_RANK", "TFY_WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    allred_task._setup_hvd_env(client)
    assert (os.environ["HOROVOD_RANK"], os.environ["HOROVOD_SIZE"], os.environ["HOROVOD_LOCAL_RANK"]) == ("2", "4", "2")
    assert os.environ["HOROVOD_GLOO_RENDEZVOUS_PORT"] == port and os.environ["TFY_WORLD_SIZE"] == "4"
    assert kv["worker:1/addr"]                                   # the reference's per-task address announcement
    for k in [k for k in os.environ if k.startswith("HOROVOD_")] + ["TFY_RANK", "TFY_WORLD_SIZE"]:
        monkeypatch.delenv(k, raising=False)
