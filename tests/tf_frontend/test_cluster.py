import json
import os

import pytest

from tf_yarn_b200.tensorflow import cluster
from tf_yarn_b200.topologies import ContainerTask

from fakes import FakeClient

TASKS = [ContainerTask("worker", 1, 1), ContainerTask("chief", 0, 1), ContainerTask("ps", 0, 1),
         ContainerTask("worker", 0, 1)]


def _client():
    return FakeClient({"chief:0/init": "1.1.1.1:1", "worker:0/init": "1.1.1.2:2", "worker:1/init": "1.1.1.3:3",
                       "ps:0/init": "1.1.1.4:4"})


def test_aggregate_spec_orders_by_task_id():
    assert cluster.aggregate_spec(_client(), TASKS) == {
        "chief": ["1.1.1.1:1"], "ps": ["1.1.1.4:4"], "worker": ["1.1.1.2:2", "1.1.1.3:3"]}


def test_start_cluster_posts_init(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "worker:1")
    client = FakeClient({"chief:0/init": "a:1", "worker:0/init": "b:2", "ps:0/init": "c:3"})
    spec = cluster.start_cluster(("1.2.3.4", 99), client, TASKS)
    assert client.kv["worker:1/init"] == b"1.2.3.4:99"
    assert spec["worker"] == ["b:2", "1.2.3.4:99"]


def test_setup_tf_config(monkeypatch):
    monkeypatch.setenv("TFY_TASK_KEY", "worker:1")
    monkeypatch.delenv("TF_CONFIG", raising=False)
    cluster.setup_tf_config({"chief": ["a:1"], "worker": ["b:2", "c:3"]})
    cfg = json.loads(os.environ["TF_CONFIG"])
    assert cfg["task"] == {"type": "worker", "index": 1} and cfg["environment"] == "google"
    with pytest.raises(RuntimeError):          # exclusive set: never overwrite an existing TF_CONFIG
        cluster.setup_tf_config({"chief": ["a:1"]})
    os.environ.pop("TF_CONFIG", None)      # set behind monkeypatch's back by xset_environ


@pytest.mark.parametrize("task,started", [("worker:0", True), ("chief:0", True), ("ps:0", False),
                                          ("evaluator:0", False)])
def test_start_tf_server_only_for_trainers(monkeypatch, task, started):
    monkeypatch.setenv("TFY_TASK_KEY", task)
    server = cluster.start_tf_server({"chief": ["a:1"], "worker": ["b:2"], "ps": ["c:3"]})
    assert (server is not None) == started
