import threading
import time

import pytest

from tf_yarn_b200 import kv


@pytest.fixture(params=["native", "python"])
def server(request):
    srv = kv.KVServer() if request.param == "native" else kv.PyKVServer()
    yield srv
    srv.stop()


def test_put_get_keys(server):
    c = kv.KVClient(server.address)
    c["worker:0/init"] = b"1.2.3.4:5"
    c.put("worker:1/init", "1.2.3.4:6")
    assert c.get("worker:0/init") == b"1.2.3.4:5"
    assert c["worker:1/init"] == b"1.2.3.4:6"
    assert c.get("missing") is None and c.get("missing", b"dflt") == b"dflt"
    assert "worker:0/init" in c and "nope" not in c
    with pytest.raises(KeyError):
        c["nope"]
    assert c.keys("worker:") == ["worker:0/init", "worker:1/init"]
    c.delete("worker:0/init")
    assert c.keys() == ["worker:1/init"]
    assert c.ping()


def test_wait_blocks_until_put(server):
    c = kv.KVClient(server.address)
    got = []
    t = threading.Thread(target=lambda: got.append(c.wait("late")))
    t.start()
    time.sleep(0.2)
    assert not got
    kv.KVClient(server.address)["late"] = b"value"
    t.join(5)
    assert got == [b"value"]
    assert c.wait("late") == b"value"          # already present: returns at once
    with pytest.raises(TimeoutError):
        c.wait("never", timeout=0.2)


def test_events_replay_and_stream(server):
    c = kv.KVClient(server.address)
    c["a"] = b"1"
    stop = threading.Event()
    seen = []

    def listen():
        for k, v in c.events(stop=stop):
            seen.append((k, v))
    t = threading.Thread(target=listen)
    t.start()
    time.sleep(0.3)
    c["b"] = b"2"
    c["a"] = b"3"
    time.sleep(0.5)
    stop.set()
    t.join(5)
    assert seen == [("a", b"1"), ("b", b"2"), ("a", b"3")]


def test_large_value(server):
    c = kv.KVClient(server.address)
    blob = bytes(range(256)) * 4096          # 1 MiB, like a pickled experiment
    c["experiment_fn"] = blob
    assert c.wait("experiment_fn") == blob


def test_in_memory_kv():
    store = kv.InMemoryKV()
    store["k"] = "v"
    assert store["k"] == b"v" and store.wait("k") == b"v" and store.keys() == ["k"]
    with pytest.raises(TimeoutError):
        store.wait("x", timeout=0.1)
