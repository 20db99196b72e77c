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


def test_wait_connections_do_not_leak_descriptors(server):
    """Every `wait` is its own connection (a parked WAIT must not block puts); a barrier over W ranks is ~W waits per
    rank, so the server has to release the descriptor of each finished one (round-1 review: it only shut them down)."""
    import os
    c = kv.KVClient(server.address)
    c["ready"] = b"1"
    for _ in range(20):
        c.wait("ready")
    time.sleep(0.3)
    before = len(os.listdir("/proc/self/fd"))
    for _ in range(400):
        assert c.wait("ready") == b"1"
    time.sleep(0.5)
    after = len(os.listdir("/proc/self/fd"))
    assert after - before < 20, (before, after)


def test_many_clients_barrier(server):
    """32 threads rendezvous through put + wait of each other's keys, the pattern of the init / stop barriers."""
    n = 32
    errors = []

    def member(i):
        try:
            c = kv.KVClient(server.address)
            c[f"m:{i}/init"] = str(i)
            for j in range(n):
                assert c.wait(f"m:{j}/init", timeout=30) == str(j).encode()
            c.close()
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)
    threads = [threading.Thread(target=member, args=(i,)) for i in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not errors and not any(t.is_alive() for t in threads)
    assert len(kv.KVClient(server.address).keys("m:")) == n


def test_unicode_keys_and_reconnect_after_the_connection_dropped(server):
    c = kv.KVClient(server.address)
    c["tâche:0/état"] = "prêt"
    assert c["tâche:0/état"] == "prêt".encode() and c.keys("tâche") == ["tâche:0/état"]
    c._sock.close()                                    # the next request notices and reconnects once
    assert c.get("tâche:0/état") == "prêt".encode()
    with pytest.raises(RuntimeError, match="no KV address"):
        import os
        old = os.environ.pop(kv.KV_ADDR_ENV, None)
        try:
            kv.KVClient()
        finally:
            if old is not None:
                os.environ[kv.KV_ADDR_ENV] = old
