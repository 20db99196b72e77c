"""Test doubles: an in-memory application client (the reference mocks skein.ApplicationClient
with a dict as ``.kv``; here the store is the real InMemoryKV so wait/keys/events work)."""
from tf_yarn_b200.kv import InMemoryKV


class FakeClient:
    def __init__(self, initial=None):
        self.kv = InMemoryKV()
        for k, v in (initial or {}).items():
            self.kv[k] = v
        self.shutdown_status = None
        self.id = "application_test"

    def shutdown(self, status):
        self.shutdown_status = status
