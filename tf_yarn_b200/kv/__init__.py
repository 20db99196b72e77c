"""Key-value rendezvous store: the control plane of the local launcher.

The reference's whole wire protocol is a handful of verbs on the skein
ApplicationMaster KV store (reference: tf_yarn/event.py:13-18,70-79;
tf_yarn/client.py:566-568,633-657).  Here the launcher process hosts the store
itself:

* :class:`KVServer`   -- the native C++ server (``kv_server.cpp``) running on
  threads inside the launcher process, loaded through ``ctypes``.
* :class:`PyKVServer` -- the same protocol in pure Python, used only when no
  C++ compiler is available.
* :class:`KVClient`   -- what the task programs use; ``kv[key] = bytes``,
  ``kv.get``, ``kv.wait``, ``kv.keys``, ``kv.events`` mirror skein's
  ``app.kv`` object.
* :class:`InMemoryKV` -- in-process store with the same interface (unit tests,
  and single-process runs that need no sockets).
"""
from __future__ import annotations

import ctypes
import os
import select
import socket
import socketserver
import struct
import threading
import time
from typing import Dict, Iterator, List, Optional, Tuple

OP_PUT, OP_GET, OP_WAIT, OP_KEYS, OP_SUBSCRIBE, OP_DEL, OP_PING = 1, 2, 3, 4, 5, 6, 7
ST_OK, ST_NOTFOUND, ST_EVENT, ST_ERROR = 0, 1, 2, 3

KV_ADDR_ENV = "TFY_KV_ADDR"


def _to_bytes(v) -> bytes:
    if isinstance(v, bytes):
        return v
    if isinstance(v, (bytearray, memoryview)):
        return bytes(v)
    if isinstance(v, str):
        return v.encode()
    raise TypeError(f"KV values must be bytes or str, got {type(v)}")


def _pack(op: int, key: str, val: bytes = b"") -> bytes:
    k = key.encode()
    body = struct.pack("<BI", op, len(k)) + k + struct.pack("<I", len(val)) + val
    return struct.pack("<I", len(body)) + body


def _read_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("KV connection closed")
        buf += chunk
    return bytes(buf)


def _read_frame(sock: socket.socket) -> Tuple[int, str, bytes]:
    (length,) = struct.unpack("<I", _read_exact(sock, 4))
    body = _read_exact(sock, length)
    status, klen = struct.unpack_from("<BI", body, 0)
    key = body[5:5 + klen].decode()
    (vlen,) = struct.unpack_from("<I", body, 5 + klen)
    val = body[9 + klen:9 + klen + vlen]
    return status, key, val


# ---------------------------------------------------------------------------
# servers
# ---------------------------------------------------------------------------
class KVServer:
    """Native (C++) KV server hosted on threads of the calling process."""

    def __init__(self, host: str = "127.0.0.1", port: int = 0):
        from tf_yarn_b200.ops import build
        lib_path = build.build_kv(verbose=False)
        self._lib = ctypes.CDLL(lib_path)
        self._lib.tfy_kv_start.restype = ctypes.c_void_p
        self._lib.tfy_kv_start.argtypes = [ctypes.c_char_p, ctypes.c_int]
        self._lib.tfy_kv_port.argtypes = [ctypes.c_void_p]
        self._lib.tfy_kv_stop.argtypes = [ctypes.c_void_p]
        self._h = self._lib.tfy_kv_start(host.encode(), port)
        if not self._h:
            raise OSError(f"cannot start KV server on {host}:{port}")
        self.host = host
        self.port = int(self._lib.tfy_kv_port(self._h))
        self.native = True

    @property
    def address(self) -> str:
        return f"{self.host}:{self.port}"

    def stop(self) -> None:
        if self._h:
            self._lib.tfy_kv_stop(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()


class InMemoryKV:
    """Dict + condition variable with the client interface (also the PyKVServer backend)."""

    def __init__(self):
        self._d: Dict[str, bytes] = {}
        self._log: List[Tuple[str, bytes]] = []
        self._cv = threading.Condition()
        self._stopped = False

    # -- client interface -------------------------------------------------
    def put(self, key: str, value) -> None:
        with self._cv:
            v = _to_bytes(value)
            self._d[key] = v
            self._log.append((key, v))
            self._cv.notify_all()

    __setitem__ = put

    def get(self, key: str, default=None):
        with self._cv:
            return self._d.get(key, default)

    def __getitem__(self, key: str) -> bytes:
        with self._cv:
            return self._d[key]

    def __contains__(self, key: str) -> bool:
        with self._cv:
            return key in self._d

    def wait(self, key: str, timeout: Optional[float] = None) -> bytes:
        deadline = None if timeout is None else time.time() + timeout
        with self._cv:
            while key not in self._d:
                if self._stopped:
                    raise ConnectionError("KV store stopped")
                remaining = None if deadline is None else deadline - time.time()
                if remaining is not None and remaining <= 0:
                    raise TimeoutError(f"timeout waiting for KV key {key!r}")
                self._cv.wait(remaining if remaining is not None else 1.0)
            return self._d[key]

    def keys(self, prefix: str = "") -> List[str]:
        with self._cv:
            return sorted(k for k in self._d if k.startswith(prefix))

    def delete(self, key: str) -> None:
        with self._cv:
            self._d.pop(key, None)

    def events(self, event_type: str = "PUT", stop: Optional[threading.Event] = None) -> Iterator[Tuple[str, bytes]]:
        cursor = 0
        while True:
            with self._cv:
                while cursor >= len(self._log):
                    if self._stopped or (stop is not None and stop.is_set()):
                        return
                    self._cv.wait(0.2)
                batch = self._log[cursor:]
                cursor = len(self._log)
            for item in batch:
                yield item

    def close(self) -> None:
        with self._cv:
            self._stopped = True
            self._cv.notify_all()


class PyKVServer:
    """Pure-Python server speaking the kv_server.cpp protocol (fallback)."""

    def __init__(self, host: str = "127.0.0.1", port: int = 0):
        store = InMemoryKV()
        self.store = store

        class Handler(socketserver.BaseRequestHandler):
            def handle(self):
                sock = self.request
                sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                try:
                    while True:
                        op, key, val = _read_frame(sock)
                        if op == OP_PUT:
                            store.put(key, val)
                            sock.sendall(_pack(ST_OK, "", b""))
                        elif op == OP_GET:
                            v = store.get(key)
                            sock.sendall(_pack(ST_NOTFOUND if v is None else ST_OK, key, v or b""))
                        elif op == OP_WAIT:
                            try:
                                v = store.wait(key)
                                sock.sendall(_pack(ST_OK, key, v))
                            except ConnectionError:
                                sock.sendall(_pack(ST_ERROR, key, b"server stopping"))
                        elif op == OP_KEYS:
                            sock.sendall(_pack(ST_OK, key, "".join(k + "\n" for k in store.keys(key)).encode()))
                        elif op == OP_DEL:
                            store.delete(key)
                            sock.sendall(_pack(ST_OK, "", b""))
                        elif op == OP_PING:
                            sock.sendall(_pack(ST_OK, "", b"pong"))
                        elif op == OP_SUBSCRIBE:
                            for k, v in store.events():
                                sock.sendall(_pack(ST_EVENT, k, v))
                            return
                        else:
                            sock.sendall(_pack(ST_ERROR, "", b"bad op"))
                except (ConnectionError, OSError):
                    return

        class Srv(socketserver.ThreadingTCPServer):
            allow_reuse_address = True
            daemon_threads = True

        self._srv = Srv((host, port), Handler)
        self.host = host
        self.port = self._srv.server_address[1]
        self.native = False
        self._thread = threading.Thread(target=self._srv.serve_forever, daemon=True)
        self._thread.start()

    @property
    def address(self) -> str:
        return f"{self.host}:{self.port}"

    def stop(self) -> None:
        self.store.close()
        self._srv.shutdown()
        self._srv.server_close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.stop()


def start_server(host: str = "127.0.0.1", port: int = 0):
    """Start the native server; fall back to the Python one without a compiler."""
    if os.environ.get("TFY_KV_PYTHON") == "1":
        return PyKVServer(host, port)
    try:
        return KVServer(host, port)
    except (RuntimeError, OSError) as exc:  # no g++ / cannot load
        import logging
        logging.getLogger(__name__).warning("native KV server unavailable (%s); using PyKVServer", exc)
        return PyKVServer(host, port)


# ---------------------------------------------------------------------------
# client
# ---------------------------------------------------------------------------
class KVClient:
    """Client of the launcher's KV store, shaped like skein's ``app.kv``."""

    def __init__(self, address: Optional[str] = None, connect_timeout: float = 30.0):
        address = address or os.environ.get(KV_ADDR_ENV)
        if not address:
            raise RuntimeError(f"no KV address given and ${KV_ADDR_ENV} is not set")
        host, port = address.rsplit(":", 1)
        self.address = address
        self._addr = (host, int(port))
        self._connect_timeout = connect_timeout
        self._lock = threading.Lock()
        self._sock: Optional[socket.socket] = None

    def _connect(self) -> socket.socket:
        deadline = time.time() + self._connect_timeout
        while True:
            try:
                s = socket.create_connection(self._addr, timeout=self._connect_timeout)
                s.settimeout(None)
                s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                return s
            except OSError:
                if time.time() > deadline:
                    raise
                time.sleep(0.05)

    def _request(self, op: int, key: str, val: bytes = b"") -> Tuple[int, str, bytes]:
        with self._lock:
            for attempt in (0, 1):
                if self._sock is None:
                    self._sock = self._connect()
                try:
                    self._sock.sendall(_pack(op, key, val))
                    return _read_frame(self._sock)
                except (ConnectionError, OSError):
                    self._sock = None
                    if attempt:
                        raise
        raise ConnectionError("unreachable")

    def put(self, key: str, value) -> None:
        status, _, v = self._request(OP_PUT, key, _to_bytes(value))
        if status != ST_OK:
            raise RuntimeError(f"KV put failed: {v!r}")

    __setitem__ = put

    def get(self, key: str, default=None):
        status, _, v = self._request(OP_GET, key)
        return v if status == ST_OK else default

    def __getitem__(self, key: str) -> bytes:
        status, _, v = self._request(OP_GET, key)
        if status != ST_OK:
            raise KeyError(key)
        return v

    def __contains__(self, key: str) -> bool:
        return self._request(OP_GET, key)[0] == ST_OK

    def wait(self, key: str, timeout: Optional[float] = None) -> bytes:
        # a dedicated connection: a parked WAIT must not block puts from other threads
        s = self._connect()
        try:
            s.settimeout(timeout)
            s.sendall(_pack(OP_WAIT, key))
            try:
                status, _, v = _read_frame(s)
            except socket.timeout:
                raise TimeoutError(f"timeout waiting for KV key {key!r}")
            if status != ST_OK:
                raise ConnectionError(f"KV wait({key!r}) aborted: {v!r}")
            return v
        finally:
            s.close()

    def keys(self, prefix: str = "") -> List[str]:
        _, _, v = self._request(OP_KEYS, prefix)
        return [k for k in v.decode().split("\n") if k]

    def delete(self, key: str) -> None:
        self._request(OP_DEL, key)

    def ping(self) -> bool:
        try:
            return self._request(OP_PING, "")[2] == b"pong"
        except OSError:
            return False

    def events(self, event_type: str = "PUT", stop: Optional[threading.Event] = None) -> Iterator[Tuple[str, bytes]]:
        """Yield ``(key, value)`` for every PUT, past and future, until the server goes away."""
        s = self._connect()
        try:
            s.sendall(_pack(OP_SUBSCRIBE, ""))
            while stop is None or not stop.is_set():
                # wait for readability with a timeout, then read one whole frame blocking
                readable, _, _ = select.select([s], [], [], 0.25)
                if not readable:
                    continue
                try:
                    status, key, val = _read_frame(s)
                except (ConnectionError, OSError):
                    return
                if status == ST_EVENT:
                    yield key, val
        finally:
            s.close()

    def close(self) -> None:
        with self._lock:
            if self._sock is not None:
                self._sock.close()
                self._sock = None
