"""Lifecycle events published by task programs through the KV store.

Key schema ``"<type>:<id>/<stage>"`` and stage names are those of the
reference (reference: tf_yarn/event.py:13-85) so that the client-side report
(`client._handle_events`) classifies tasks identically.
"""
from __future__ import annotations

import logging
import time
import traceback
from typing import Optional

logger = logging.getLogger(__name__)


def wait(client, key: str, timeout: Optional[float] = None) -> str:
    """Block until ``key`` exists in the KV store and return its (decoded) value."""
    logger.info("Waiting for %s", key)
    kv = client.kv
    try:
        raw = kv.wait(key, timeout) if timeout is not None else kv.wait(key)
    except TypeError:  # stores whose wait() takes no timeout
        raw = kv.wait(key)
    return raw.decode() if isinstance(raw, (bytes, bytearray)) else raw


def broadcast(client, key: str, value: str = "") -> None:
    """Publish ``key`` = ``value`` to every listener (tasks and the client's aggregator)."""
    logger.info("Broadcasting %s = %r", key, value if len(value) < 200 else value[:200] + "...")
    client.kv[key] = value.encode()


def maybe_format_exception(exception: Optional[BaseException]) -> str:
    if exception is None:
        return ""
    return "".join(traceback.format_exception(type(exception), exception, exception.__traceback__))


def logs_event(client, task: str, logs: str) -> None:
    broadcast(client, f"{task}/logs", logs)


def url_event(client, task: str, url: str) -> None:
    broadcast(client, f"{task}/url", url)


def init_event(client, task: str, sock_addr: str) -> None:
    broadcast(client, f"{task}/init", sock_addr)


def start_event(client, task: str) -> None:
    broadcast(client, f"{task}/start")


def stop_event(client, task: str, e: Optional[BaseException] = None) -> None:
    """Empty value = clean stop; a formatted traceback marks the task FAILED."""
    broadcast(client, f"{task}/stop", maybe_format_exception(e))


def broadcast_train_eval_start_timer(client, task: str) -> None:
    broadcast(client, f"{task}/train_eval_start_time", str(time.time()))


def broadcast_train_eval_stop_timer(client, task: str) -> None:
    broadcast(client, f"{task}/train_eval_stop_time", str(time.time()))


def broadcast_container_start_time(client, task: str) -> None:
    broadcast(client, f"{task}/container_start_time", str(time.time()))


def broadcast_container_stop_time(client, task: str) -> None:
    broadcast(client, f"{task}/container_stop_time", str(time.time()))
