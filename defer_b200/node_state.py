"""Shared per-node state (reference ``/root/reference/src/node_state.py:6-41``).

Same fields and the same "empty string means not set yet" sentinel (``node_state.py:9-11``); the
reference polls these with ``time.sleep(5)`` (``src/node.py:32-33,95-96``) - here setters also
notify a condition variable so waiters wake immediately.  ``chunk_size`` is kept for interface
parity; the NVLink hop has no chunking.  The TCP framing helpers (``socket_send``/``socket_recv``,
``node_state.py:43-101``) have no counterpart on the hot path.
"""
from __future__ import annotations

import threading


class NodeState:
    def __init__(self, chunk_size) -> None:
        self._chunk_size = chunk_size
        self._next_node = ""
        self._model = ""
        self._weights = ""
        self._lock = threading.Lock()
        self._cond = threading.Condition(self._lock)

    @property
    def chunk_size(self):
        with self._lock:
            return self._chunk_size

    @property
    def next_node(self):
        with self._lock:
            return self._next_node

    @next_node.setter
    def next_node(self, nx):
        with self._cond:
            self._next_node = nx
            self._cond.notify_all()

    @property
    def model(self):
        with self._lock:
            return self._model

    @model.setter
    def model(self, m):
        with self._cond:
            self._model = m
            self._cond.notify_all()

    @property
    def weights(self):
        with self._lock:
            return self._weights

    @weights.setter
    def weights(self, w):
        with self._cond:
            self._weights = w
            self._cond.notify_all()

    def wait_for(self, field: str, timeout: float = None):
        """Block until ``field`` ('weights' | 'model' | 'next_node') is set; returns its value."""
        attr = "_" + field
        with self._cond:
            ok = self._cond.wait_for(lambda: not _is_unset(getattr(self, attr)), timeout=timeout)
            if not ok:
                raise TimeoutError(f"NodeState.{field} not set within {timeout}s")
            return getattr(self, attr)


def _is_unset(v) -> bool:
    return isinstance(v, str) and v == ""
