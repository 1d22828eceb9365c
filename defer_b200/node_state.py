"""Shared per-node state - the interface of the reference's ``NodeState``
(``/root/reference/src/node_state.py:6-41``) on a condition variable.

Same constructor and the same four attributes (``chunk_size`` read-only; ``next_node``, ``model``,
``weights`` settable), including the reference's "empty string means not set yet" sentinel
(``node_state.py:9-11``).  The reference's consumers poll these fields with ``time.sleep(5)``
(``src/node.py:32-33,95-96``); here every assignment notifies waiters, so ``wait_for`` returns as
soon as the field is published.  ``chunk_size`` only exists for interface parity - the NVLink hop
has no chunking - and the TCP framing helpers of the reference file (``socket_send`` /
``socket_recv``, ``node_state.py:43-101``) have no counterpart on this hot path.
"""
from __future__ import annotations

import threading
from typing import Any, Optional

_UNSET = ""          # the reference's sentinel


class _Published:
    """Data descriptor: a field guarded by the owner's condition variable; writes wake waiters."""

    def __init__(self, read_only: bool = False):
        self.read_only = read_only
        self.slot = ""

    def __set_name__(self, owner, name):
        self.slot = "_" + name

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        with obj._cond:
            return getattr(obj, self.slot)

    def __set__(self, obj, value):
        if self.read_only:
            raise AttributeError(f"{self.slot[1:]} is read-only")
        with obj._cond:
            setattr(obj, self.slot, value)
            obj._cond.notify_all()


class NodeState:
    chunk_size = _Published(read_only=True)
    next_node = _Published()
    model = _Published()
    weights = _Published()

    def __init__(self, chunk_size) -> None:
        self._cond = threading.Condition(threading.Lock())
        self._chunk_size = chunk_size
        self._next_node: Any = _UNSET
        self._model: Any = _UNSET
        self._weights: Any = _UNSET

    def is_set(self, field: str) -> bool:
        v = getattr(self, field)
        return not (isinstance(v, str) and v == _UNSET)

    def wait_for(self, field: str, timeout: Optional[float] = None):
        """Block until ``field`` ('weights' | 'model' | 'next_node') has been published; return its value."""
        slot = "_" + field

        def published() -> bool:
            v = getattr(self, slot)
            return not (isinstance(v, str) and v == _UNSET)

        with self._cond:
            if not self._cond.wait_for(published, timeout=timeout):
                raise TimeoutError(f"NodeState.{field} not set within {timeout}s")
            return getattr(self, slot)
