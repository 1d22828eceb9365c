"""Dispatcher with the reference's API (``/root/reference/src/dispatcher.py:20-115``).

``DEFER(computeNodes).run_defer(model, partition_layers, input_stream, output_stream)`` - same names,
same arguments, same blocking behaviour (callers run it in a daemon thread, ``test/test.py:42``).
What changed underneath:

* ``computeNodes[i]`` is a GPU ordinal (or ``"cuda:i"``) of one 8xB200 box instead of an IP;
* ``_dispatchModels`` still ships ``to_json()`` + ``get_weights()`` per stage (``dispatcher.py:49,57``)
  but "shipping" is an upload into that GPU's HBM through ``defer_stage_create`` (same process) or a
  ``torch.distributed`` object send to the rank that owns the GPU (one process per GPU);
* ``_startDistEdgeInference`` / ``_result_server`` (``dispatcher.py:85-105``) keep their roles - feed
  the first stage from ``input_stream``, drain the last into ``output_stream`` in FIFO order - over
  pinned-memory DMA instead of ZFP+LZ4+TCP.

Non-reference additions: ``close()`` / context manager (the reference can only be killed), keyword-only
``dtype``, ``depth`` (in-flight microbatches per stage), ``batch`` (samples per queue item; reference: 1) and
``coalesce``.

Coalescing: the reference's queue items are single images and every node runs them one at a time
(``src/node.py:103-108``), re-reading its weights per image.  Here up to ``coalesce`` in-flight queue items are
gathered into ONE engine microbatch (one kernel chain launch per stage, weights streamed once per group); results
are split back into per-item arrays and delivered in FIFO order, so the API contract (item in, ``(batch, 1000)``
out, same order) is unchanged.  A group is launched as soon as ``coalesce`` items are there or the input queue has
been empty for ``linger_us``; unused sample slots of a partial group are computed and dropped.
"""
from __future__ import annotations

import queue
import threading
import time
from typing import List, Optional

import numpy as np

from . import keras_like as K
from .dag_util import construct_model
from .node import DTYPE_TO_FMT, StageRunner, parse_device


class DEFER:
    def __init__(self, computeNodes, *, dtype: str = "float32", depth: int = 4, batch: Optional[int] = None,
                 coalesce: int = 1, linger_us: float = 200.0, conv_backend: int = 0, dist=None,
                 wait_timeout_ms: int = 0, max_inflight: int = 0) -> None:
        self.computeNodes = list(computeNodes)
        self.dispatchIP = "localhost"       # reference: socket.gethostbyname(...) (dispatcher.py:23); no sockets here
        self.chunk_size = 512 * 1000        # kept for interface parity (dispatcher.py:24)
        self.dtype = dtype
        self.depth = int(depth)
        self.batch = batch                  # samples per queue item
        self.coalesce = max(1, int(coalesce))
        self.linger_s = max(0.0, float(linger_us)) * 1e-6
        self.conv_backend = conv_backend
        self.dist = dist                    # DistContext when launched one-process-per-GPU
        self.wait_timeout_ms = wait_timeout_ms
        # microbatches between ingress and egress.  One process drives all stages => bounded by the lanes of
        # the last stage (depth).  One process per GPU => every stage has its own `depth` lanes and the chain
        # back-pressures itself through the device flags, so depth x stages may be in flight.
        self.max_inflight = int(max_inflight)
        self.stages: List[StageRunner] = []
        self._stop = threading.Event()
        self._ready = threading.Event()
        self._inflight: Optional[threading.Semaphore] = None
        self._submitted = 0                 # engine microbatches (groups of <= coalesce items) stepped so far
        self._group_n: List[int] = []       # items in microbatch seq (ring indexed by seq)
        self._threads: List[threading.Thread] = []
        self._error: Optional[BaseException] = None
        self.results_delivered = 0          # queue items delivered
        self.items_submitted = 0

    @property
    def engine_batch(self) -> int:
        return (self.batch or 1) * self.coalesce

    # ------------------------------------------------------------------ partition (dispatcher.py:27-42)
    def _partition(self, model: K.Model, layer_parts: List[str]) -> List[K.Model]:
        models = []
        for p in range(len(layer_parts) + 1):
            if p == 0:
                start = model.input._keras_history[0].name
            else:
                start = layer_parts[p - 1]
            if p == len(layer_parts):
                end = model.output._keras_history[0].name
            else:
                end = layer_parts[p]
            part = construct_model(model, start, end, part_name=f"part{p+1}")
            models.append(part)
        return models

    # ------------------------------------------------------------------ placement (dispatcher.py:44-65)
    def _dispatchModels(self, models: list, nodeIPs: List) -> None:
        if len(nodeIPs) < len(models):
            raise ValueError(f"{len(models)} stages but only {len(nodeIPs)} compute nodes")
        n = len(models)
        batch = self.engine_batch           # samples per engine microbatch = item batch x coalesced items
        if self.dist is not None:
            # one process per GPU: ship (json, weights, next hop) to each rank; ranks build + link themselves
            for i in range(n):
                next_node = nodeIPs[i + 1] if i != n - 1 else self.dispatchIP
                self.dist.send_stage(i, {"json": models[i].to_json(), "weights": models[i].get_weights(),
                                         "next_node": str(next_node), "fmt": self.dtype, "batch": batch,
                                         "depth": self.depth, "conv_backend": self.conv_backend,
                                         "wait_timeout_ms": self.wait_timeout_ms})
            self.dist.wait_all_ready()      # the 1-byte ACK of dispatcher.py:64-65
            return
        runners = []
        for i in range(n):
            model_json = models[i].to_json()
            weights = models[i].get_weights()
            r = StageRunner.from_wire(model_json, weights, device=parse_device(nodeIPs[i]), dtype=self.dtype,
                                      max_batch=batch, depth=self.depth, is_first=(i == 0), is_last=(i == n - 1),
                                      finalize=False, conv_backend=self.conv_backend,
                                      wait_timeout_ms=self.wait_timeout_ms)
            r.name = f"part{i+1}"
            runners.append(r)
        for i in range(n - 1):              # next hop = nodeIPs[i+1] (dispatcher.py:51-55)
            runners[i].link_to(runners[i + 1])
        for r in runners:
            r.finalize()
        self.stages = runners

    # ------------------------------------------------------------------ ingress (dispatcher.py:85-93)
    def _startDistEdgeInference(self, input: queue.Queue):
        first = self.stages[0] if self.stages else self.dist.local_runner()
        G, B = self.coalesce, self.batch or 1
        hold, nh = self._hold, len(self._hold)
        get_nowait = input.get_nowait
        submit_items = first.submit_items if hasattr(first, "submit_items") else None
        if submit_items is None:             # duck-typed stages (tests): fall back to one call per item
            def submit_items(seq, group):
                for i, x in enumerate(group):
                    first.submit_part(seq, i * B, x)
        try:
            while not self._stop.is_set():
                try:
                    model_input = input.get(timeout=0.05)
                except queue.Empty:
                    continue
                while not self._inflight.acquire(timeout=0.05):
                    if self._stop.is_set():
                        return
                seq = self._submitted
                n = 0
                deadline = None
                group = []
                in_shape = None
                while True:
                    x = model_input
                    if not (isinstance(x, np.ndarray) and x.dtype == np.float32 and x.flags["C_CONTIGUOUS"]):
                        x = np.ascontiguousarray(x, dtype=np.float32)
                    if x.shape[0] != B:
                        raise ValueError(f"queue item has batch {x.shape[0]}, DEFER was built for batch {B}")
                    if in_shape is None:
                        in_shape = x.shape
                    elif x.shape != in_shape:
                        raise ValueError(f"queue items of one group differ in shape: {x.shape} vs {in_shape}")
                    hold[self.items_submitted % nh] = x      # keep alive until the DMA has certainly happened
                    self.items_submitted += 1
                    group.append(x)
                    n += 1
                    if n == G:
                        break
                    try:                                     # coalesce whatever is already waiting ...
                        model_input = get_nowait()
                    except queue.Empty:                      # ... or arrives within the linger window
                        now = time.perf_counter()
                        if deadline is None:
                            deadline = now + self.linger_s
                        if now >= deadline or self._stop.is_set():
                            break
                        try:
                            model_input = input.get(timeout=deadline - now)
                        except queue.Empty:
                            break
                submit_items(seq, group)                     # one C call: a cudaMemcpyAsync per item on the lane's stream
                self._group_n[seq % len(self._group_n)] = n
                if self.dist is not None:
                    first.step(seq)
                    self.dist.mark_submitted(seq + 1)
                else:
                    for r in self.stages:
                        r.step(seq)
                self._submitted = seq + 1
        except BaseException as e:  # surface in close()/run_defer instead of dying silently
            self._error = e
            self._stop.set()

    # ------------------------------------------------------------------ egress (dispatcher.py:95-105)
    def _result_server(self, output: queue.Queue):
        try:
            self._ready.wait()
            seq = 0
            B = self.batch or 1
            local_last = bool(self.stages) or (self.dist is not None and self.dist.world == 1)
            last = (self.stages[-1] if self.stages else self.dist.local_runner()) if local_last else None
            while not self._stop.is_set():
                if seq >= self._submitted:
                    time.sleep(20e-6)
                    continue
                if last is not None:
                    pred = last.result(seq)
                else:
                    pred = self.dist.wait_result(seq, self._stop)
                    if pred is None:
                        return
                    pred = pred.reshape(self.engine_batch, -1)
                n = self._group_n[seq % len(self._group_n)]
                self._inflight.release()
                seq += 1
                for i in range(n):                           # split the group back into queue items, FIFO
                    item = pred[i * B:(i + 1) * B]
                    while not self._stop.is_set():
                        try:
                            output.put(item, timeout=0.05)
                            break
                        except queue.Full:
                            continue
                    self.results_delivered += 1
        except BaseException as e:
            self._error = e
            self._stop.set()

    # ------------------------------------------------------------------ orchestration (dispatcher.py:107-115)
    def run_defer(self, model: K.Model, partition_layers, input_stream: queue.Queue, output_stream: queue.Queue):
        if self.batch is None:
            self.batch = 1
        models_to_dispatch = self._partition(model, partition_layers)
        # the last stage has `depth` result buffers (one process) / the result ring and every stage's lanes bound the
        # chain (one process per GPU): more microbatches in flight than that would overwrite a result before it is read
        cap = self.depth * (self.dist.world if self.dist is not None else 1)
        if self.dist is not None:
            cap = min(cap, self.dist.ring)
        if self.max_inflight <= 0:
            self.max_inflight = cap
        elif self.max_inflight > cap:
            raise ValueError(f"max_inflight {self.max_inflight} exceeds what the pipeline can hold ({cap} = depth "
                             f"{self.depth} x {'ranks' if self.dist is not None else '1'}, result ring included)")
        self._inflight = threading.Semaphore(self.max_inflight)
        self._hold = [None] * ((2 * self.max_inflight + 2) * self.coalesce)
        self._group_n = [0] * (2 * self.max_inflight + 2)
        a = threading.Thread(target=self._result_server, args=(output_stream,), name="defer-result")
        a.start()
        try:
            self._dispatchModels(models_to_dispatch, self.computeNodes)
        except BaseException as e:
            self._error = e
            self._stop.set()
            self._ready.set()
            a.join()
            raise
        self._ready.set()
        b = threading.Thread(target=self._startDistEdgeInference, args=(input_stream,), daemon=True, name="defer-feed")
        b.start()
        self._threads = [a, b]
        a.join()                            # blocks until close(), like the reference blocks forever
        b.join()
        if self._error is not None:
            raise self._error

    # ------------------------------------------------------------------ non-reference: clean shutdown
    def wait_ready(self, timeout: Optional[float] = None) -> bool:
        return self._ready.wait(timeout)

    def close(self):
        self._stop.set()
        for t in self._threads:
            if t is not threading.current_thread():
                t.join(timeout=10)
        if self.dist is not None:
            self.dist.request_stop()
        # Stages write into each other's arenas (outputs + ready flags downstream, free flags upstream): first let
        # EVERY stage finish what is enqueued, then drop every cross-stage mapping, and only then free memory.
        for r in self.stages:
            try:
                r.sync()
            except Exception:
                pass
        for r in self.stages:
            try:
                r.unlink()
            except Exception:
                pass
        for r in self.stages:
            r.close()
        self.stages = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
