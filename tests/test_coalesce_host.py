"""Host logic of DEFER's coalesced ingress / egress (no GPU): queue items are gathered into engine microbatches of
`coalesce` items, partial groups go out after the linger window, results come back per item in FIFO order
(reference contract: one item in, one `(batch, 1000)` array out, same order - test/test.py:32-34,47-49)."""
import queue
import threading
import time

import numpy as np
import pytest

from defer_b200.dispatcher import DEFER


class FakeStage:
    """Stands in for StageRunner: y[i] = mean(x[i]) broadcast over 5 outputs, computed at step()."""

    def __init__(self, batch, depth):
        self.batch, self.depth = batch, depth
        self.out_shape = (batch, 5)
        self.slots = [np.full((batch, 2, 2, 1), np.nan, np.float32) for _ in range(depth)]
        self.outs = {}
        self.groups = []          # samples written per microbatch
        self._n = {}

    def submit_part(self, seq, index, x):
        self.slots[seq % self.depth][index:index + x.shape[0]] = x
        self._n[seq] = self._n.get(seq, 0) + x.shape[0]

    def step(self, seq):
        x = self.slots[seq % self.depth]
        self.outs[seq] = np.repeat(x.reshape(self.batch, -1).mean(axis=1, keepdims=True), 5, axis=1).astype(np.float32)
        self.groups.append(self._n.pop(seq, 0))

    def result(self, seq, out=None):
        return self.outs.pop(seq)

    def sync(self):
        pass

    def unlink(self):
        pass

    def close(self):
        pass


class FakeDefer(DEFER):
    def _partition(self, model, layer_parts):
        return [None]

    def _dispatchModels(self, models, nodeIPs):
        self.stages = [FakeStage(self.engine_batch, self.depth)]


def _run(defer, n_items, pace_s=0.0, item_batch=1):
    in_q, out_q = queue.Queue(), queue.Queue()
    t = threading.Thread(target=defer.run_defer, args=(None, [], in_q, out_q), daemon=True)
    t.start()
    assert defer.wait_ready(10)
    got = []
    try:
        for i in range(n_items):
            in_q.put(np.full((item_batch, 2, 2, 1), float(i), np.float32))
            if pace_s:
                time.sleep(pace_s)
        for _ in range(n_items):
            got.append(out_q.get(timeout=10))
    finally:
        stage = defer.stages[0]
        defer.close()
        t.join(timeout=10)
    assert not t.is_alive()
    return got, stage


@pytest.mark.parametrize("coalesce", [1, 4, 8])
def test_fifo_and_per_item_results(coalesce):
    got, stage = _run(FakeDefer([0], depth=3, coalesce=coalesce, linger_us=2000), 37)
    assert [g.shape for g in got] == [(1, 5)] * 37
    assert [float(g[0, 0]) for g in got] == [float(i) for i in range(37)]      # FIFO, one result per item
    assert sum(stage.groups) == 37 and max(stage.groups) <= coalesce


def test_flooded_queue_fills_the_groups():
    d = FakeDefer([0], depth=2, coalesce=8, linger_us=50000)
    got, stage = _run(d, 64)
    assert len(got) == 64
    assert stage.groups.count(8) >= 7          # back-to-back items: (almost) every launch carries a full group


def test_partial_group_leaves_after_the_linger_window():
    d = FakeDefer([0], depth=2, coalesce=8, linger_us=500)
    t0 = time.perf_counter()
    got, stage = _run(d, 3, pace_s=0.02)       # items 20 ms apart: far beyond the 0.5 ms linger
    assert [float(g[0, 0]) for g in got] == [0.0, 1.0, 2.0]
    assert stage.groups == [1, 1, 1]           # nobody waited for a full group
    assert time.perf_counter() - t0 < 5.0


def test_item_batch_larger_than_one():
    d = FakeDefer([0], depth=2, batch=2, coalesce=4, linger_us=2000)
    got, stage = _run(d, 10, item_batch=2)
    assert [g.shape for g in got] == [(2, 5)] * 10
    assert [float(g[1, 0]) for g in got] == [float(i) for i in range(10)]
    assert stage.batch == 8


def test_max_inflight_is_validated():
    d = FakeDefer([0], depth=2, coalesce=2, max_inflight=5)
    with pytest.raises(ValueError):
        d.run_defer(None, [], queue.Queue(), queue.Queue())
    d.close()


def test_wrong_item_batch_surfaces_as_error():
    d = FakeDefer([0], depth=2, batch=1, coalesce=2)
    in_q, out_q = queue.Queue(), queue.Queue()
    t = threading.Thread(target=lambda: pytest.raises(ValueError, d.run_defer, None, [], in_q, out_q), daemon=True)
    t.start()
    assert d.wait_ready(10)
    in_q.put(np.zeros((3, 2, 2, 1), np.float32))
    t.join(timeout=10)
    assert not t.is_alive()
    d.close()
