"""World-size-2 `gloo` test of the one-process-per-GPU control plane (no GPU): stage shipment, link-token
exchange order, shared-memory counters / result ring, and the node data loop's host logic with a fake stage."""
import os
import socket
import sys
import time
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeRunner:
    """Stands in for StageRunner: records the wiring calls, 'computes' y = x + rank on the host."""

    def __init__(self, rank, world, depth=2):
        self.rank, self.world, self.depth = rank, world, depth
        self.finalized = False
        self.imported = {}
        self.steps = []
        self.out_shape = (1, 4)

    def export_link(self, role):
        return f"tok-r{self.rank}-role{role}".encode()

    def import_link(self, role, token):
        self.imported[role] = bytes(token)

    def finalize(self):
        self.finalized = True

    def step(self, seq):
        self.steps.append(seq)

    def result(self, seq, out):
        out[...] = float(seq) + 0.5
        return out

    def sync(self):
        pass


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import threading
    from defer_b200.dist import DistContext
    from defer_b200.node import Node
    ctx = DistContext(backend="gloo", ring=8, out_elems=4)
    try:
        runner = FakeRunner(rank, world)
        # --- shipment: rank 0 "dispatcher" scatters one message per stage
        if rank == 0:
            for i in range(world):
                ctx.send_stage(i, {"json": f"stage{i}", "weights": [np.full(3, i, np.float32)], "next_node": str(i + 1)})
            t = threading.Thread(target=ctx.wait_all_ready, daemon=True)
            t.start()
        msg = ctx.recv_stage()
        assert msg["json"] == f"stage{rank}" and msg["weights"][0][0] == rank
        ctx.exchange_links(runner)
        runner.finalize()
        ctx.ack_ready()
        if rank == 0:
            t.join(timeout=30)
            assert not t.is_alive()
        assert ctx.local_runner() is runner
        # neighbours' tokens landed on the right side
        if rank < world - 1:
            assert runner.imported[0] == f"tok-r{rank+1}-role0".encode()
        if rank > 0:
            assert runner.imported[1] == f"tok-r{rank-1}-role1".encode()
        # --- data loop: rank 0 marks microbatches as submitted, the last rank publishes results in order
        node = Node(dist_ctx=ctx, device=rank, poll_s=1e-4)
        th = threading.Thread(target=node._data_loop, args=(runner,), daemon=True)
        th.start()
        n = 20
        if rank == 0:
            got = []
            for s in range(n):
                while s - len(got) >= runner.depth:      # the dispatcher's in-flight throttle
                    got.append(float(ctx.wait_result(len(got), timeout=30)[0]))
                ctx.mark_submitted(s + 1)
            while len(got) < n:
                got.append(float(ctx.wait_result(len(got), timeout=30)[0]))
            assert got == [s + 0.5 for s in range(n)]
            ctx.request_stop()
        th.join(timeout=30)
        assert not th.is_alive()
        if rank == world - 1:
            assert runner.steps == list(range(n))       # every microbatch stepped exactly once, in order
        assert ctx.max_over_ranks(float(rank)) == world - 1
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()))
    finally:
        ctx.close()


@pytest.mark.timeout(180)
def test_control_plane_world2():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    procs = [mpctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, s = q.get(timeout=150)
        res[r] = s
    for p in procs:
        p.join(timeout=30)
    assert res == {0: "ok", 1: "ok"}, res
