"""The real cross-GPU hop (reference: src/node.py:76-79,107-108 -> 89-90): one process per GPU under torchrun,
`Node.run` + CUDA-IPC link tokens + device flags over NVLink, checked numerically; plus the one-process
peer-access variant on distinct devices.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from defer_b200 import _cabi as A
from defer_b200 import applications

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
ROOT = Path(__file__).resolve().parents[1]


def _n_gpus():
    try:
        return A.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("coalesce", [1, 4])
def test_cross_process_hop_parity(coalesce):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, HOP_COALESCE=str(coalesce), HOP_ITEMS="14")
    env.pop("CUDA_VISIBLE_DEVICES", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "dist_hop_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540, env=env, cwd=str(ROOT))
    tail = (r.stdout[-3000:] + "\n--- stderr ---\n" + r.stderr[-3000:])
    assert r.returncode == 0 and "HOP_OK" in r.stdout, tail
    print(r.stdout[-500:])


def test_pipeline_on_distinct_devices_bitwise(resnet50, x224):
    """One process, stage i on GPU i (peer access): same answer as one stage on one GPU, bit for bit."""
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    from test_gpu_model import _oracle, _pipeline_on_one_gpu, _rel
    from defer_b200.node import StageRunner
    k = min(n, 4)
    cuts = applications.default_cuts(resnet50, k)
    outs = _pipeline_on_one_gpu(resnet50, cuts, x224, "float32", depth=3, n_items=7, devices=list(range(k)))
    r = StageRunner.from_model(resnet50, device=0, dtype="float32", max_batch=1, depth=1)
    try:
        whole = r.predict(x224)
    finally:
        r.close()
    ref = _oracle(resnet50, x224)
    for y in outs:
        assert _rel(y, ref) <= 1e-3
        assert np.array_equal(y, whole)
