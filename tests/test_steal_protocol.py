"""Host-side stress test of the tile-stealing ticket board (defer_b200/csrc/steal_board.h).

The header holds the claim / complete / arm protocol once, over a small atomics shim; the CUDA kernel
(conv_steal_kernel, DEFER_STEAL=1) and this model (tests/steal_model.cpp, std::threads standing in for CTAs) compile
the same code.  Invariants checked by the model: every tile exactly once, op o+1 never before op o is complete (and
its data visible), lanes re-armed while other lanes poll them, a lane's workers leave when their lane is complete."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def model_binary(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("steal") / "steal_model"
    subprocess.run([gxx, "-O2", "-std=c++17", "-pthread", str(ROOT / "tests" / "steal_model.cpp"), "-o", str(out)],
                   check=True, capture_output=True, text=True)
    return out


@pytest.mark.parametrize("lanes,workers,runs,seed", [(6, 3, 150, 1), (8, 2, 150, 2), (2, 5, 200, 3), (1, 4, 100, 4)])
def test_ticket_board_protocol(model_binary, lanes, workers, runs, seed):
    r = subprocess.run([str(model_binary), str(lanes), str(workers), str(runs), str(seed)], capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "errors 0" in r.stdout
