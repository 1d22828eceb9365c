"""Random-schedule model check of conv_steal_kernel's mbarrier choreography (see tests/steal_pipeline_model.py)."""
import random

import pytest

from steal_pipeline_model import Violation, simulate


@pytest.mark.parametrize("fifo", [2, 4])
@pytest.mark.parametrize("stages", [1, 2, 3, 6])
def test_role_choreography_random_schedules(stages, fifo, monkeypatch):
    import steal_pipeline_model as M
    monkeypatch.setattr(M, "FIFO", fifo)
    rng = random.Random(100 + stages)
    for trial in range(60):
        n = rng.randint(1, 14)
        tiles = [(rng.randint(1, 9), rng.random() < 0.4, rng.random() < 0.15) for _ in range(n)]
        simulate(tiles, stages, seed=trial)


def test_model_catches_a_broken_protocol(monkeypatch):
    """Negative control: with one epilogue arrival missing per accumulator hand-back the model must report it."""
    import steal_pipeline_model as M
    monkeypatch.setattr(M, "EPI_WARPS", 8)
    orig = M.CTA.__init__

    def broken(self, tiles, stages, rng):
        orig(self, tiles, stages, rng)
        self.tempty = [M.MBar(M.EPI_WARPS + 1, f"tempty{b}") for b in range(2)]   # MMA waits for an arrival that never comes
    monkeypatch.setattr(M.CTA, "__init__", broken)
    with pytest.raises(Violation):
        simulate([(2, False, False)] * 5, 2, seed=0)
