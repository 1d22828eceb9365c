"""Balanced cut selection: DP optimality, legality of the chosen cuts, oracle parity."""
import itertools

import numpy as np
import pytest

from defer_b200 import applications, autocut, planner
from defer_b200.dispatcher import DEFER
from oracle import keras_ref as R


def test_minimax_partition_is_optimal():
    rng = np.random.default_rng(0)
    for _ in range(30):
        n = int(rng.integers(5, 11))
        costs = rng.uniform(0.1, 5.0, n).tolist()
        allowed = sorted(rng.choice(n - 1, size=int(rng.integers(3, n - 1)), replace=False).tolist())
        k = int(rng.integers(2, min(4, len(allowed) + 1) + 1))
        got = autocut.minimax_partition(costs, k, allowed)
        assert len(got) == k - 1 and all(g in allowed for g in got)

        def worst(cuts):
            b = [-1] + list(cuts) + [n - 1]
            return max(sum(costs[b[i] + 1:b[i + 1] + 1]) for i in range(len(b) - 1))
        brute = min(worst(c) for c in itertools.combinations(allowed, k - 1))
        assert abs(worst(got) - brute) < 1e-12


def test_articulation_points_resnet50(resnet50):
    arts = autocut.articulation_layers(resnet50)
    assert "add_2" in arts and "activation_9" in arts and "max_pooling2d" in arts and "avg_pool" in arts
    assert "res2b_branch2a" not in arts and "bn3a_branch2b" not in arts      # the shortcut bypasses them


def test_balanced_cuts_are_legal_and_better_balanced(x224):
    m = applications.ResNet50()
    cuts, stage_costs = autocut.balanced_cuts(m, 8)
    assert len(cuts) == 7 and len(stage_costs) == 8
    # chosen cuts keep every conv+BN+Add+ReLU fusion: they are post-ReLU activations / pools, never an Add
    assert all(m.get_layer(c).class_name in ("Activation", "MaxPooling2D") for c in cuts)
    # the partition is legal for the reference partitioner and numerically identical on the oracle
    parts = DEFER(list(range(8)))._partition(m, cuts)
    ref = R.predict(m.to_json(), m.get_weights(), x224)
    y = R.pipeline_predict([(p.to_json(), p.get_weights()) for p in parts], x224)
    assert np.array_equal(y, ref)
    # and it is better balanced (under the same cost model) than the reference's hand-made list
    plan = planner.plan_stage(applications.ResNet50(), True, True)
    costs = autocut.analytic_op_costs(plan)
    ends = [i for i, op in enumerate(plan.ops) if any(l in applications.RESNET50_TEST_CUTS for l in op.layers)]
    b = [-1] + ends + [len(costs) - 1]
    ref_worst = max(sum(costs[b[i] + 1:b[i + 1] + 1]) for i in range(8))
    assert max(stage_costs) < ref_worst


def test_balanced_cuts_other_models():
    v = applications.VGG16(weights=None)
    cuts, _ = autocut.balanced_cuts(v, 4)
    assert len(cuts) == 3
    DEFER([0] * 4)._partition(v, cuts)
    m = applications.ResNet152(weights=None)
    cuts, sc = autocut.balanced_cuts(m, 8)
    DEFER([0] * 8)._partition(m, cuts)
    assert max(sc) / (sum(sc) / 8) < 1.35
