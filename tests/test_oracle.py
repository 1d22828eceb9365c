"""Pin the oracle: hand-computed cases, a pure-loop convolution, a second executor, golden vectors."""
import json
from pathlib import Path

import numpy as np
import pytest

from defer_b200 import applications
from oracle import keras_ref as R

GOLD = Path(__file__).parent / "golden"


def test_conv_known_answer():
    x = np.arange(16, dtype=np.float32).reshape(1, 4, 4, 1)
    w = np.ones((3, 3, 1, 1), np.float32)
    y = R.conv2d(x, w, np.array([0.5], np.float32), (1, 1), "valid")
    assert y.reshape(2, 2).tolist() == [[45.5, 54.5], [81.5, 90.5]]
    y = R.conv2d(x, w, None, (2, 2), "same")       # TF SAME on 4 with k=3,s=2: pad (0,1)
    assert y.reshape(2, 2).tolist() == [[45.0, 39.0], [66.0, 50.0]]


def test_same_pad_rule():
    assert R.same_pad(224, 3, 1) == (1, 1)
    assert R.same_pad(224, 3, 2) == (0, 1)
    assert R.same_pad(7, 3, 1) == (1, 1)
    assert R.same_pad(5, 1, 2) == (0, 0)


@pytest.mark.parametrize("k,s,pad", [(1, 1, "valid"), (3, 1, "same"), (3, 2, "same"), (7, 2, "valid"), (1, 2, "valid")])
def test_conv_im2col_matches_loops(k, s, pad):
    rng = np.random.default_rng(k * 10 + s)
    x = rng.standard_normal((2, 9, 8, 3))
    w = rng.standard_normal((k, k, 3, 4))
    b = rng.standard_normal(4)
    assert np.allclose(R.conv2d(x, w, b, (s, s), pad), R.conv2d_loops(x, w, b, (s, s), pad), atol=1e-12)


def test_bn_pool_softmax_known_answers():
    x = np.array([[[[1.0, -2.0]]]], np.float32)
    y = R.batchnorm(x, np.array([2.0, 1.0], np.float32), np.array([0.5, 0.0], np.float32),
                    np.array([1.0, 0.0], np.float32), np.array([3.0, 0.0], np.float32), 1.0)
    assert np.allclose(y, [[[[0.5, -2.0]]]])
    x = np.array([[-1, -2, -3], [-4, -5, -6], [-7, -8, -9]], np.float32).reshape(1, 3, 3, 1)
    # ZeroPadding2D then max-pool: the pad value 0 wins over negative inputs
    assert R.maxpool2d(R.zeropad2d(x, ((1, 1), (1, 1))), (3, 3), (2, 2)).reshape(2, 2).tolist() == [[0, 0], [0, 0]]
    assert np.allclose(R.softmax(np.array([[0.0, np.log(3.0)]])), [[0.25, 0.75]])


def test_two_executors_agree(resnet50, x224):
    from oracle.torch_cpu import TorchCpuModel
    js, ws = resnet50.to_json(), resnet50.get_weights()
    y64 = R.predict(js, ws, x224, dtype=np.float64, final_activation=False)
    y32 = R.predict(js, ws, x224, final_activation=False)
    yt = TorchCpuModel(js, ws).predict(x224, final_activation=False)
    assert R.rel_err(y32, y64) < 1e-5
    assert R.rel_err(yt, y64) < 1e-5


def test_vgg_flatten_order_and_executors():
    from oracle.torch_cpu import TorchCpuModel
    m = applications.VGG16(input_shape=(64, 64, 3))
    x = applications.synthetic_input(1, shape=(64, 64, 3), seed=2)
    js, ws = m.to_json(), m.get_weights()
    y = R.predict(js, ws, x, final_activation=False)
    yt = TorchCpuModel(js, ws).predict(x, final_activation=False)
    assert R.rel_err(yt, y) < 1e-4


def test_golden_vectors():
    """Committed fixtures (tools/make_golden.py) - guard the oracle + synthetic weights against drift."""
    meta = json.loads((GOLD / "meta.json").read_text())
    g = np.load(GOLD / "resnet50_seed1_input0.npz")
    m = applications.ResNet50(seed=meta["weight_seed"])
    x = applications.synthetic_input(1, seed=meta["input_seed"])
    vals = R.WireModel(m.to_json(), m.get_weights()).predict(x, return_all=True)
    assert R.rel_err(vals["fc1000"], g["probs"]) < 1e-5
    logits = R.predict(m.to_json(), m.get_weights(), x, final_activation=False)
    assert R.rel_err(logits, g["logits"]) < 1e-5
    for name in meta["layers"]:
        v = vals[name]
        sub = v[0, ::meta["stride"], ::meta["stride"], :meta["channels"]] if v.ndim == 4 else v[0, :meta["channels"]]
        assert R.rel_err(sub, g[name]) < 1e-5, name
