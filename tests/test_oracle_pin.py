"""Pins the oracle (oracle/keras_ref.py) and the model zoo (defer_b200/applications.py) against an implementation
we did NOT write: torchvision's VGG16 and ResNet50 classes (torchvision 0.26, CPU).

The reference imports `tensorflow.python.keras.applications` ResNet50 / VGG16 (test/test.py:3,14;
test/local_infer.py:3,8); TensorFlow is not installable here, so parity with the reference itself
stays UNPINNED.  What this test removes is the failure mode where the builder's graph JSON and the builder's oracle
share a topology mistake (stride placement, padding order, flatten order) that cancels out:

* VGG16: torchvision's `vgg16` is arithmetically the Keras VGG16 (138 357 544 parameters).  Seeded Keras-layout
  weights are transplanted (HWIO -> OIHW; fc1 rows re-ordered from Keras' H,W,C flatten to torch's C,H,W).
* ResNet50: torchvision's `resnet50` is the v1.5 variant (stride on the 3x3 conv, bias-free convs, BN eps 1e-5).  The
  Keras `resnet50.py` graph differs exactly by: stride on the first 1x1 conv, conv biases, eps 1e-3.  We move the
  strides on torchvision's own Bottleneck modules, fold each conv bias into the following BN's running mean
  (BN(conv + b) == BN'(conv), mean' = mean - b) and set eps - the block wiring, padding and pooling stay torchvision's.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
tvm = pytest.importorskip("torchvision.models")

from defer_b200 import applications  # noqa: E402
from oracle import keras_ref  # noqa: E402


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).double()


def _conv_w(k_hwio):
    return _t(np.transpose(k_hwio, (3, 2, 0, 1)))      # HWIO -> OIHW


def _softmax(z):
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


@pytest.mark.timeout(600)
def test_vgg16_oracle_matches_torchvision():
    m = applications.VGG16()
    assert m.count_params() == 138_357_544
    tv = tvm.vgg16(weights=None).double().eval()
    assert sum(p.numel() for p in tv.parameters()) == 138_357_544
    convs = [l for l in tv.features if isinstance(l, torch.nn.Conv2d)]
    names = [f"block{b}_conv{c}" for b, n in enumerate([2, 2, 3, 3, 3], start=1) for c in range(1, n + 1)]
    with torch.no_grad():
        for conv, name in zip(convs, names):
            k, b = m.get_layer(name).get_weights()
            conv.weight.copy_(_conv_w(k))
            conv.bias.copy_(_t(b))
        fcs = [l for l in tv.classifier if isinstance(l, torch.nn.Linear)]
        for i, (fc, name) in enumerate(zip(fcs, ["fc1", "fc2", "predictions"])):
            k, b = m.get_layer(name).get_weights()        # (in, out)
            if i == 0:                                    # Keras flattens (H, W, C); torchvision flattens (C, H, W)
                k = k.reshape(7, 7, 512, -1).transpose(2, 0, 1, 3).reshape(25088, -1)
            fc.weight.copy_(_t(k.T))
            fc.bias.copy_(_t(b))
    x = applications.synthetic_input(1, seed=21)
    with torch.no_grad():
        logits = tv(_t(np.transpose(x, (0, 3, 1, 2)))).numpy()
    want = _softmax(logits)
    got64 = keras_ref.predict(m.to_json(), m.get_weights(), x, dtype=np.float64)
    got32 = keras_ref.predict(m.to_json(), m.get_weights(), x, dtype=np.float32)
    assert keras_ref.rel_err(got64, want) <= 1e-9
    assert keras_ref.rel_err(got32, want) <= 1e-5
    assert int(np.argmax(got32)) == int(np.argmax(want))


def _move_stride_to_first_conv(tv):
    for layer in (tv.layer1, tv.layer2, tv.layer3, tv.layer4):
        for blk in layer:
            if blk.conv2.stride != (1, 1):
                blk.conv1.stride = blk.conv2.stride
                blk.conv2.stride = (1, 1)


@pytest.mark.timeout(600)
def test_resnet50_oracle_matches_stride_moved_torchvision():
    m = applications.ResNet50()
    assert m.count_params() == 25_636_712
    tv = tvm.resnet50(weights=None).double().eval()
    _move_stride_to_first_conv(tv)

    def load(conv, bn, conv_name, bn_name):
        k, b = m.get_layer(conv_name).get_weights()
        gamma, beta, mean, var = m.get_layer(bn_name).get_weights()
        with torch.no_grad():
            conv.weight.copy_(_conv_w(k))
            bn.weight.copy_(_t(gamma))
            bn.bias.copy_(_t(beta))
            bn.running_mean.copy_(_t(mean.astype(np.float64) - b.astype(np.float64)))   # conv bias folded into BN
            bn.running_var.copy_(_t(var))
        bn.eps = 1e-3                                     # keras_applications resnet50.py default

    load(tv.conv1, tv.bn1, "conv1", "bn_conv1")
    for si, layer in enumerate((tv.layer1, tv.layer2, tv.layer3, tv.layer4), start=2):
        for bi, blk in enumerate(layer):
            tag = f"{si}{'abcdef'[bi]}"
            load(blk.conv1, blk.bn1, f"res{tag}_branch2a", f"bn{tag}_branch2a")
            load(blk.conv2, blk.bn2, f"res{tag}_branch2b", f"bn{tag}_branch2b")
            load(blk.conv3, blk.bn3, f"res{tag}_branch2c", f"bn{tag}_branch2c")
            if blk.downsample is not None:
                load(blk.downsample[0], blk.downsample[1], f"res{tag}_branch1", f"bn{tag}_branch1")
    k, b = m.get_layer("fc1000").get_weights()
    with torch.no_grad():
        tv.fc.weight.copy_(_t(k.T))
        tv.fc.bias.copy_(_t(b))
    x = applications.synthetic_input(1, seed=22)
    feats = {}
    tv.layer2.register_forward_hook(lambda mod, i, o: feats.__setitem__("layer2", o.detach().numpy()))
    with torch.no_grad():
        logits = tv(_t(np.transpose(x, (0, 3, 1, 2)))).numpy()
    want = _softmax(logits)
    all64 = keras_ref.predict(m.to_json(), m.get_weights(), x, dtype=np.float64, return_all=True)
    got64 = all64["fc1000"]
    # an intermediate tensor too (end of conv3_x = the 2-stage cut of SURVEY 8d), NCHW -> NHWC
    mid = np.transpose(feats["layer2"], (0, 2, 3, 1))
    assert keras_ref.rel_err(all64["activation_21"], mid) <= 1e-9
    assert keras_ref.rel_err(got64, want) <= 1e-9
    got32 = keras_ref.predict(m.to_json(), m.get_weights(), x, dtype=np.float32)
    assert keras_ref.rel_err(got32, want) <= 1e-5
    assert int(np.argmax(got32)) == int(np.argmax(want))


@pytest.mark.timeout(900)
def test_resnet152_oracle_matches_stride_moved_torchvision():
    """Same pin for the `resnet_common` family (BASELINE config 5): torchvision's resnet152 with the stride moved to the
    first 1x1 conv of each down-sampling block, conv biases folded into BN and eps = 1.001e-5 (keras_applications
    `resnet_common.block1`), at a reduced 96x96 resolution to keep the CPU suite short."""
    m = applications.ResNet152(input_shape=(96, 96, 3))
    tv = tvm.resnet152(weights=None).double().eval()
    assert m.count_params() == 60_419_944
    _move_stride_to_first_conv(tv)

    def load(conv, bn, name):
        k, b = m.get_layer(name + "_conv").get_weights()
        gamma, beta, mean, var = m.get_layer(name + "_bn").get_weights()
        with torch.no_grad():
            conv.weight.copy_(_conv_w(k))
            bn.weight.copy_(_t(gamma))
            bn.bias.copy_(_t(beta))
            bn.running_mean.copy_(_t(mean.astype(np.float64) - b.astype(np.float64)))
            bn.running_var.copy_(_t(var))
        bn.eps = 1.001e-5

    load(tv.conv1, tv.bn1, "conv1")
    for si, layer in enumerate((tv.layer1, tv.layer2, tv.layer3, tv.layer4), start=2):
        for bi, blk in enumerate(layer, start=1):
            tag = f"conv{si}_block{bi}"
            load(blk.conv1, blk.bn1, tag + "_1")
            load(blk.conv2, blk.bn2, tag + "_2")
            load(blk.conv3, blk.bn3, tag + "_3")
            if blk.downsample is not None:
                load(blk.downsample[0], blk.downsample[1], tag + "_0")
    k, b = m.get_layer("probs").get_weights()
    with torch.no_grad():
        tv.fc.weight.copy_(_t(k.T))
        tv.fc.bias.copy_(_t(b))
    x = applications.synthetic_input(1, shape=(96, 96, 3), seed=23)
    with torch.no_grad():
        want = _softmax(tv(_t(np.transpose(x, (0, 3, 1, 2)))).numpy())
    got64 = keras_ref.predict(m.to_json(), m.get_weights(), x, dtype=np.float64)
    assert keras_ref.rel_err(got64, want) <= 1e-9
    got32 = keras_ref.predict(m.to_json(), m.get_weights(), x, dtype=np.float32)
    assert keras_ref.rel_err(got32, want) <= 2e-5
