"""The dense-layer arithmetic of the oracle (conv1x1 + bias + inference batch norm + ReLU, group max, head activations)
cannot be pinned to the reference itself: TensorFlow 1.10 is not installable here ("unpinned part").
This file cross-checks the C restatement against an INDEPENDENT implementation of the same documented semantics --
torch.nn.functional conv2d (1x1, NHWC weights [1,1,cin,cout] as TF stores them), batch_norm(training=False, eps=1e-3 =
tf.contrib.layers.batch_norm's default, tf_util.py:512-531), relu, max over nsample -- evaluated in f64."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
F = torch.nn.functional

from oracle import pnpp  # noqa: E402


def _layer_weights(rng, cin, cout, scope, bn=True):
    w = {scope + "/weights": rng.normal(0, 1 / np.sqrt(cin), (1, 1, cin, cout)).astype(np.float32),
         scope + "/biases": rng.normal(0, 0.1, cout).astype(np.float32)}
    if bn:
        w[scope + "/bn/gamma"] = rng.uniform(0.5, 1.5, cout).astype(np.float32)
        w[scope + "/bn/beta"] = rng.normal(0, 0.1, cout).astype(np.float32)
        w[scope + "/bn/moving_mean"] = rng.normal(0, 0.1, cout).astype(np.float32)
        w[scope + "/bn/moving_variance"] = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    return w


def _torch_layer(x, w, scope, bn, relu):
    """x (B,m,S,cin) f32 -> (B,m,S,cout) f64 with torch ops (NCHW inside)."""
    t = torch.from_numpy(x.astype(np.float64)).permute(0, 3, 1, 2)
    W = torch.from_numpy(w[scope + "/weights"].astype(np.float64)).reshape(-1, w[scope + "/weights"].shape[-1])
    y = F.conv2d(t, W.t()[:, :, None, None], torch.from_numpy(w[scope + "/biases"].astype(np.float64)))
    if bn:
        y = F.batch_norm(y, torch.from_numpy(w[scope + "/bn/moving_mean"].astype(np.float64)),
                         torch.from_numpy(w[scope + "/bn/moving_variance"].astype(np.float64)),
                         torch.from_numpy(w[scope + "/bn/gamma"].astype(np.float64)),
                         torch.from_numpy(w[scope + "/bn/beta"].astype(np.float64)), training=False, eps=1e-3)
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize("cin,cout,bn,relu", [(3, 64, True, True), (131, 128, True, True), (259, 256, True, True),
                                              (128, 13, False, False)])
def test_conv_bn_relu_matches_torch(cin, cout, bn, relu):
    rng = np.random.default_rng(cin * 1000 + cout)
    w = _layer_weights(rng, cin, cout, "s/conv0", bn)
    x = rng.normal(0, 1, (2, 5, 7, cin)).astype(np.float32)
    got = pnpp.conv1x1(x, w, "s/conv0", bn=bn, relu=relu)
    ref = _torch_layer(x, w, "s/conv0", bn, relu)
    assert got.dtype == np.float32 and got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5)        # f32 accumulation vs f64


def test_shared_mlp_stack_and_group_max_match_torch():
    """Three stacked layers + max over nsample (pointnet_util.py:124-134) against the torch chain in f64."""
    rng = np.random.default_rng(3)
    dims = [6, 32, 32, 64]
    w = {}
    for i in range(3):
        w.update(_layer_weights(rng, dims[i], dims[i + 1], "l/conv%d" % i))
    x = rng.normal(0, 1, (2, 9, 16, dims[0])).astype(np.float32)
    y, yt = x, x
    for i in range(3):
        y = pnpp.conv1x1(y, w, "l/conv%d" % i, bn=True, relu=True)
        yt = _torch_layer(yt.astype(np.float32), w, "l/conv%d" % i, True, True)
    np.testing.assert_allclose(pnpp.group_max(y), yt.max(axis=2), rtol=1e-4, atol=1e-4)


def test_head_activations_match_torch():
    rng = np.random.default_rng(5)
    x = rng.normal(0, 3, (4, 100, 3)).astype(np.float32)
    np.testing.assert_allclose(pnpp._softmax(x), torch.softmax(torch.from_numpy(x).double(), -1).numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(pnpp._sigmoid(x), torch.sigmoid(torch.from_numpy(x).double()).numpy(), rtol=1e-6, atol=1e-7)
