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


# ------------------------------------------------------------------------------------------------------------------
# Whole network: an independent torch (f64) re-derivation of the TF graph -- pointnet_plusplus/architectures.py:56-95,
# utils/pointnet_util.py:29-91,94-161,206-236, lib/architecture.py:86-161,195-208 -- written from the reference sources,
# not from oracle/pnpp.py.  Only the index-producing ops (FPS, ball query, three_nn: pinned bit-exactly to the reference's own
# kernels elsewhere) are taken from the oracle's trace; gathers, concat orders, layer stacks, pooling, interpolation
# weights, heads and activations are redone here.
# ------------------------------------------------------------------------------------------------------------------
def _t(x):
    return torch.from_numpy(np.asarray(x, np.float64))


def _tconv(x, w, scope, bn=True, relu=True):
    """x (..., cin) f64 tensor -> (..., cout): tf_util.conv2d / conv1d with a 1x1 kernel (tf_util.py:52-185)."""
    W = _t(w[scope + "/weights"]).reshape(-1, w[scope + "/weights"].shape[-1])
    y = x @ W + _t(w[scope + "/biases"])
    if bn:
        y = (y - _t(w[scope + "/bn/moving_mean"])) / torch.sqrt(_t(w[scope + "/bn/moving_variance"]) + 1e-3) * \
            _t(w[scope + "/bn/gamma"]) + _t(w[scope + "/bn/beta"])
    return torch.relu(y) if relu else y


def _group(points, idx):
    """group_point: points (B,n,c), idx (B,m,S) -> (B,m,S,c)"""
    B = points.shape[0]
    return torch.stack([points[b][torch.from_numpy(idx[b].astype(np.int64))] for b in range(B)])


def _torch_forward(P, w, K, tr, mixed_pred, early_split_nocs, prefix="SPFN"):
    e = prefix + "/est_net/"
    xyz0 = _t(P)

    def sa(xyz, points, scope):
        fps = tr[scope + "/fps_idx"].astype(np.int64)
        new_xyz = torch.stack([xyz[b][torch.from_numpy(fps[b])] for b in range(xyz.shape[0])])
        g_xyz = _group(xyz, tr[scope + "/ball_idx"]) - new_xyz[:, :, None, :]          # pointnet_util.py:52-53
        x = g_xyz if points is None else torch.cat([g_xyz, _group(points, tr[scope + "/ball_idx"])], -1)   # :57 [xyz, feat]
        for i in range(3):
            x = _tconv(x, w, "%s/conv%d" % (scope, i))
        return new_xyz, x.max(dim=2).values                                            # :134

    l1_xyz, l1 = sa(xyz0, None, e + "layer1")
    l2_xyz, l2 = sa(l1_xyz, l1, e + "layer2")
    x = torch.cat([l2_xyz, l2], 2)[:, None]                                            # group_all: [xyz, points], :84-88
    for i in range(3):
        x = _tconv(x, w, e + "layer3/conv%d" % i)
    l3 = x.max(dim=2).values                                                           # (B,1,1024)

    def fp(points1, points2, scope, n_mlp):
        idx = tr[scope + "/nn_idx"].astype(np.int64)
        d = torch.clamp(_t(tr[scope + "/nn_dist"]), min=1e-10)                         # :219
        wgt = (1.0 / d) / (1.0 / d).sum(dim=2, keepdim=True)                           # :220-222
        B = points2.shape[0]
        interp = torch.stack([(points2[b][torch.from_numpy(idx[b])] * wgt[b][:, :, None]).sum(dim=1) for b in range(B)])
        x = interp if points1 is None else torch.cat([interp, points1], 2)             # :226 [interp, skip]
        for i in range(n_mlp):
            x = _tconv(x, w, "%s/conv_%d" % (scope, i))
        return x

    l2f = fp(l2, l3, e + "fa_layer1", 2)
    l1f = fp(l1, l2f, e + "fa_layer2", 2)
    l0f = fp(xyz0, l1f, e + "fa_layer3", 3)                                            # points1 = [l0_xyz, 0 channels]
    net = _tconv(l0f, w, e + "fc1")
    out_dims = [K, 3 * K] + ([K, 3 * K] if mixed_pred else []) + [1]
    nn = prefix + "/nocs_net/"
    res = []
    for i in range(len(out_dims)):
        h = net
        if early_split_nocs and i == 1:
            h = _tconv(h, w, nn + "fc11_1", bn=False, relu=False)
        res.append(_tconv(h, w, nn + "fc2_%d" % i, bn=False, relu=False))
    jn = prefix + "/joint_net/"
    X = net
    for j in range(2):
        X = _tconv(X, w, jn + "fc3_%d" % j)
    heads = [_tconv(X, w, jn + "fc4_%d" % k, bn=False, relu=False) for k in range(4)]
    pred = {"W": torch.softmax(res[0], 2), "nocs_per_point": torch.sigmoid(res[1]), "confi_per_point": torch.sigmoid(res[-1]),
            "joint_axis_per_point": torch.tanh(heads[0]), "unitvec_per_point": torch.tanh(heads[1]),
            "heatmap_per_point": torch.sigmoid(heads[2]), "index_per_point": torch.softmax(heads[3], 2)}
    if mixed_pred:
        scale, trans = torch.sigmoid(res[2]), torch.tanh(res[3])
        tiled = scale[..., None].repeat(1, 1, 1, 3).reshape(scale.shape[0], scale.shape[1], 3 * K)     # architecture.py:150
        pred.update(gocs_per_point=pred["nocs_per_point"] * tiled + trans, global_scale=scale, global_translation=trans)
    return {k: v.numpy() for k, v in pred.items()}


@pytest.mark.parametrize("cat,K,mixed,early", [("eyeglasses", 3, True, True), ("eyeglasses", 3, False, False), ("drawer", 4, True, True)])
def test_whole_forward_matches_independent_torch_graph(cat, K, mixed, early):
    from articulated_pose_b200 import synthetic, weights
    P, _ = synthetic.make_batch(range(2), cat, n_points=512)
    w = weights.synthetic_weights(K, mixed, early, seed=11)
    tr = {}
    got = pnpp.forward(P, w, K, nsample=16, mixed_pred=mixed, early_split_nocs=early, npoint1=128, npoint2=32, trace=tr)
    ref = _torch_forward(P, w, K, tr, mixed, early)
    assert set(got) == set(ref)
    for k in ref:
        assert got[k].shape == ref[k].shape, k
        err = np.abs(got[k].astype(np.float64) - ref[k]) / np.maximum(np.abs(ref[k]), 1e-2)
        assert err.max() < 1e-4, (k, err.max())
