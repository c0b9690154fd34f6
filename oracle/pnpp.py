"""oracle/pnpp.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy glue over oracle/pnpp_ref.c) of the ANCSH network forward pass, following

  pointnet_plusplus/utils/pointnet_util.py:29-63   sample_and_group
  pointnet_plusplus/utils/pointnet_util.py:66-91   sample_and_group_all
  pointnet_plusplus/utils/pointnet_util.py:94-161  pointnet_sa_module
  pointnet_plusplus/utils/pointnet_util.py:206-236 pointnet_fp_module
  pointnet_plusplus/architectures.py:56-95         build_pointnet2_shared
  lib/architecture.py:86-161, 195-208              get_per_point_model_new, joint_est_model

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.

TensorFlow 1.10 cannot be installed, so the conv / batch-norm / activation arithmetic is a
restatement of TF semantics ("parity unpinned" for that part); the native ops are
pinned against the reference's own kernels (see oracle/build.py).
"""
import ctypes
import os

import numpy as np

from . import build as _build

_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.ORACLE_SO
        if not os.path.exists(path):
            _build.build_oracle()
        _lib = ctypes.CDLL(path)
        _lib.orc_get_max_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    return a.ctypes.data_as(_F)


def _i(a):
    return a.ctypes.data_as(_I)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def set_threads(n):
    lib().orc_set_threads(ctypes.c_int(int(n)))


def max_threads():
    return int(lib().orc_get_max_threads())


# ------------------------------------------------------------------ native ops (reference wrapper names)
def farthest_point_sample(npoint, inp):
    """tf_sampling.py:48-56 -> (B,npoint) int32"""
    inp = _c(inp, np.float32)
    b, n, _ = inp.shape
    out = np.zeros((b, npoint), np.int32)
    lib().orc_fps(b, n, npoint, _f(inp), _i(out))
    return out


def gather_point(inp, idx):
    """tf_sampling.py:29-37 -> (B,m,3)"""
    inp = _c(inp, np.float32)
    idx = _c(idx, np.int32)
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = np.zeros((b, m, 3), np.float32)
    lib().orc_gather_point(b, n, m, _f(inp), _i(idx), _f(out))
    return out


def query_ball_point(radius, nsample, xyz1, xyz2):
    """tf_grouping.py:8-20 -> idx (B,m,nsample) int32, pts_cnt (B,m) int32"""
    xyz1 = _c(xyz1, np.float32)
    xyz2 = _c(xyz2, np.float32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.zeros((b, m, nsample), np.int32)
    cnt = np.zeros((b, m), np.int32)
    lib().orc_ball_query(b, n, m, ctypes.c_float(radius), nsample, _f(xyz1), _f(xyz2), _i(idx), _i(cnt))
    return idx, cnt


def group_point(points, idx):
    """tf_grouping.py:33-41 -> (B,m,nsample,c)"""
    points = _c(points, np.float32)
    idx = _c(idx, np.int32)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.zeros((b, m, ns, c), np.float32)
    if c > 0:
        lib().orc_group_point(b, n, c, m, ns, _f(points), _i(idx), _f(out))
    return out


def three_nn(xyz1, xyz2):
    """tf_interpolate.py:8-17 -> dist (B,n,3) f32 (squared), idx (B,n,3) int32"""
    xyz1 = _c(xyz1, np.float32)
    xyz2 = _c(xyz2, np.float32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.zeros((b, n, 3), np.float32)
    idx = np.zeros((b, n, 3), np.int32)
    with np.errstate(over="ignore"):
        lib().orc_three_nn(b, n, m, _f(xyz1), _f(xyz2), _f(dist), _i(idx))
    return dist, idx


def three_interpolate(points, idx, weight):
    """tf_interpolate.py:19-28 -> (B,n,c)"""
    points = _c(points, np.float32)
    idx = _c(idx, np.int32)
    weight = _c(weight, np.float32)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.zeros((b, n, c), np.float32)
    lib().orc_three_interpolate(b, m, c, n, _f(points), _i(idx), _f(weight), _f(out))
    return out


def three_weights(dist):
    """pointnet_util.py:219-222"""
    dist = _c(dist, np.float32)
    w = np.zeros_like(dist)
    lib().orc_three_weights(int(dist.size // 3), _f(dist), _f(w))
    return w


# ------------------------------------------------------------------ layers
def conv1x1(x, weights, scope, bn, relu):
    """tf_util.py:120-185 / :52-115.  x (..., cin) -> (..., cout).  `scope` is the TF variable scope."""
    W = _c(weights[scope + "/weights"], np.float32)
    W = W.reshape(-1, W.shape[-1])  # [1,1,cin,cout] / [1,cin,cout] -> (cin,cout)
    bias = _c(weights[scope + "/biases"], np.float32)
    cin, cout = W.shape
    shp = x.shape
    x2 = _c(x, np.float32).reshape(-1, cin)
    y = np.zeros((x2.shape[0], cout), np.float32)
    if bn:
        g = _c(weights[scope + "/bn/gamma"], np.float32)
        be = _c(weights[scope + "/bn/beta"], np.float32)
        mu = _c(weights[scope + "/bn/moving_mean"], np.float32)
        var = _c(weights[scope + "/bn/moving_variance"], np.float32)
        args = (_f(g), _f(be), _f(mu), _f(var))
    else:
        args = (None, None, None, None)
    lib().orc_conv1x1(ctypes.c_long(x2.shape[0]), cin, cout, _f(x2), _f(W), _f(bias), *args, int(relu), _f(y))
    return y.reshape(shp[:-1] + (cout,))


def group_max(x):
    """pointnet_util.py:134: (B,m,S,c) -> (B,m,c)"""
    x = _c(x, np.float32)
    b, m, s, c = x.shape
    y = np.zeros((b, m, c), np.float32)
    lib().orc_group_max(ctypes.c_long(b * m), s, c, _f(x), _f(y))
    return y


def sa_module(xyz, points, npoint, radius, nsample, n_mlp, group_all, weights, scope, trace=None):
    """pointnet_util.py:94-161 (pooling='max', mlp2=None, knn=False, use_xyz=True)."""
    b, n, _ = xyz.shape
    if group_all:
        new_xyz = np.zeros((b, 1, 3), np.float32)                       # :80
        new_points = np.concatenate([xyz, points], axis=2)[:, None]     # :84-88 (un-centred)
        idx = None
    else:
        fps_idx = farthest_point_sample(npoint, xyz)                    # :47
        new_xyz = gather_point(xyz, fps_idx)
        idx, cnt = query_ball_point(radius, nsample, xyz, new_xyz)      # :51
        grouped_xyz = group_point(xyz, idx)                             # :52
        grouped_xyz = grouped_xyz - new_xyz[:, :, None, :]              # :53
        if points is not None and points.shape[-1] > 0:
            new_points = np.concatenate([grouped_xyz, group_point(points, idx)], axis=-1)   # :57
        else:
            new_points = grouped_xyz
        if trace is not None:
            trace[scope + "/fps_idx"] = fps_idx
            trace[scope + "/ball_idx"] = idx
            trace[scope + "/pts_cnt"] = cnt
    for i in range(n_mlp):
        new_points = conv1x1(new_points, weights, "%s/conv%d" % (scope, i), bn=True, relu=True)   # :124-129
    new_points = group_max(new_points)                                  # :134 + squeeze :160
    return new_xyz, new_points, idx


def fp_module(xyz1, xyz2, points1, points2, n_mlp, weights, scope, trace=None):
    """pointnet_util.py:206-236"""
    dist, idx = three_nn(xyz1, xyz2)
    weight = three_weights(dist)
    interpolated = three_interpolate(points2, idx, weight)
    if trace is not None:
        trace[scope + "/nn_idx"] = idx
        trace[scope + "/nn_dist"] = dist
    if points1 is not None:
        x = np.concatenate([interpolated, points1], axis=2)            # :226
    else:
        x = interpolated
    for i in range(n_mlp):
        x = conv1x1(x, weights, "%s/conv_%d" % (scope, i), bn=True, relu=True)   # :230-234
    return x


def _sigmoid(x):
    x = x.astype(np.float32)
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(np.float32)


def _softmax(x):
    x64 = x.astype(np.float64)
    e = np.exp(x64 - x64.max(axis=-1, keepdims=True))
    return (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)


def forward(P, weights, n_max_parts, nsample=64, mixed_pred=True, early_split_nocs=True, prefix="SPFN",
            npoint1=512, npoint2=128, radius1=0.2, radius2=0.4, trace=None):
    """sess.run(pred_dict) restated: lib/architecture.py:86-161 on top of architectures.py:56-95.

    P: (B,N,3) f32.  Returns the pred dict of lib/architecture.py:141-159 (all f32).
    `trace`, when a dict, receives the intermediate index / feature tensors.
    """
    P = _c(P, np.float32)
    K = n_max_parts
    e = prefix + "/est_net/"
    l0_xyz = P[:, :, :3]
    l0_points = P[:, :, 3:3]                                                       # architectures.py:58-59
    l1_xyz, l1_points, _ = sa_module(l0_xyz, l0_points, npoint1, radius1, nsample, 3, False, weights, e + "layer1", trace)
    l2_xyz, l2_points, _ = sa_module(l1_xyz, l1_points, npoint2, radius2, nsample, 3, False, weights, e + "layer2", trace)
    l3_xyz, l3_points, _ = sa_module(l2_xyz, l2_points, None, None, None, 3, True, weights, e + "layer3", trace)
    l2_points = fp_module(l2_xyz, l3_xyz, l2_points, l3_points, 2, weights, e + "fa_layer1", trace)
    l1_points = fp_module(l1_xyz, l2_xyz, l1_points, l2_points, 2, weights, e + "fa_layer2", trace)
    l0_points = fp_module(l0_xyz, l1_xyz, np.concatenate([l0_xyz, l0_points], axis=-1), l1_points, 3, weights,
                          e + "fa_layer3", trace)
    net = conv1x1(l0_points, weights, e + "fc1", bn=True, relu=True)              # architectures.py:89-90
    if trace is not None:
        trace.update(l1_xyz=l1_xyz, l1_points_sa=None, l2_xyz=l2_xyz, l3_points=l3_points, l2_points=l2_points,
                     l1_points=l1_points, l0_points=l0_points, net=net)

    out_dims = [K, 3 * K] + ([K, 3 * K] if mixed_pred else []) + [1]              # architecture.py:98-102
    nn = prefix + "/nocs_net/"
    res = []
    for i, _d in enumerate(out_dims):
        h = net
        if early_split_nocs and i == 1:
            h = conv1x1(h, weights, nn + "fc11_1", bn=False, relu=False)          # :111-112
        res.append(conv1x1(h, weights, nn + "fc2_%d" % i, bn=False, relu=False))  # :113 / :119
    if mixed_pred:
        W, nocs, scale, trans, confi = res
        scale = _sigmoid(scale)                                                   # :124
        trans = np.tanh(trans.astype(np.float64)).astype(np.float32)              # :125
    else:
        W, nocs, confi = res

    jn = prefix + "/joint_net/"
    X = net
    for j in range(2):
        X = conv1x1(X, weights, jn + "fc3_%d" % j, bn=True, relu=True)            # :198-202
    joint_axis = conv1x1(X, weights, jn + "fc4_0", bn=False, relu=False)
    unitvec = conv1x1(X, weights, jn + "fc4_1", bn=False, relu=False)
    heatmap = conv1x1(X, weights, jn + "fc4_2", bn=False, relu=False)
    joint_cls = conv1x1(X, weights, jn + "fc4_3", bn=False, relu=False)

    pred = {
        "W": _softmax(W),                                                         # :131
        "nocs_per_point": _sigmoid(nocs),
        "confi_per_point": _sigmoid(confi),
        "heatmap_per_point": _sigmoid(heatmap),
        "unitvec_per_point": np.tanh(unitvec.astype(np.float64)).astype(np.float32),
        "joint_axis_per_point": np.tanh(joint_axis.astype(np.float64)).astype(np.float32),
        "index_per_point": _softmax(joint_cls),
    }
    if mixed_pred:
        scale_tiled = np.repeat(scale, 3, axis=2)                                 # :150 (k -> k,k,k)
        pred["gocs_per_point"] = (pred["nocs_per_point"] * scale_tiled + trans).astype(np.float32)
        pred["global_scale"] = scale
        pred["global_translation"] = trans
    return pred
