"""Pins the CPU oracle (oracle/pnpp_ref.c) against
  * the reference's own three_nn / three_interpolate loops (oracle/_ref/libref_interp.so, compiled from
    /root/reference/.../tf_interpolate.cpp by oracle/build.py -- the .so travels, the sources do not),
  * a literal numpy simulation of the reference FPS kernel's thread/tree structure,
  * hand-checked known answers, incl. the 4-point toy of 3d_interpolation/visu_interpolation.py:12-14.
The GPU-side pinning against the reference's compiled CUDA kernels lives in test_ops_gpu.py.
"""
import ctypes
import os

import numpy as np
import pytest

from oracle import build as obuild
from oracle import pnpp

F = ctypes.POINTER(ctypes.c_float)
I = ctypes.POINTER(ctypes.c_int)


def _ref_interp():
    if not os.path.exists(obuild.REF_INTERP_SO):
        pytest.skip("oracle/_ref/libref_interp.so not built (reference checkout absent)")
    return ctypes.CDLL(obuild.REF_INTERP_SO)


def _clouds(seed, b, n, dup=False):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-0.5, 0.5, size=(b, n, 3)).astype(np.float32)
    if dup:  # tiled clouds (lib/dataset.py:290-317) create exact distance ties
        x[:, n // 2:] = x[:, :n - n // 2]
    return x


@pytest.mark.parametrize("n,m", [(1024, 512), (512, 128), (128, 1), (100, 7), (3, 2)])
def test_three_nn_matches_reference_loop(n, m):
    ref = _ref_interp()
    b = 3
    xyz1 = _clouds(1, b, n)
    xyz2 = _clouds(2, b, m, dup=m > 4)
    d0 = np.zeros((b, n, 3), np.float32); i0 = np.zeros((b, n, 3), np.int32)
    with np.errstate(over="ignore"):
        ref.ref_three_nn(b, n, m, xyz1.ctypes.data_as(F), xyz2.ctypes.data_as(F), d0.ctypes.data_as(F), i0.ctypes.data_as(I))
    d1, i1 = pnpp.three_nn(xyz1, xyz2)
    np.testing.assert_array_equal(i0, i1)
    np.testing.assert_array_equal(d0, d1)          # bit-exact, inf for missing neighbours included


def test_three_interpolate_matches_reference_loop():
    ref = _ref_interp()
    rng = np.random.default_rng(5)
    b, m, c, n = 2, 128, 64, 512
    pts = rng.normal(size=(b, m, c)).astype(np.float32)
    idx = rng.integers(0, m, size=(b, n, 3)).astype(np.int32)
    w = rng.uniform(size=(b, n, 3)).astype(np.float32)
    o0 = np.zeros((b, n, c), np.float32)
    ref.ref_three_interpolate(b, m, c, n, pts.ctypes.data_as(F), idx.ctypes.data_as(I), w.ctypes.data_as(F), o0.ctypes.data_as(F))
    np.testing.assert_array_equal(o0, pnpp.three_interpolate(pts, idx, w))


def test_visu_interpolation_toy():
    # visu_interpolation.py:12-14: 4 known points on a square; a query at a corner gets that corner's value
    xyz2 = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]]], np.float32)
    feat = np.array([[[1.0], [2.0], [3.0], [4.0]]], np.float32)
    xyz1 = np.array([[[0, 0, 0], [1, 1, 0], [0.5, 0.5, 0]]], np.float32)
    d, i = pnpp.three_nn(xyz1, xyz2)
    assert i[0, 0, 0] == 0 and i[0, 1, 0] == 3
    assert list(i[0, 2]) == [0, 1, 2]               # equidistant: lowest indices win, in order
    w = pnpp.three_weights(d)
    out = pnpp.three_interpolate(feat, i, w)
    np.testing.assert_allclose(out[0, :2, 0], [1.0, 4.0], rtol=1e-6)
    np.testing.assert_allclose(out[0, 2, 0], 2.0, rtol=1e-6)


def _fps_literal(xyz, m):
    """numpy transcription of tf_sampling_g.cu:105-170 with explicit per-thread scan + tree (f32)."""
    n = xyz.shape[0]
    T = 512
    temp = np.full(n, 1e38, np.float32)
    out = [0]
    old = 0
    x = xyz.astype(np.float32)
    for _ in range(1, m):
        dxyz = x - x[old]
        # fma(dz,dz,fma(dy,dy,dx*dx)) emulated in float64 then rounded once per fma
        t0 = (dxyz[:, 0].astype(np.float64) * dxyz[:, 0]).astype(np.float32)
        t1 = (dxyz[:, 1].astype(np.float64) * dxyz[:, 1] + t0).astype(np.float32)
        d = (dxyz[:, 2].astype(np.float64) * dxyz[:, 2] + t1).astype(np.float32)
        temp = np.minimum(d, temp)
        best = np.full(T, -1.0, np.float32); besti = np.zeros(T, np.int64)
        for t in range(min(T, n)):
            ks = np.arange(t, n, T)
            v = temp[ks]
            a = int(np.argmax(v))                    # first max in scan order
            best[t], besti[t] = v[a], ks[a]
        u = 0
        while (1 << u) < T:
            for t in range(T >> (u + 1)):
                i1, i2 = (t * 2) << u, (t * 2 + 1) << u
                if best[i1] < best[i2]:
                    best[i1], besti[i1] = best[i2], besti[i2]
            u += 1
        old = int(besti[0])
        out.append(old)
    return np.array(out, np.int32)


@pytest.mark.parametrize("n,m,dup", [(1024, 64, False), (1024, 64, True), (700, 40, True), (40, 40, False)])
def test_fps_matches_literal_kernel_simulation(n, m, dup):
    xyz = _clouds(11, 1, n, dup=dup)
    got = pnpp.farthest_point_sample(m, xyz)[0]
    np.testing.assert_array_equal(got, _fps_literal(xyz[0], m))


def test_fps_tie_rule_prefers_low_lane_not_low_index():
    # n > 512 with all remaining points at identical distance: winner = smallest (k mod 512, k div 512)
    n = 1030
    xyz = np.zeros((1, n, 3), np.float32)
    xyz[0, 0] = [1, 0, 0]                            # start point; all others coincide at the origin
    got = pnpp.farthest_point_sample(3, xyz)[0]
    assert got[0] == 0
    assert got[1] == 512                             # lane 0 scans k=0,512,1024: k=512 is the first max (k=0 has d=0)
    assert got[2] == 0 or got[2] == 512              # everything is at distance 0 afterwards -> lane 0's first


def test_ball_query_semantics():
    xyz1 = np.array([[[0, 0, 0], [0.05, 0, 0], [1, 1, 1], [0.1, 0, 0], [0.15, 0, 0], [0.19, 0, 0]]], np.float32)
    xyz2 = np.array([[[0, 0, 0], [1, 1, 1], [5, 5, 5]]], np.float32)
    idx, cnt = pnpp.query_ball_point(0.2, 4, xyz1, xyz2)
    assert list(idx[0, 0]) == [0, 1, 3, 4] and cnt[0, 0] == 4      # first nsample in index order
    assert list(idx[0, 1]) == [2, 2, 2, 2] and cnt[0, 1] == 1      # first hit pads the row
    assert cnt[0, 2] == 0 and list(idx[0, 2]) == [0, 0, 0, 0]      # empty ball: defined as zeros here


def test_group_point_and_gather():
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(2, 10, 5)).astype(np.float32)
    idx = rng.integers(0, 10, size=(2, 4, 3)).astype(np.int32)
    out = pnpp.group_point(pts, idx)
    for b in range(2):
        np.testing.assert_array_equal(out[b], pts[b][idx[b]])
    g = pnpp.gather_point(pts[:, :, :3].copy(), idx[:, :, 0].copy())
    for b in range(2):
        np.testing.assert_array_equal(g[b], pts[b, idx[b, :, 0], :3])


def test_forward_shapes_and_ranges():
    from articulated_pose_b200 import synthetic, weights
    P, _ = synthetic.make_batch([0])
    for K, mixed, early in ((3, True, True), (3, False, False)):
        w = weights.synthetic_weights(K, mixed, early)
        pred = pnpp.forward(P, w, K, nsample=32, mixed_pred=mixed, early_split_nocs=early)
        assert pred["W"].shape == (1, 1024, K) and pred["nocs_per_point"].shape == (1, 1024, 3 * K)
        np.testing.assert_allclose(pred["W"].sum(-1), 1.0, rtol=1e-5)
        assert ("gocs_per_point" in pred) == mixed
        if mixed:
            np.testing.assert_allclose(pred["gocs_per_point"],
                                       pred["nocs_per_point"] * np.repeat(pred["global_scale"], 3, 2)
                                       + pred["global_translation"], rtol=1e-6, atol=1e-6)
