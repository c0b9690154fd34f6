"""Parity of the op-level CUDA kernels (through the C ABI / tf_ops mirrors) against
  (1) the CPU oracle (oracle/pnpp_ref.c) and
  (2) the reference's OWN kernels compiled for sm_100a (oracle/_ref/libref_tfops.so) -- which also pins the
      oracle itself against the reference on this GPU.
Bar: bit-exact for every index tensor and for three_nn / three_interpolate / gather / group floats.
"""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _clouds(seed, b, n, dup=False, scale=0.5):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-scale, scale, size=(b, n, 3)).astype(np.float32)
    if dup:
        x[:, n // 2:] = x[:, :n - n // 2]
    return x


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _ref_tfops():
    from oracle import build as obuild
    if not os.path.exists(obuild.REF_TFOPS_SO):
        pytest.skip("oracle/_ref/libref_tfops.so not built")
    lib = ctypes.CDLL(obuild.REF_TFOPS_SO)
    return lib


FPS_CASES = [(4, 1024, 512, False), (4, 512, 128, False), (3, 1024, 512, True), (2, 2048, 512, True), (2, 700, 300, True),
             (2, 40, 40, False), (1, 4096, 64, False), (2, 8192, 16, True), (1, 1, 1, False), (2, 5, 9, False)]


@pytest.mark.parametrize("b,n,m,dup", FPS_CASES)
def test_fps_bit_exact_vs_oracle(b, n, m, dup):
    from articulated_pose_b200 import tf_ops
    from oracle import pnpp
    xyz = _clouds(100 + n + m, b, n, dup)
    got = tf_ops.farthest_point_sample(m, _dev(xyz)).cpu().numpy()
    np.testing.assert_array_equal(got, pnpp.farthest_point_sample(m, xyz))


@pytest.mark.parametrize("b,n,m,dup", [(4, 1024, 512, False), (3, 1024, 512, True), (2, 2048, 256, True), (2, 700, 300, True)])
def test_fps_bit_exact_vs_reference_kernel(b, n, m, dup):
    from articulated_pose_b200 import tf_ops
    from oracle import pnpp
    ref = _ref_tfops()
    xyz = _clouds(200 + n, b, n, dup)
    d = _dev(xyz)
    temp = torch.empty(32 * n, dtype=torch.float32, device="cuda")
    out = torch.empty((b, m), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    assert ref.ref_fps(b, n, m, ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(temp.data_ptr()),
                       ctypes.c_void_p(out.data_ptr()), 1) == 0
    ref_idx = out.cpu().numpy()
    np.testing.assert_array_equal(tf_ops.farthest_point_sample(m, d).cpu().numpy(), ref_idx)
    np.testing.assert_array_equal(pnpp.farthest_point_sample(m, xyz), ref_idx)      # pins the oracle


def test_fps_all_points_identical():
    from articulated_pose_b200 import tf_ops
    from oracle import pnpp
    xyz = np.zeros((2, 1030, 3), np.float32)
    xyz[:, 0] = [1, 0, 0]
    got = tf_ops.farthest_point_sample(5, _dev(xyz)).cpu().numpy()
    np.testing.assert_array_equal(got, pnpp.farthest_point_sample(5, xyz))
    assert got[0, 1] == 512


@pytest.mark.parametrize("n,m1,m2,kind", [(1024, 512, 128, "random"), (1024, 512, 128, "dup"), (2048, 512, 128, "dup"),
                                          (700, 300, 300, "dup"), (1024, 512, 128, "few"), (1024, 512, 128, "one"),
                                          (640, 512, 128, "tiled")])
def test_fps_two_level_equals_two_launches(n, m1, m2, kind):
    """The second sampling level is served as the prefix of the first (ancsh_fps_two_level); it must equal a real second
    launch on the sampled points -- duplicated points (ties), clouds with fewer distinct points than samples (the
    kernel re-picks index 0 once all distances are zero) and loader-style tiled clouds (lib/dataset.py:290-317) included."""
    from articulated_pose_b200 import tf_ops
    b = 3
    xyz = _clouds(300 + n + m1, b, n, kind == "dup")
    if kind == "few":
        xyz[:] = xyz[:, :90].repeat(n // 90 + 1, axis=1)[:, :n]          # 90 distinct points < m2
    if kind == "one":
        xyz[:] = xyz[:, :1]
    if kind == "tiled":
        xyz[:] = np.concatenate([xyz[:, :400], xyz[:, :240]], axis=1)     # 400 distinct points: between m2 and m1
    d = _dev(xyz)
    idx1, xyz1, idx2, xyz2 = tf_ops.farthest_point_sample_two_level(m1, m2, d)
    ref1 = tf_ops.farthest_point_sample(m1, d)
    ref_xyz1 = tf_ops.gather_point(d, ref1)
    ref2 = tf_ops.farthest_point_sample(m2, ref_xyz1)
    ref_xyz2 = tf_ops.gather_point(ref_xyz1, ref2)
    assert torch.equal(idx1, ref1) and torch.equal(xyz1, ref_xyz1)
    assert torch.equal(idx2, ref2), (idx2[0, :8], ref2[0, :8])
    assert torch.equal(xyz2, ref_xyz2)


BQ_CASES = [(4, 1024, 512, 0.2, 64, 0.5), (4, 512, 128, 0.4, 64, 0.5), (2, 1024, 512, 0.2, 32, 0.5),
            (2, 1024, 512, 0.05, 64, 1.0), (2, 2048, 512, 0.2, 64, 0.7), (2, 100, 37, 0.3, 16, 0.5), (1, 33, 5, 10.0, 48, 0.5)]


@pytest.mark.parametrize("b,n,m,r,ns,scale", BQ_CASES)
def test_ball_query_bit_exact(b, n, m, r, ns, scale):
    from articulated_pose_b200 import tf_ops
    from oracle import pnpp
    xyz1 = _clouds(300 + n, b, n, dup=True, scale=scale)
    sel = pnpp.farthest_point_sample(m, xyz1)
    xyz2 = pnpp.gather_point(xyz1, sel)
    idx, cnt = tf_ops.query_ball_point(r, ns, _dev(xyz1), _dev(xyz2))
    idx0, cnt0 = pnpp.query_ball_point(r, ns, xyz1, xyz2)
    np.testing.assert_array_equal(cnt.cpu().numpy(), cnt0)
    np.testing.assert_array_equal(idx.cpu().numpy(), idx0)
    assert cnt0.min() >= 1
    # reference kernel
    ref = _ref_tfops()
    d1, d2 = _dev(xyz1), _dev(xyz2)
    ridx = torch.zeros((b, m, ns), dtype=torch.int32, device="cuda")
    rcnt = torch.zeros((b, m), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    assert ref.ref_ball_query(b, n, m, ctypes.c_float(r), ns, ctypes.c_void_p(d1.data_ptr()), ctypes.c_void_p(d2.data_ptr()),
                              ctypes.c_void_p(ridx.data_ptr()), ctypes.c_void_p(rcnt.data_ptr()), 1) == 0
    np.testing.assert_array_equal(rcnt.cpu().numpy(), cnt0)
    np.testing.assert_array_equal(ridx.cpu().numpy(), idx0)


def test_ball_query_empty_ball_and_external_centroids():
    from articulated_pose_b200 import tf_ops
    from oracle import pnpp
    xyz1 = _clouds(7, 2, 256, scale=0.5)
    xyz2 = _clouds(8, 2, 50, scale=2.0)              # many centroids far from every point
    idx, cnt = tf_ops.query_ball_point(0.2, 8, _dev(xyz1), _dev(xyz2))
    idx0, cnt0 = pnpp.query_ball_point(0.2, 8, xyz1, xyz2)
    assert (cnt0 == 0).any()
    np.testing.assert_array_equal(cnt.cpu().numpy(), cnt0)
    np.testing.assert_array_equal(idx.cpu().numpy(), idx0)


def test_gather_and_group_bit_exact():
    from articulated_pose_b200 import tf_ops
    from oracle import pnpp
    rng = np.random.default_rng(4)
    b, n, m, s, c = 3, 512, 128, 64, 131
    xyz = _clouds(9, b, n)
    pts = rng.normal(size=(b, n, c)).astype(np.float32)
    idx2 = rng.integers(0, n, size=(b, m)).astype(np.int32)
    idx3 = rng.integers(0, n, size=(b, m, s)).astype(np.int32)
    np.testing.assert_array_equal(tf_ops.gather_point(_dev(xyz), _dev(idx2)).cpu().numpy(), pnpp.gather_point(xyz, idx2))
    got = tf_ops.group_point(_dev(pts), _dev(idx3)).cpu().numpy()
    np.testing.assert_array_equal(got, pnpp.group_point(pts, idx3))
    ref = _ref_tfops()
    dp, di = _dev(pts), _dev(idx3)
    out = torch.zeros((b, m, s, c), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    assert ref.ref_group_point(b, n, c, m, s, ctypes.c_void_p(dp.data_ptr()), ctypes.c_void_p(di.data_ptr()),
                               ctypes.c_void_p(out.data_ptr()), 1) == 0
    np.testing.assert_array_equal(out.cpu().numpy(), got)


@pytest.mark.parametrize("b,n,m", [(3, 1024, 512), (3, 512, 128), (2, 128, 1), (2, 100, 2), (1, 2048, 700), (2, 7, 1300)])
def test_three_nn_bit_exact(b, n, m):
    from articulated_pose_b200 import tf_ops
    from oracle import pnpp
    xyz1 = _clouds(20 + n, b, n)
    xyz2 = _clouds(21 + m, b, m, dup=m > 4)
    dist, idx = tf_ops.three_nn(_dev(xyz1), _dev(xyz2))
    d0, i0 = pnpp.three_nn(xyz1, xyz2)
    np.testing.assert_array_equal(idx.cpu().numpy(), i0)
    np.testing.assert_array_equal(dist.cpu().numpy(), d0)


def test_three_interpolate_bit_exact():
    from articulated_pose_b200 import tf_ops
    from oracle import pnpp
    rng = np.random.default_rng(6)
    b, m, c, n = 2, 128, 256, 512
    pts = rng.normal(size=(b, m, c)).astype(np.float32)
    idx = rng.integers(0, m, size=(b, n, 3)).astype(np.int32)
    w = rng.uniform(size=(b, n, 3)).astype(np.float32)
    got = tf_ops.three_interpolate(_dev(pts), _dev(idx), _dev(w)).cpu().numpy()
    np.testing.assert_array_equal(got, pnpp.three_interpolate(pts, idx, w))


def test_shape_errors_raise_like_op_requires():
    from articulated_pose_b200 import tf_ops
    with pytest.raises(ValueError):
        tf_ops.farthest_point_sample(4, torch.zeros((2, 8, 4), device="cuda"))
    with pytest.raises(ValueError):
        tf_ops.query_ball_point(0.2, 4, torch.zeros((2, 8, 3), device="cuda"), torch.zeros((3, 4, 3), device="cuda"))
    with pytest.raises(ValueError):
        tf_ops.three_nn(torch.zeros((2, 8, 3), device="cuda"), torch.zeros((2, 8, 2), device="cuda"))


def test_golden_ops_from_reference_kernels():
    """Committed goldens produced by the REFERENCE kernels on a B200 (tests/golden/make_ops_golden.py)."""
    from articulated_pose_b200 import tf_ops
    path = os.path.join(os.path.dirname(__file__), "golden", "ops_ref_b200.npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    g = np.load(path)
    xyz = g["xyz"]
    np.testing.assert_array_equal(tf_ops.farthest_point_sample(int(g["fps_idx"].shape[1]), _dev(xyz)).cpu().numpy(), g["fps_idx"])
    idx, cnt = tf_ops.query_ball_point(float(g["radius"]), int(g["ball_idx"].shape[2]), _dev(xyz), _dev(g["new_xyz"]))
    np.testing.assert_array_equal(idx.cpu().numpy(), g["ball_idx"])
    np.testing.assert_array_equal(cnt.cpu().numpy(), g["ball_cnt"])
