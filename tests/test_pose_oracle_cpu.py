"""Pins oracle/pose_np.py against the REFERENCE pose code:
   * tests/golden/pose_ref.npz -- outputs of the imported reference (tests/golden/make_pose_golden.py), always;
   * the live imported reference (oracle/ref_loader.py) when /root/reference is mounted.
Bit-exact for scores / inlier masks; models to 1e-12 (same numpy/scipy calls in the same order)."""
import os

import numpy as np
import pytest

from articulated_pose_b200 import synthetic
from oracle import pose_np, ref_loader

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pose_ref.npz")


def _part(cloud, pred, j):
    cls = np.argmax(pred["W"], axis=1)
    pidx = np.where(cls == j)[0]
    return pred["nocs_per_point"][pidx, 3 * j:3 * j + 3].astype(np.float64), cloud["P"][pidx].astype(np.float64)


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _cases(gold):
    for ci, c in enumerate(gold["cases"]):
        cat, cid = str(c).split(":")
        cloud = synthetic.make_cloud(int(cid), cat)
        yield ci, cloud, synthetic.teacher_predictions(cloud)


def test_single_ransac_matches_reference_golden(gold):
    th = float(gold["inlier_th"])
    for ci, cloud, pred in _cases(gold):
        for j in range(cloud["n_parts"]):
            src, tgt = _part(cloud, pred, j)
            k = "c%d_p%d_" % (ci, j)
            m, inl, scores = pose_np.ransac_single(src, tgt, th, gold[k + "idx"], return_scores=True)
            np.testing.assert_array_equal(scores, gold[k + "scores"])
            np.testing.assert_array_equal(inl, gold[k + "inl"])
            np.testing.assert_allclose(m["rotation"], gold[k + "R"], atol=1e-12)
            np.testing.assert_allclose(m["scale"], gold[k + "s"], rtol=1e-12)
            np.testing.assert_allclose(m["translation"], gold[k + "t"], atol=1e-12)
            # teacher data: the fit recovers the ground-truth similarity
            assert abs(m["scale"] - cloud["scale_gt"][j]) < 0.02
            assert np.abs(m["rotation"] - cloud["R_gt"][j]).max() < 0.2


def test_joint_ransac_matches_reference_golden(gold):
    th = float(gold["inlier_th"])
    for ci, cloud, pred in _cases(gold):
        src0, tgt0 = _part(cloud, pred, 0)
        for j in range(1, cloud["n_parts"]):
            src1, tgt1 = _part(cloud, pred, j)
            k = "c%d_j%d_" % (ci, j)
            m, inl, scores = pose_np.ransac_joint(src0, tgt0, src1, tgt1, gold[k + "axis"], th, gold[k + "idx0"],
                                                  gold[k + "idx1"], return_scores=True)
            np.testing.assert_array_equal(scores, gold[k + "scores"])
            np.testing.assert_array_equal(inl[0], gold[k + "inl0"])
            np.testing.assert_array_equal(inl[1], gold[k + "inl1"])
            for f in ("rotation0", "scale0", "translation0", "rotation1", "scale1", "translation1"):
                np.testing.assert_allclose(m[f], gold[k + f], atol=1e-10, err_msg=f)


def test_umeyama_matches_reference_golden(gold):
    for ci, cloud, pred in _cases(gold):
        for j in range(cloud["n_parts"]):
            if "c%d_u%d_s" % (ci, j) not in gold:
                pytest.skip("golden has no Umeyama entries")
            m = cloud["cls_gt"] == j
            a = np.hstack([cloud["nocs_gt"][m].astype(np.float64), np.ones((m.sum(), 1))]).T
            c = np.hstack([cloud["P"][m].astype(np.float64), np.ones((m.sum(), 1))]).T
            s, r, t, rt = pose_np.estimate_similarity_umeyama(a, c)
            np.testing.assert_allclose(s, gold["c%d_u%d_s" % (ci, j)], rtol=1e-12)
            np.testing.assert_allclose(r, gold["c%d_u%d_R" % (ci, j)], atol=1e-12)
            np.testing.assert_allclose(t, gold["c%d_u%d_t" % (ci, j)], atol=1e-12)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted")
def test_live_reference_primitives():
    pap, d3, al = ref_loader.load()
    rng = np.random.default_rng(3)
    for n in (3, 3, 3, 17, 200):
        src, tgt = rng.uniform(size=(n, 3)), rng.normal(size=(n, 3))
        np.testing.assert_array_equal(pose_np.rotate_pts(src, tgt), d3.rotate_pts(src, tgt))
        assert pose_np.scale_pts(src, tgt) == d3.scale_pts(src, tgt)
        r0, s0, t0 = d3.transform_pts(src, tgt)
        r1, s1, t1 = pose_np.transform_pts(src, tgt)
        np.testing.assert_array_equal(r0, r1); assert s0 == s1; np.testing.assert_array_equal(t0, t1)
    pts, rv = rng.normal(size=(9, 3)), rng.normal(size=(1, 3))
    np.testing.assert_array_equal(pose_np.rotate_points_with_rotvec(pts, rv), d3.rotate_points_with_rotvec(pts, rv))
    np.testing.assert_array_equal(pose_np.rotate_points_with_rotvec(pts, rv * 0), pts)      # theta = 0 guard
    x = rng.normal(size=6)
    a = pose_np.objective_eval(x, pts[:3], pts[3:6], pts[2:5], pts[4:7], np.ones((3, 1)) * rv, False)
    b = pap.objective_eval(x, pts[:3], pts[3:6], pts[2:5], pts[4:7], np.ones((3, 1)) * rv, False)
    np.testing.assert_array_equal(a, b)


def test_solve_cloud_shapes():
    cloud = synthetic.make_cloud(5)
    pred = synthetic.teacher_predictions(cloud)
    rng = np.random.default_rng(1)
    cls = np.argmax(pred["W"], 1)
    n = [int((cls == j).sum()) for j in range(3)]
    idx_s = [rng.integers(0, n[j], size=(32, 3)) for j in range(3)]
    idx_j0 = [rng.integers(0, n[0], size=(8, 3)) for j in (1, 2)]
    idx_j1 = [rng.integers(0, n[j], size=(8, 3)) for j in (1, 2)]
    out = pose_np.solve_cloud(cloud["P"], pred["nocs_per_point"], pred["W"], pred["joint_axis_per_point"],
                              cloud["joint_cls_gt"], 3, 0.1, idx_s, idx_j0, idx_j1)
    assert len(out["baseline"]) == 3 and len(out["nonlinear"]) == 2
    for j in range(3):
        assert abs(out["baseline"][j]["scale"] - cloud["scale_gt"][j]) < 0.03
    for j in (1, 2):
        assert np.abs(out["nonlinear"][j - 1]["rotation1"] - cloud["R_gt"][j]).max() < 0.2
