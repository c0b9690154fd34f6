"""oracle/iou_np.py (metric kernels) against the reference's own lib/d3_utils.py: the committed golden
vectors (tests/golden/iou_ref.npz, minted by tests/golden/make_iou_golden.py) and, when /root/reference is mounted, the
live functions."""
import os

import numpy as np
import pytest

from oracle import iou_np, ref_loader

GOLD = os.path.join(os.path.dirname(__file__), "golden", "iou_ref.npz")


@pytest.mark.parametrize("nres", [50, 17])
def test_iou_matches_reference_golden(nres):
    g = np.load(GOLD)
    b1, b2, want = g["nres%d_bbox1" % nres], g["nres%d_bbox2" % nres], g["nres%d_iou" % nres]
    got = np.array([float(iou_np.iou_3d(b1[i], b2[i], nres)) for i in range(len(b1))])
    np.testing.assert_array_equal(got, want)                     # integer counts -> identical ratios
    assert (want == 0).any() and (want == 1).any() and ((want > 0) & (want < 1)).sum() > 10


def test_degenerate_union_is_one():
    g = np.load(GOLD)
    z = g["degenerate_bbox"][0]
    assert iou_np.iou_3d(z, z) == 1 == g["degenerate_iou"][0]


def test_get_3d_bbox_matches_reference_golden():
    g = np.load(GOLD)
    out = iou_np.get_3d_bbox(g["bbox_f32_in"], shift=np.array([0.5, 0.5, 0.5]))
    np.testing.assert_array_equal(out, g["bbox_f32_out"])
    assert out.dtype == g["bbox_f32_out"].dtype
    np.testing.assert_array_equal(iou_np.get_3d_bbox(0.8, 0), g["bbox_scalar_out"])


def test_amodal_extent_and_part_boxes():
    from articulated_pose_b200 import synthetic
    cloud = synthetic.make_cloud(11)
    pred = synthetic.teacher_predictions(cloud)
    ext, cnt = iou_np.amodal_extent(pred["nocs_per_point"], pred["W"], 3)
    assert ext.dtype == np.float32 and cnt.sum() == 1024 and (ext > 0).all() and (ext <= 1.0 + 1e-6).all()
    box = iou_np.part_boxes(ext[0], 1.3, np.eye(3), np.zeros(3))
    assert box.shape == (8, 3)
    np.testing.assert_allclose(box.max(0) - box.min(0), ext[0].astype(np.float64) * 1.3, rtol=1e-6)
    # empty part -> NaN extent, zero count
    W = pred["W"].copy(); W[:, 2] = -1.0
    ext2, cnt2 = iou_np.amodal_extent(pred["nocs_per_point"], W, 3)
    assert cnt2[2] == 0 and np.isnan(ext2[2]).all()


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not mounted")
def test_live_reference_iou_and_inside():
    _, d3, _ = ref_loader.load()
    rng = np.random.default_rng(9)
    for _ in range(6):
        e1, e2 = rng.uniform(0.1, 1, 3), rng.uniform(0.1, 1, 3)
        a = d3.get_3d_bbox(e1, 0).T + rng.normal(0, 0.1, 3)
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        b = np.dot(d3.get_3d_bbox(e2, 0).T, q.T) + rng.normal(0, 0.1, 3)
        assert float(d3.iou_3d(a, b, nres=23)) == float(iou_np.iou_3d(a, b, 23))
        pts = rng.uniform(-1, 1, (500, 3))
        np.testing.assert_array_equal(d3.pts_inside_box(pts, b).reshape(-1), iou_np.pts_inside_box(pts, b))


def test_joint_vote_restatement_on_teacher_predictions():
    """With teacher predictions the voted axis is the GT axis up to the injected noise and every joint has voters."""
    from articulated_pose_b200 import synthetic
    cloud = synthetic.make_cloud(3)
    pred = synthetic.teacher_predictions(cloud)
    K = 3
    N = cloud["P"].shape[0]
    index = np.eye(3, dtype=np.float32)[cloud["joint_cls_gt"]]
    gocs = np.tile(cloud["nocs_gt"], (1, K)).astype(np.float32)
    unit = np.zeros((N, 3), np.float32); heat = np.ones((N, 1), np.float32)
    out = iou_np.joint_vote(gocs, pred["W"], unit, heat, pred["joint_axis_per_point"], index, K)
    assert len(out) == K - 1
    for j in range(1, K):
        sel = cloud["joint_cls_gt"] == j
        assert sel.sum() > 0
        np.testing.assert_array_equal(out[j - 1]["l"], np.median(pred["joint_axis_per_point"][sel], axis=0))
        assert out[j - 1]["p"].dtype == np.float32
