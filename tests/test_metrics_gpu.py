"""CUDA metric kernels (ancsh_box_iou_3d, ancsh_amodal_extent) through the C ABI against
  * tests/golden/iou_ref.npz -- outputs of the REFERENCE's own d3_utils.iou_3d;
  * oracle/iou_np.py on seeded inputs, including the whole per-cloud IoU loop of compute_miou.py:196-231.
Bars: intersect / union counts and part sizes bit-exact (integer work), extents bit-exact (f32 max), IoU ratios identical."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

GOLD = os.path.join(os.path.dirname(__file__), "golden", "iou_ref.npz")


@pytest.mark.parametrize("nres", [50, 17])
def test_iou_matches_reference_golden(nres):
    from articulated_pose_b200 import metrics
    g = np.load(GOLD)
    b1, b2, want = g["nres%d_bbox1" % nres], g["nres%d_bbox2" % nres], g["nres%d_iou" % nres]
    got = metrics.iou_3d_batch(b1, b2, nres)
    np.testing.assert_array_equal(got, want)
    assert metrics.iou_3d(b1[0], b2[0], nres) == want[0]
    z = g["degenerate_bbox"][0]
    assert metrics.iou_3d(z, z) == 1


def test_iou_counts_match_oracle_and_properties():
    from articulated_pose_b200 import metrics
    from oracle import iou_np
    rng = np.random.default_rng(3)
    n = 40
    b1, b2 = np.zeros((n, 8, 3)), np.zeros((n, 8, 3))
    for i in range(n):
        q1, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        q2, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        b1[i] = np.dot(iou_np.get_3d_bbox(rng.uniform(0.1, 1, 3), 0).T, q1.T) + rng.normal(0, 0.1, 3)
        b2[i] = np.dot(iou_np.get_3d_bbox(rng.uniform(0.1, 1, 3), 0).T, q2.T) + rng.normal(0, 0.1, 3)
    iou, inter, uni = metrics.iou_3d_batch(b1, b2, 50, counts=True)
    for i in range(n):
        assert (int(inter[i]), int(uni[i])) == iou_np.iou_counts(b1[i], b2[i], 50), i
    # symmetry, self-IoU, bounds (size independent)
    iou_r = metrics.iou_3d_batch(b2, b1, 50)
    np.testing.assert_array_equal(iou, iou_r)
    np.testing.assert_array_equal(metrics.iou_3d_batch(b1, b1, 50), np.ones(n))
    assert ((iou >= 0) & (iou <= 1)).all() and (inter <= uni).all()
    # empty batch and bad shapes
    assert metrics.iou_3d_batch(np.zeros((0, 8, 3)), np.zeros((0, 8, 3))).shape == (0,)
    with pytest.raises(ValueError):
        metrics.iou_3d_batch(b1, b2[:3])


def test_amodal_extent_and_part_ious_match_oracle():
    from articulated_pose_b200 import metrics, synthetic
    from articulated_pose_b200.pose import PoseSolver, compute_gt_pose
    from oracle import iou_np
    ids = [11, 12, 13, 14]
    clouds = [synthetic.make_cloud(i) for i in ids]
    preds = [synthetic.teacher_predictions(c) for c in clouds]
    K, B = 3, len(ids)
    nocs = np.stack([p["nocs_per_point"] for p in preds]); W = np.stack([p["W"] for p in preds])
    W[3, :, 2] = -1.0                                                           # cloud 3: part 2 predicted empty
    ext, cnt = metrics.amodal_extent(nocs, W)
    for b in range(B):
        e0, c0 = iou_np.amodal_extent(nocs[b], W[b], K)
        np.testing.assert_array_equal(cnt[b], c0)
        np.testing.assert_array_equal(ext[b], e0)                               # NaN == NaN under assert_array_equal
    assert cnt[3, 2] == 0 and np.isnan(ext[3, 2]).all()
    # the whole compute_miou loop: poses from the CUDA pose stage, GT poses from the CUDA Umeyama
    P = np.stack([c["P"] for c in clouds]); ax = np.stack([p["joint_axis_per_point"] for p in preds])
    jc = np.stack([c["joint_cls_gt"] for c in clouds])
    W_ok = np.stack([p["W"] for p in preds])
    res = PoseSolver(K, niter_single=128, niter_joint=16, seed=2).solve(P, nocs, W_ok, ax, jc)
    gts = [compute_gt_pose(c["P"], c["nocs_gt"], c["cls_gt"], K) for c in clouds]
    rot = np.stack([[r["baseline"][j]["rotation"] for j in range(K)] for r in res])
    tr = np.stack([[r["baseline"][j]["translation"] for j in range(K)] for r in res])
    sc = np.stack([[r["baseline"][j]["scale"] for j in range(K)] for r in res])
    rt_gt = np.stack([np.stack(g["rt"]["gt"]) for g in gts]); s_gt = np.stack([[g["scale"]["gt"][j][0] for j in range(K)] for g in gts])
    ext_gt = np.stack([[2 * np.abs(c["nocs_gt"][c["cls_gt"] == j] - 0.5).max(0) for j in range(K)] for c in clouds])
    iou, ext2 = metrics.part_ious(nocs, W_ok, rot, tr, sc, rt_gt, s_gt, ext_gt)
    # (the teacher NOCS carry 10 % uniform outliers, so the amodal extents are inflated towards the unit cube and the IoU
    #  of the thin eyeglasses parts is small -- the values are compared with the oracle, not with a quality bar)
    assert iou.shape == (B, K) and ((iou > 0) & (iou <= 1)).all(), iou
    for b in range(B):
        for j in range(K):
            g = iou_np.part_boxes(ext_gt[b, j], s_gt[b, j], rt_gt[b, j][:3, :3], rt_gt[b, j][:3, 3])
            p = iou_np.part_boxes(ext2[b, j], sc[b, j], rot[b, j], tr[b, j])
            assert iou[b, j] == iou_np.iou_3d(g, p, 50), (b, j)


def test_part_ious_of_exact_predictions_are_high():
    """Sanity of the whole metric chain: predictions equal to the ground truth (NOCS, labels) and the GT pose give boxes
    that coincide up to f32 rounding -> IoU near 1 for every part."""
    from articulated_pose_b200 import metrics, synthetic
    from articulated_pose_b200.pose import compute_gt_pose
    K = 3
    clouds = [synthetic.make_cloud(i) for i in (15, 16)]
    B = len(clouds)
    nocs = np.stack([np.tile(c["nocs_gt"], (1, K)) for c in clouds]).astype(np.float32)
    W = np.stack([np.eye(K, dtype=np.float32)[c["cls_gt"]] for c in clouds])
    gts = [compute_gt_pose(c["P"], c["nocs_gt"], c["cls_gt"], K) for c in clouds]
    rt_gt = np.stack([np.stack(g["rt"]["gt"]) for g in gts])
    s_gt = np.stack([[g["scale"]["gt"][j][0] for j in range(K)] for g in gts])
    ext_gt = np.stack([[2 * np.abs(c["nocs_gt"][c["cls_gt"] == j] - 0.5).max(0) for j in range(K)] for c in clouds])
    iou, ext = metrics.part_ious(nocs, W, rt_gt[:, :, :3, :3], rt_gt[:, :, :3, 3], s_gt, rt_gt, s_gt, ext_gt)
    np.testing.assert_array_equal(ext, ext_gt.astype(np.float32))
    assert (iou > 0.97).all(), iou


@pytest.mark.parametrize("cat,gn3", [("eyeglasses", False), ("drawer", False), ("eyeglasses", True)])
def test_joint_vote_matches_oracle_bit_exact(cat, gn3):
    """ancsh_joint_vote on the network's own outputs (random-weight network: every branch of the argmaxes is exercised)
    against the NumPy restatement of eval_joint_params.py:178-190 -- medians are order statistics: bit-exact."""
    from articulated_pose_b200 import metrics, synthetic, weights
    from articulated_pose_b200.network import AncshNet
    from oracle import iou_np
    ids = range(70, 74)
    P, clouds = synthetic.make_batch(ids, cat)
    K = clouds[0]["n_parts"]
    pred = AncshNet(weights.synthetic_weights(K), K, nsample=32).forward(P)
    gocs = pred["gocs_per_point"][:, :, :3].copy() if gn3 else pred["gocs_per_point"]
    # a random-weight network votes for one joint class only: scatter the votes (seeded) so that every joint has voters
    rng = np.random.default_rng(8)
    pred["index_per_point"] = rng.dirichlet(np.ones(3), size=pred["index_per_point"].shape[:2]).astype(np.float32)
    pred["W"] = rng.dirichlet(np.ones(K), size=pred["W"].shape[:2]).astype(np.float32)
    joints, cnt = metrics.joint_vote(gocs, pred["W"], pred["unitvec_per_point"], pred["heatmap_per_point"],
                                     pred["joint_axis_per_point"], pred["index_per_point"])
    total = 0
    for b in range(len(clouds)):
        ref = iou_np.joint_vote(gocs[b], pred["W"][b], pred["unitvec_per_point"][b], pred["heatmap_per_point"][b],
                                pred["joint_axis_per_point"][b], pred["index_per_point"][b], K)
        jc = np.argmax(pred["index_per_point"][b], axis=1)
        for j in range(K - 1):
            assert cnt[b, j] == int((jc == j + 1).sum())
            total += cnt[b, j]
            np.testing.assert_array_equal(joints[b][j]["l"], ref[j]["l"])          # NaN == NaN for empty joints
            np.testing.assert_array_equal(joints[b][j]["p"], ref[j]["p"])
            assert joints[b][j]["l"].dtype == np.float32 == ref[j]["l"].dtype
    assert total > 0 and (cnt[:, :2] > 100).all()
    if cat == "drawer":          # index_per_point is 3 wide (architecture.py:129): joint 3 can never be voted for
        assert (cnt[:, 2] == 0).all() and all(np.isnan(joints[b][2]["p"]).all() for b in range(len(clouds)))
