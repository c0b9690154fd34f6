"""The reference's REAL dtype path.  h5py hands `solver_ransac_nonlinear` float32 datasets (parallel_ancsh_pose.py:232-260),
so its single-part RANSAC arithmetic (rotate_pts / scale_pts / transform_pts, residual norms) runs in f32; the product and
the oracle promote the same f32 values to f64 first.  tests/golden/pose_ref_f32.npz holds the outputs
of the imported, unmodified reference on f32-typed inputs (tests/golden/make_pose_golden_f32.py: 4 clouds, 19 part / joint
problems at BASELINE's 500 / 200 hypotheses).  Bar this file asserts for both the oracle (CPU) and the CUDA path (GPU):
  * the winning hypothesis' inlier masks are IDENTICAL,
  * the refit models agree within 1e-6 absolute (measured: <= 2.7e-7 -- the f32 rounding of the reference itself),
  * per-hypothesis scores agree for >= 97 % of the non-degenerate hypotheses (a 3-sample covariance evaluated in f32 vs
    f64 can flip single points across the inlier threshold; it never changed a winner).
"""
import os

import numpy as np
import pytest

from articulated_pose_b200 import synthetic

GOLD = os.path.join(os.path.dirname(__file__), "golden", "pose_ref_f32.npz")
MODEL_ATOL = 1e-6
SCORE_AGREE = 0.97


def _cases(g, limit=None):
    for ci, c in enumerate(g["cases"][:limit]):
        cat, cid = str(c).split(":")
        cloud = synthetic.make_cloud(int(cid), cat)
        yield ci, cloud, synthetic.teacher_predictions(cloud)


def _part(cloud, pred, j):
    pidx = np.where(np.argmax(pred["W"], axis=1) == j)[0]
    return pred["nocs_per_point"][pidx, 3 * j:3 * j + 3].astype(np.float64), cloud["P"][pidx].astype(np.float64)


def _distinct(*idx_sets):
    return np.array([all(len(set(r)) == 3 for r in rows) for rows in zip(*idx_sets)])


def test_oracle_f64_contract_vs_reference_on_f32_inputs():
    from oracle import pose_np
    g = np.load(GOLD)
    th = float(g["inlier_th"])
    worst = 0.0
    for ci, cloud, pred in _cases(g, limit=2):               # two clouds keep the CPU suite short; the GPU test runs all four
        K = cloud["n_parts"]
        for j in range(K):
            k = "c%d_p%d_" % (ci, j)
            assert str(g[k + "dtype"]) == "float32"          # the reference really computed this model in f32
            src, tgt = _part(cloud, pred, j)
            m, inl, sc = pose_np.ransac_single(src, tgt, th, g[k + "idx"], return_scores=True)
            ok = _distinct(g[k + "idx"])
            assert np.mean(sc[ok] == g[k + "scores"][ok]) >= SCORE_AGREE
            np.testing.assert_array_equal(inl, g[k + "inl"])
            for a, b in ((m["rotation"], g[k + "R"]), (m["scale"], g[k + "s"]), (m["translation"], g[k + "t"])):
                worst = max(worst, float(np.abs(np.asarray(a) - b).max()))
        s0, t0 = _part(cloud, pred, 0)
        for j in range(1, K):
            k = "c%d_j%d_" % (ci, j)
            s1, t1 = _part(cloud, pred, j)
            m, inl, sc = pose_np.ransac_joint(s0, t0, s1, t1, g[k + "axis"], th, g[k + "idx0"], g[k + "idx1"], return_scores=True)
            ok = _distinct(g[k + "idx0"], g[k + "idx1"])
            assert np.mean(sc[ok] == g[k + "scores"][ok]) >= SCORE_AGREE
            assert int(np.argmax(sc)) == int(np.argmax(g[k + "scores"]))
            np.testing.assert_array_equal(inl[0], g[k + "inl0"])
            np.testing.assert_array_equal(inl[1], g[k + "inl1"])
            for f in ("rotation0", "scale0", "translation0", "rotation1", "scale1", "translation1"):
                worst = max(worst, float(np.abs(np.asarray(m[f]) - g[k + f]).max()))
    assert worst <= MODEL_ATOL, worst


@pytest.mark.gpu
def test_cuda_vs_reference_on_f32_inputs():
    from articulated_pose_b200.pose import PoseSolver
    g = np.load(GOLD)
    worst = 0.0
    for ci, cloud, pred in _cases(g):
        K = cloud["n_parts"]
        ns, nj = g["c%d_p0_idx" % ci].shape[0], g["c%d_j1_idx0" % ci].shape[0]
        solver = PoseSolver(K, niter_single=ns, niter_joint=nj, inlier_th=float(g["inlier_th"]))
        idx_s = np.stack([g["c%d_p%d_idx" % (ci, j)] for j in range(K)])[None]
        idx_0 = np.stack([g["c%d_j%d_idx0" % (ci, j)] for j in range(1, K)])[None]
        idx_1 = np.stack([g["c%d_j%d_idx1" % (ci, j)] for j in range(1, K)])[None]
        res = solver.solve(cloud["P"][None], pred["nocs_per_point"][None], pred["W"][None], pred["joint_axis_per_point"][None],
                           cloud["joint_cls_gt"][None], idx_single=idx_s, idx_joint0=idx_0, idx_joint1=idx_1)[0]
        inter = {k: v.cpu().numpy() for k, v in solver.intermediates().items() if hasattr(v, "cpu")}
        assert (res["status"] == 0).all()
        for j in range(K):
            k = "c%d_p%d_" % (ci, j)
            ok = _distinct(g[k + "idx"])
            assert np.mean(inter["single_scores"][0, j][ok] == g[k + "scores"][ok]) >= SCORE_AGREE
            np.testing.assert_array_equal(res["inliers_single"][j], g[k + "inl"])
            m = res["baseline"][j]
            for a, b in ((m["rotation"], g[k + "R"]), (m["scale"], g[k + "s"]), (m["translation"], g[k + "t"])):
                worst = max(worst, float(np.abs(np.asarray(a) - b).max()))
        for j in range(1, K):
            k = "c%d_j%d_" % (ci, j)
            ok = _distinct(g[k + "idx0"], g[k + "idx1"])
            assert np.mean(inter["joint_scores"][0, j - 1][ok] == g[k + "scores"][ok]) >= 0.95
            assert int(inter["joint_best"][0, j - 1]) == int(np.argmax(g[k + "scores"]))
            np.testing.assert_array_equal(res["inliers_joint"][j - 1][0], g[k + "inl0"])
            np.testing.assert_array_equal(res["inliers_joint"][j - 1][1], g[k + "inl1"])
            m = res["nonlinear"][j - 1]
            for f in ("rotation0", "scale0", "translation0", "rotation1", "scale1", "translation1"):
                worst = max(worst, float(np.abs(np.asarray(m[f]) - g[k + f]).max()))
    assert worst <= MODEL_ATOL, worst
