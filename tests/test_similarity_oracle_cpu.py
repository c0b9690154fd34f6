"""oracle/pose_np.py::estimate_similarity_transform (lib/aligning.py estimateSimilarityTransform =
set_config + getRANSACInliers + evaluateModel + estimateSimilarityUmeyama) against the reference's own code: the
committed goldens with recorded np.random.randint draws (tests/golden/similarity_ref.npz) and, when /root/reference is
mounted, the live function."""
import contextlib
import io
import os

import numpy as np
import pytest

from oracle import pose_np, ref_loader

GOLD = os.path.join(os.path.dirname(__file__), "golden", "similarity_ref.npz")


def test_matches_reference_golden():
    g = np.load(GOLD)
    kinds = set()
    for k in range(int(g["n"])):
        key = "p%d_" % k
        res = pose_np.estimate_similarity_transform(g[key + "src"], g[key + "tgt"], g[key + "idx"], return_info=True)
        info = res[4]
        kinds.add(str(g[key + "kind"]))
        assert info["iters"] == int(g[key + "iters"]), key
        assert (res[0] is None) == bool(g[key + "none"]), key
        if res[0] is not None:
            np.testing.assert_array_equal(res[0], g[key + "scales"])
            np.testing.assert_array_equal(res[1], g[key + "R"])
            np.testing.assert_array_equal(res[2], g[key + "t"])
            np.testing.assert_array_equal(res[3], g[key + "T"])
    assert kinds == {"natural", "scaled50", "tiny", "garbage"}


def test_count_nonzero_quirk_and_first_best():
    """evaluateModel counts the non-zero inlier INDICES (aligning.py:545): with every point an inlier the ratio is
    (n-1)/n, and the strict `>` keeps the first hypothesis that reaches it."""
    rng = np.random.default_rng(0)
    src = rng.uniform(0, 1, (50, 3))
    R, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    if np.linalg.det(R) < 0:
        R[:, 0] = -R[:, 0]
    tgt = 1.7 * src @ R.T + np.array([0.1, -0.2, 0.3])
    idx = rng.integers(0, 50, size=(100, 5))
    res = pose_np.estimate_similarity_transform(src, tgt, idx, return_info=True)
    assert res[4]["inlier_ratio"] == 49 / 50 and len(res[4]["inlier_idx"]) == 50
    np.testing.assert_allclose(res[0], 1.7, rtol=1e-9)
    np.testing.assert_allclose(res[1].T, R, atol=1e-9)          # reference convention: Rotation = (U Vh)^T


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not mounted")
def test_live_reference():
    _, _, al = ref_loader.load()
    rng = np.random.default_rng(5)
    for n, scale in ((40, 1.0), (300, 30.0), (120, 1e-3)):
        src = rng.uniform(0, 1, (n, 3)) * scale
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        tgt = (0.8 * src @ q.T + rng.normal(0, 0.02 * scale, (n, 3)))
        tgt[rng.random(n) < 0.2] += rng.normal(0, 0.5 * scale, 3)
        idx = rng.integers(0, n, size=(100, 5))
        with ref_loader.injected_randint(iter(idx)), contextlib.redirect_stdout(io.StringIO()):
            ref = al.estimateSimilarityTransform(src, tgt)
        got = pose_np.estimate_similarity_transform(src, tgt, idx)
        assert (ref[0] is None) == (got[0] is None)
        if ref[0] is not None:
            for a, b in zip(ref, got):
                np.testing.assert_array_equal(a, b)
