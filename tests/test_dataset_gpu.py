"""ancsh_unit_data (device-side sampling / normalisation of lib/dataset.py:290-317, 346-372) against the NumPy restatement
oracle/dataset_np.py: every gathered array, the scaled coordinates and both masks bit-exact; the sapien rotation branch to
1e-6 (np.dot's f64 summation order is BLAS's)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _cloud(rng, n, K):
    return {"pts": rng.normal(size=(n, 3)).astype(np.float32), "cls": rng.integers(0, K, n).astype(np.float32),
            "heatmap": rng.random(n).astype(np.float32), "unitvec": rng.normal(size=(n, 3)).astype(np.float32),
            "orient": rng.normal(size=(n, 3)).astype(np.float32), "joint_cls": rng.integers(0, K, n).astype(np.float32),
            "nocs_p": rng.random((n, 3)).astype(np.float32), "nocs_g": rng.random((n, 3)).astype(np.float32)}


@pytest.mark.parametrize("with_rot", [False, True])
def test_unit_data_matches_oracle(with_rot):
    from scipy.spatial.transform import Rotation
    from articulated_pose_b200 import dataset
    from oracle import dataset_np
    rng = np.random.default_rng(12)
    K, num_points = 3, 1024
    sizes = [5000, 1024, 700, 37]                       # more points than needed, exactly enough, tiled 2x, tiled 28x
    clouds = [_cloud(rng, n, K) for n in sizes]
    perm = np.stack([rng.permutation(n if n >= num_points else (int(num_points / n) + 1) * n)[:num_points] for n in sizes])
    nf = rng.uniform(0.5, 2.0, len(sizes)).astype(np.float32)
    rot = Rotation.random(len(sizes), random_state=3).as_matrix() if with_rot else None
    got = dataset.subsample_normalize(clouds, perm, nf, K, rot=rot)
    for b, c in enumerate(clouds):
        ref = dataset_np.unit_data(c, perm[b], nf[b], K, num_points, rot[b] if with_rot else None)
        assert set(ref) == set(got)
        for k, v in ref.items():
            if with_rot and k in ("nocs_gt", "nocs_gt_g", "unitvec_gt", "orient_gt"):
                np.testing.assert_allclose(got[k][b], v, rtol=0, atol=1e-6, err_msg=k)
            else:
                np.testing.assert_array_equal(got[k][b], v, err_msg=k)
    assert got["mask_array"].sum() == len(sizes) * num_points
