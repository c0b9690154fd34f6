"""Synthetic articulated-object clouds for tests and benchmarks (no dataset ships with the reference and
there is no network): SURVEY.md section 8(d).

`make_cloud(cloud_id, category)` mimics what lib/dataset.py:251-432 (create_unit_data_from_hdf5) hands to the
network and what the pose stage later reads back from the per-cloud h5 (lib/prediction_io.py:73-92):
  P (N,3) camera-space points, cls_gt (N,), nocs_gt (N,3) part-normalised coordinates, joint_cls_gt (N,),
  joint axis per joint (canonical frame), and the ground-truth per-part similarity (scale, R, t) with
  P = scale * R @ nocs + t  -- exactly the model solver_ransac_nonlinear fits
  (evaluation/parallel_ancsh_pose.py:258-269).

`teacher_predictions(cloud)` builds network-like outputs (noisy GT) so RANSAC / the joint solve behave as
on real data.
"""
import numpy as np

CATEGORIES = {
    # part boxes: (size xyz, rest-state centre); joints: (child part, type, pivot, axis)
    "eyeglasses": dict(
        n_points=1024,
        boxes=[((0.90, 0.30, 0.06), (0.0, 0.0, 0.0)),
               ((0.06, 0.06, 0.80), (-0.45, 0.0, -0.43)),
               ((0.06, 0.06, 0.80), (0.45, 0.0, -0.43))],
        split=(0.5, 0.25, 0.25),
        joints=[(1, "revolute", (-0.45, 0.0, -0.03), (0.0, 1.0, 0.0), (0.0, np.pi / 2)),
                (2, "revolute", (0.45, 0.0, -0.03), (0.0, -1.0, 0.0), (0.0, np.pi / 2))],
    ),
    "drawer": dict(
        n_points=2048,
        boxes=[((0.60, 0.90, 0.60), (0.0, 0.0, 0.0)),
               ((0.54, 0.24, 0.56), (0.0, 0.29, 0.04)),
               ((0.54, 0.24, 0.56), (0.0, 0.0, 0.04)),
               ((0.54, 0.24, 0.56), (0.0, -0.29, 0.04))],
        split=(0.4, 0.2, 0.2, 0.2),
        joints=[(1, "prismatic", (0.0, 0.29, 0.0), (0.0, 0.0, 1.0), (0.0, 0.3)),
                (2, "prismatic", (0.0, 0.0, 0.0), (0.0, 0.0, 1.0), (0.0, 0.3)),
                (3, "prismatic", (0.0, -0.29, 0.0), (0.0, 0.0, 1.0), (0.0, 0.3))],
    ),
    # the three remaining categories of the reference (global_info.py:30-82): two parts, one revolute joint, N = 1024
    "oven": dict(
        n_points=1024,
        boxes=[((0.80, 0.60, 0.60), (0.0, 0.0, 0.0)),
               ((0.70, 0.50, 0.04), (0.0, 0.0, 0.32))],
        split=(0.65, 0.35),
        joints=[(1, "revolute", (0.0, -0.25, 0.30), (1.0, 0.0, 0.0), (0.0, np.pi / 2))],
    ),
    "laptop": dict(
        n_points=1024,
        boxes=[((0.80, 0.04, 0.55), (0.0, 0.0, 0.0)),
               ((0.80, 0.55, 0.03), (0.0, 0.295, -0.275))],
        split=(0.5, 0.5),
        joints=[(1, "revolute", (0.0, 0.02, -0.275), (1.0, 0.0, 0.0), (-np.pi / 3, np.pi / 6))],
    ),
    "washing_machine": dict(
        n_points=1024,
        boxes=[((0.65, 0.85, 0.65), (0.0, 0.0, 0.0)),
               ((0.40, 0.40, 0.05), (0.0, 0.05, 0.35))],
        split=(0.75, 0.25),
        joints=[(1, "revolute", (-0.20, 0.05, 0.33), (0.0, 1.0, 0.0), (-np.pi / 2, 0.0))],
    ),
}

# the reference's five benchmark categories (global_info.py: eyeglasses, oven, laptop, washing_machine, drawer)
ALL_CATEGORIES = ("eyeglasses", "oven", "laptop", "washing_machine", "drawer")


def n_parts(category):
    return len(CATEGORIES[category]["boxes"])


def _rodrigues(axis, angle):
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def _euler(pitch, roll, yaw):
    rx = _rodrigues((1, 0, 0), pitch)
    ry = _rodrigues((0, 1, 0), yaw)
    rz = _rodrigues((0, 0, 1), roll)
    return ry @ rx @ rz


def _sample_box_surface(rng, size, n):
    """n points uniformly on the surface of an axis-aligned box centred at 0; returns (pts, normals)."""
    sx, sy, sz = size
    areas = np.array([sy * sz, sy * sz, sx * sz, sx * sz, sx * sy, sx * sy])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    u = rng.uniform(-0.5, 0.5, size=(n, 3)) * np.array(size)
    nrm = np.zeros((n, 3))
    for f in range(6):
        m = face == f
        ax, sgn = f // 2, (1.0 if f % 2 == 0 else -1.0)
        u[m, ax] = sgn * 0.5 * size[ax]
        nrm[m, ax] = sgn
    return u, nrm


def make_cloud(cloud_id, category="eyeglasses", n_points=None, noise=0.002, seed=1234):
    cat = CATEGORIES[category]
    N = int(n_points or cat["n_points"])
    rng = np.random.default_rng(seed + int(cloud_id))
    boxes, joints = cat["boxes"], cat["joints"]
    K = len(boxes)
    lo = np.min([np.array(c) - 0.5 * np.array(s) for s, c in boxes], axis=0)
    hi = np.max([np.array(c) + 0.5 * np.array(s) for s, c in boxes], axis=0)
    obj_scale = 1.0 / np.linalg.norm(hi - lo)                         # rest-state bbox diagonal = 1 (dataset.py:351)
    R_obj = _euler(rng.uniform(np.deg2rad(-90), np.deg2rad(5)), rng.uniform(np.deg2rad(-10), np.deg2rad(10)),
                   rng.uniform(-np.pi, np.pi))
    t_obj = np.array([0.0, 0.0, -1.0]) * 1.0 + rng.uniform(-0.1, 0.1, size=3)

    # per-part articulation: x_art = Rj (x - pivot) + pivot + dj
    Rj = [np.eye(3) for _ in range(K)]
    dj = [np.zeros(3) for _ in range(K)]
    pj = [np.zeros(3) for _ in range(K)]
    joint_axis_cam = []
    for child, jtype, pivot, axis, (a0, a1) in joints:
        q = rng.uniform(a0, a1)
        ax = np.asarray(axis, np.float64)
        if jtype == "revolute":
            Rj[child] = _rodrigues(ax, q)
            pj[child] = np.asarray(pivot, np.float64)
        else:
            dj[child] = ax * q
        # The reference's joint residual is Rod(r0) u - Rod(r1) u (parallel_ancsh_pose.py:63): it is consistent with
        # the part poses only for u expressed in the canonical (rest / NOCS) frame, where R_gt[0] u == R_gt[j] u.
        joint_axis_cam.append(ax.copy())

    n_draw = 4 * N
    counts = np.maximum(1, (np.asarray(cat["split"]) * n_draw).astype(int))
    pts, nocs, cls = [], [], []
    gt = []
    for j, ((size, centre), cnt) in enumerate(zip(boxes, counts)):
        size = np.asarray(size, np.float64)
        centre = np.asarray(centre, np.float64)
        local, nrm = _sample_box_surface(rng, size, cnt)
        diag = np.linalg.norm(size)
        npcs = local / diag + 0.5                                      # part-normalised coordinates in [0,1]^3
        rest = local + centre
        art = (rest - pj[j]) @ Rj[j].T + pj[j] + dj[j]
        cam = obj_scale * (art @ R_obj.T) + t_obj
        ncam = nrm @ Rj[j].T @ R_obj.T
        vis = ncam[:, 2] > 0.0                                         # one-sided visibility (camera on +z)
        pts.append(cam[vis]); nocs.append(npcs[vis]); cls.append(np.full(vis.sum(), j))
        # ground-truth similarity nocs -> camera:  P = s R nocs + t
        s = obj_scale * diag
        R = R_obj @ Rj[j]
        t = obj_scale * (R_obj @ (Rj[j] @ (centre - pj[j] - 0.5 * diag * np.ones(3)) + pj[j] + dj[j])) + t_obj
        gt.append((s, R, t))
    pts = np.concatenate(pts); nocs = np.concatenate(nocs); cls = np.concatenate(cls)
    pts = pts + rng.normal(0.0, noise, size=pts.shape)
    n_tot = pts.shape[0]
    if n_tot < N:                                                      # tile like lib/dataset.py:290-317
        tile_n = int(N / n_tot) + 1
        pts, nocs, cls = np.concatenate([pts] * tile_n), np.concatenate([nocs] * tile_n), np.concatenate([cls] * tile_n)
        n_tot = pts.shape[0]
    perm = rng.permutation(n_tot)[:N]                                  # dataset.py:346-351
    pts, nocs, cls = pts[perm], nocs[perm], cls[perm]
    joint_cls = cls.copy()                                             # points of moving part j vote for joint j
    return {
        "category": category, "n_parts": K, "cloud_id": int(cloud_id),
        "P": pts.astype(np.float32), "nocs_gt": nocs.astype(np.float32), "cls_gt": cls.astype(np.int32),
        "joint_cls_gt": joint_cls.astype(np.int32),
        "joint_axis_gt": np.asarray(joint_axis_cam, np.float32),
        "scale_gt": np.array([g[0] for g in gt]), "R_gt": np.stack([g[1] for g in gt]),
        "t_gt": np.stack([g[2] for g in gt]),
    }


def make_batch(cloud_ids, category="eyeglasses", n_points=None):
    clouds = [make_cloud(i, category, n_points) for i in cloud_ids]
    return np.stack([c["P"] for c in clouds]), clouds


def teacher_predictions(cloud, seed=99, nocs_sigma=0.01, outlier_frac=0.10, label_noise=0.05, axis_sigma=0.05):
    """Network-like outputs built from ground truth (SURVEY.md 8d 'teacher' predictions):
    nocs_per_point (N,3K), W (N,K), joint_axis_per_point (N,3)."""
    rng = np.random.default_rng([seed, int(cloud["cloud_id"])])
    N, K = cloud["P"].shape[0], cloud["n_parts"]
    cls = cloud["cls_gt"].copy()
    flip = rng.uniform(size=N) < label_noise
    cls[flip] = rng.integers(0, K, size=flip.sum())
    W = np.full((N, K), 0.02 / max(K - 1, 1), np.float32)
    W[np.arange(N), cls] = 0.98
    nocs = np.tile(cloud["nocs_gt"], (1, K)).astype(np.float64) + rng.normal(0, nocs_sigma, size=(N, 3 * K))
    out = rng.uniform(size=N) < outlier_frac
    nocs[out] = rng.uniform(0, 1, size=(out.sum(), 3 * K))
    nocs = np.clip(nocs, 0.0, 1.0)
    axis = np.zeros((N, 3))
    for j in range(1, K):
        m = cloud["joint_cls_gt"] == j
        axis[m] = cloud["joint_axis_gt"][j - 1]
    axis += rng.normal(0, axis_sigma, size=axis.shape)
    return {"nocs_per_point": nocs.astype(np.float32), "W": W, "joint_axis_per_point": axis.astype(np.float32)}


def mixed_stream(n_clouds, categories=ALL_CATEGORIES, seed=2024, first_id=0):
    """BASELINE.json configs[4]: a shuffled stream of clouds drawn uniformly from `categories`.
    Returns [(category, cloud_id)], the same list on every rank (the ranks shard it, dist.shard_range)."""
    rng = np.random.default_rng(seed)
    cats = rng.integers(0, len(categories), size=int(n_clouds))
    return [(categories[int(c)], int(first_id + i)) for i, c in enumerate(cats)]
