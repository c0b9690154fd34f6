"""Mints tests/golden/iou_ref.npz from the REFERENCE's own lib/d3_utils.py (iou_3d, get_3d_bbox), imported unmodified
through oracle/ref_loader.py.  Run in the container where /root/reference is mounted:
    python tests/golden/make_iou_golden.py
Cases: random oriented box pairs (overlapping, nested, disjoint, identical, degenerate zero-volume), nres = 50 and 17.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def cases(d3, rng, n):
    b1, b2 = [], []
    for i in range(n):
        e1 = rng.uniform(0.05, 1.0, 3)
        box = d3.get_3d_bbox(e1, shift=np.array([0.5, 0.5, 0.5])).transpose()
        R1, t1, s1 = random_rotation(rng), rng.uniform(-1, 1, 3), rng.uniform(0.5, 2.0)
        a = np.dot(box * s1, R1.T) + t1
        kind = i % 6
        if kind == 0:                                   # perturbed copy (the metric's usual case: pred ~ gt)
            e2 = e1 * rng.uniform(0.8, 1.2, 3)
            R2 = R1 @ random_small(rng)
            t2, s2 = t1 + rng.normal(0, 0.03, 3), s1 * rng.uniform(0.9, 1.1)
        elif kind == 1:                                 # unrelated box nearby
            e2, R2, t2, s2 = rng.uniform(0.05, 1.0, 3), random_rotation(rng), t1 + rng.normal(0, 0.3, 3), rng.uniform(0.5, 2.0)
        elif kind == 2:                                 # far away: empty intersection
            e2, R2, t2, s2 = rng.uniform(0.05, 1.0, 3), random_rotation(rng), t1 + 10.0, s1
        elif kind == 3:                                 # identical
            e2, R2, t2, s2 = e1, R1, t1, s1
        elif kind == 4:                                 # nested
            e2, R2, t2, s2 = e1 * 0.5, R1, t1, s1
        else:                                           # axis-aligned pair (grid planes parallel to faces)
            R1 = np.eye(3)
            a = box * s1 + t1
            e2, R2, t2, s2 = rng.uniform(0.05, 1.0, 3), np.eye(3), t1 + rng.normal(0, 0.1, 3), s1
        box2 = d3.get_3d_bbox(e2, shift=np.array([0.5, 0.5, 0.5])).transpose()
        b = np.dot(box2 * s2, R2.T) + t2
        b1.append(a)
        b2.append(b)
    return np.stack(b1), np.stack(b2)


def random_small(rng):
    v = rng.normal(0, 0.05, 3)
    th = np.linalg.norm(v)
    k = v / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def main():
    _, d3, _ = ref_loader.load()
    rng = np.random.default_rng(2024)
    out = {}
    for nres, n in ((50, 36), (17, 24)):
        b1, b2 = cases(d3, rng, n)
        iou = np.array([float(d3.iou_3d(b1[i], b2[i], nres=nres)) for i in range(n)])
        out["nres%d_bbox1" % nres], out["nres%d_bbox2" % nres], out["nres%d_iou" % nres] = b1, b2, iou
    # degenerate: both boxes collapsed to a point -> union == 0 -> 1
    z = np.zeros((1, 8, 3)) + 0.25
    out["degenerate_bbox"] = z
    out["degenerate_iou"] = np.array([float(d3.iou_3d(z[0], z[0]))])
    # get_3d_bbox corner order (f32 and scalar inputs)
    out["bbox_f32_in"] = np.array([0.3, 0.5, 0.7], np.float32)
    out["bbox_f32_out"] = d3.get_3d_bbox(out["bbox_f32_in"], shift=np.array([0.5, 0.5, 0.5]))
    out["bbox_scalar_out"] = d3.get_3d_bbox(0.8, 0)
    out["versions"] = np.array("numpy %s" % np.__version__)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "iou_ref.npz"), **out)
    print({k: getattr(v, "shape", None) for k, v in out.items()})
    print("iou nres50:", np.round(out["nres50_iou"], 4))


if __name__ == "__main__":
    main()
