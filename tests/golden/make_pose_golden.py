"""Mints tests/golden/pose_ref.npz by running the REFERENCE's own pose code (imported unmodified through
oracle/ref_loader.py from /root/reference -- works only in the build container) on seeded synthetic 'teacher'
predictions with RECORDED sample indices.  The file pins oracle/pose_np.py (and, through it, the CUDA path) to
the reference where /root/reference is not available (the GPU box).

    python tests/golden/make_pose_golden.py

Recorded: numpy / scipy versions (third-party arithmetic: LAPACK gesdd, MINPACK lmder, scipy Rotation).
"""
import os
import sys

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from articulated_pose_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402

N_SINGLE, N_JOINT, TH = 96, 24, 0.1


def part_data(cloud, pred, j):
    cls = np.argmax(pred["W"], axis=1)
    pidx = np.where(cls == j)[0]
    return (pred["nocs_per_point"][pidx, 3 * j:3 * j + 3].astype(np.float64), cloud["P"][pidx].astype(np.float64), pidx)


def main():
    pap, d3, al = ref_loader.load()
    out = {"numpy_version": np.__version__, "scipy_version": scipy.__version__, "inlier_th": TH}
    rng = np.random.default_rng(4242)
    cases = [("eyeglasses", 0), ("eyeglasses", 1), ("drawer", 2)]
    out["cases"] = np.array(["%s:%d" % c for c in cases])
    for ci, (cat, cid) in enumerate(cases):
        cloud = synthetic.make_cloud(cid, cat)
        pred = synthetic.teacher_predictions(cloud)
        K = cloud["n_parts"]
        for j in range(K):
            src, tgt, pidx = part_data(cloud, pred, j)
            idx = rng.integers(0, len(pidx), size=(N_SINGLE, 3))
            scores = []
            orig_ver = pap.single_transformation_verifier

            def ver(ds, model, th, _s=scores):
                s, inl = orig_ver(ds, model, th)
                _s.append(s)
                return s, inl

            ds = {"source": src, "target": tgt, "nsource": src.shape[0]}
            with ref_loader.injected_randint(iter(idx)):
                m, inl = pap.ransac(ds, pap.single_transformation_estimator, ver, TH, N_SINGLE)
            k = "c%d_p%d_" % (ci, j)
            out[k + "idx"] = idx.astype(np.int32)
            out[k + "scores"] = np.array(scores, np.int64)
            out[k + "R"], out[k + "s"], out[k + "t"], out[k + "inl"] = m["rotation"], m["scale"], m["translation"], inl
        for j in range(1, K):
            src0, tgt0, p0 = part_data(cloud, pred, 0)
            src1, tgt1, p1 = part_data(cloud, pred, j)
            jidx = np.where(cloud["joint_cls_gt"] == j)[0]
            axis = np.median(pred["joint_axis_per_point"][jidx].astype(np.float64), 0)
            i0 = rng.integers(0, len(p0), size=(N_JOINT, 3))
            i1 = rng.integers(0, len(p1), size=(N_JOINT, 3))
            stream = iter([x for pair in zip(i0, i1) for x in pair])
            scores = []
            orig_ver = pap.joint_transformation_verifier

            def jver(ds, model, th, _s=scores):
                s, inl = orig_ver(ds, model, th)
                _s.append(s)
                return s, inl

            ds = {"source0": src0, "target0": tgt0, "nsource0": len(p0), "source1": src1, "target1": tgt1,
                  "nsource1": len(p1), "joint_direction": axis}
            with ref_loader.injected_randint(stream):
                m, inl = pap.ransac(ds, pap.joint_transformation_estimator, jver, TH, N_JOINT)
            k = "c%d_j%d_" % (ci, j)
            out[k + "idx0"], out[k + "idx1"] = i0.astype(np.int32), i1.astype(np.int32)
            out[k + "axis"] = axis
            out[k + "scores"] = np.array(scores, np.float64)
            for f in ("rotation0", "scale0", "translation0", "rotation1", "scale1", "translation1"):
                out[k + f] = np.asarray(m[f])
            out[k + "inl0"], out[k + "inl1"] = inl[0], inl[1]
        # Umeyama GT pose (lib/aligning.py:580-622 via evaluation/compute_gt_pose.py:80-97)
        if al is not None:
            for j in range(K):
                m = cloud["cls_gt"] == j
                a = np.hstack([cloud["nocs_gt"][m].astype(np.float64), np.ones((m.sum(), 1))]).T
                c = np.hstack([cloud["P"][m].astype(np.float64), np.ones((m.sum(), 1))]).T
                s, r, t, rt = al.estimateSimilarityUmeyama(a, c)
                out["c%d_u%d_s" % (ci, j)], out["c%d_u%d_R" % (ci, j)], out["c%d_u%d_t" % (ci, j)] = s, r, t
    path = os.path.join(ROOT, "tests", "golden", "pose_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
