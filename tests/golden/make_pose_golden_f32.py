"""Mints tests/golden/pose_ref_f32.npz: the REFERENCE's own pose code (imported unmodified through oracle/ref_loader.py)
run on **float32-typed** inputs, which is what it really sees: h5py hands `solver_ransac_nonlinear` f32 datasets
(`nocs_pred[partidx[j], 3*j:3*(j+1)]`, `f['P'][partidx[j], :3]`, parallel_ancsh_pose.py:232-260), so its single-part
RANSAC (rotate_pts / scale_pts / transform_pts, the residual norms) runs in f32 while scipy's LM promotes to f64.
The product's contract is f64 arithmetic on the same f32 values; this file pins how far the
two are apart: tests/test_pose_f32_reference.py asserts equal inlier sets and models within 1e-5.

    python tests/golden/make_pose_golden_f32.py
"""
import os
import sys

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from articulated_pose_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402

N_SINGLE, N_JOINT, TH = 500, 200, 0.1      # BASELINE.json configs[2] hypothesis counts
CASES = [("eyeglasses", 30), ("eyeglasses", 31), ("drawer", 32), ("laptop", 33)]


def part_data32(cloud, pred, j):
    cls = np.argmax(pred["W"], axis=1)
    pidx = np.where(cls == j)[0]
    return (pred["nocs_per_point"][pidx, 3 * j:3 * j + 3].astype(np.float32), cloud["P"][pidx].astype(np.float32), pidx)


def main():
    pap, d3, al = ref_loader.load()
    out = {"numpy_version": np.__version__, "scipy_version": scipy.__version__, "inlier_th": TH,
           "cases": np.array(["%s:%d" % c for c in CASES])}
    rng = np.random.default_rng(777)
    for ci, (cat, cid) in enumerate(CASES):
        cloud = synthetic.make_cloud(cid, cat)
        pred = synthetic.teacher_predictions(cloud)
        K = cloud["n_parts"]
        for j in range(K):
            src, tgt, pidx = part_data32(cloud, pred, j)
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
            out[k + "dtype"] = str(np.asarray(m["rotation"]).dtype)
            out[k + "R"], out[k + "s"], out[k + "t"], out[k + "inl"] = (np.asarray(m["rotation"], np.float64), np.float64(m["scale"]),
                                                                         np.asarray(m["translation"], np.float64), inl)
        for j in range(1, K):
            src0, tgt0, p0 = part_data32(cloud, pred, 0)
            src1, tgt1, p1 = part_data32(cloud, pred, j)
            jidx = np.where(cloud["joint_cls_gt"] == j)[0]
            axis = np.median(pred["joint_axis_per_point"][jidx].astype(np.float32), 0)      # f32, :295
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
            out[k + "axis"] = axis.astype(np.float64)
            out[k + "scores"] = np.array(scores, np.float64)
            for f in ("rotation0", "scale0", "translation0", "rotation1", "scale1", "translation1"):
                out[k + f] = np.asarray(m[f], np.float64)
            out[k + "inl0"], out[k + "inl1"] = inl[0], inl[1]
    path = os.path.join(ROOT, "tests", "golden", "pose_ref_f32.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
