"""Mints tests/golden/similarity_ref.npz from the REFERENCE's own lib/aligning.py::estimateSimilarityTransform
(imported unmodified through oracle/ref_loader.py) with RECORDED np.random.randint draws.
    python tests/golden/make_similarity_golden.py
Cases (per part of seeded synthetic clouds with teacher predictions):
  natural    NOCS source / camera-space target: PassT >= 1, almost everything is an inlier (the reference's thresholds)
  scaled50   both sides x50: residuals straddle PassT -> selective inlier sets, the count_nonzero(index) quirk matters
  tiny       both sides x1e-3: BestResidual < StopT after the first iteration -> early stop
  garbage    target shuffled and x50: best inlier ratio < 0.1 -> four Nones
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from articulated_pose_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402

NITER = 100


def problems():
    rng = np.random.default_rng(77)
    out = []
    for cid, cat in ((0, "eyeglasses"), (1, "eyeglasses"), (2, "drawer")):
        cloud = synthetic.make_cloud(cid, cat)
        pred = synthetic.teacher_predictions(cloud)
        cls = np.argmax(pred["W"], axis=1)
        for j in range(cloud["n_parts"]):
            p = np.where(cls == j)[0]
            src = pred["nocs_per_point"][p, 3 * j:3 * j + 3].astype(np.float32)
            tgt = cloud["P"][p].astype(np.float32)
            out.append(("natural", src, tgt))
            out.append(("scaled50", src * np.float32(50), tgt * np.float32(50)))
            if j == 0:
                out.append(("tiny", src * np.float32(1e-3), tgt * np.float32(1e-3)))
                out.append(("garbage", src * np.float32(50), rng.permutation(tgt) * np.float32(50)))
    return out, rng


def main():
    _, _, al = ref_loader.load()
    probs, rng = problems()
    out = {"n": np.array(len(probs)), "numpy_version": np.array(np.__version__)}
    for k, (kind, src, tgt) in enumerate(probs):
        idx = rng.integers(0, len(src), size=(NITER, 5))
        used = []

        def stream():
            for r in idx:
                used.append(1)
                yield r

        with ref_loader.injected_randint(stream()), contextlib.redirect_stdout(io.StringIO()):
            res = al.estimateSimilarityTransform(src.astype(np.float64), tgt.astype(np.float64))
        key = "p%d_" % k
        out[key + "kind"], out[key + "src"], out[key + "tgt"], out[key + "idx"] = np.array(kind), src, tgt, idx.astype(np.int32)
        out[key + "iters"] = np.array(len(used))
        out[key + "none"] = np.array(res[0] is None)
        if res[0] is not None:
            out[key + "scales"], out[key + "R"], out[key + "t"], out[key + "T"] = res
        print(k, kind, len(src), "iters", len(used), "none" if res[0] is None else float(res[0][0]))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "similarity_ref.npz"), **out)


if __name__ == "__main__":
    main()
