"""GPU probe: where does the tensor-core path's extra error come from?  Per-tensor error of f32 / f16x3 paths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from articulated_pose_b200 import synthetic, weights
from articulated_pose_b200.network import AncshNet
from oracle import pnpp
ns = 32
P, _ = synthetic.make_batch(range(0, 6))
w = weights.synthetic_weights(3)
tr = {}
ref = pnpp.forward(P, w, 3, nsample=ns, trace=tr)
def err(a, b, floor=1e-2):
    e = np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), floor)
    return e.max(), e.mean()
for prec in ("f32", "f16x3"):
    net = AncshNet(w, 3, nsample=ns, precision=prec)
    out = net.forward(P)
    it = {k: v.cpu().numpy() for k, v in net.intermediates().items()}
    feat = net.features(P)
    print("==", prec)
    rows = [("l1_points(sa1)", it["l1_points"], None), ("l2_points(sa2)", it["l2_points"], None), ("l3_points", it["l3_points"], tr["l3_points"][:, 0]),
            ("l2_fp", it["l2_points_fp"], tr["l2_points"]), ("l1_fp", it["l1_points_fp"], tr["l1_points"]), ("net", feat, tr["net"])]
    for name, a, r in rows:
        if r is not None:
            print("  %-18s max %.2e mean %.2e   |ref| mean %.3f max %.2f" % ((name,) + err(a, r) + (np.abs(r).mean(), np.abs(r).max())))
    worst = max(((k,) + err(out[k], ref[k]) for k in ref), key=lambda t: t[1])
    print("  worst output:", worst, " abs err max", max(np.abs(out[k] - ref[k]).max() for k in ref))
