"""GPU probe: are the slow LM solves the degenerate (repeated-index) hypotheses?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from articulated_pose_b200 import synthetic
from articulated_pose_b200.pose import PoseSolver
B = 64
clouds = [synthetic.make_cloud(i) for i in range(B)]
preds = [synthetic.teacher_predictions(c) for c in clouds]
jc = np.stack([c["joint_cls_gt"] for c in clouds]).astype(np.int32)
P = np.stack([c["P"] for c in clouds]); nocs = np.stack([p["nocs_per_point"] for p in preds])
W = np.stack([p["W"] for p in preds]); ax = np.stack([p["joint_axis_per_point"] for p in preds])
s = PoseSolver(3, niter_single=100, niter_joint=200, seed=1)
res = s.solve(P, nocs, W, ax, jc)
inter = {k: v.cpu().numpy() for k, v in s.intermediates().items()}
cnt = np.stack([r["part_count"] for r in res])
K = 3
i0 = s.sample_indices(1, np.repeat(cnt[:, :1], K - 1, 1).reshape(-1), 200).reshape(B, K - 1, 200, 3)
i1 = s.sample_indices(2, cnt[:, 1:].reshape(-1), 200).reshape(B, K - 1, 200, 3)
dup = np.zeros((B, K - 1, 200), bool)
for b in range(B):
    for j in range(K - 1):
        for h in range(200):
            dup[b, j, h] = len(set(i0[b, j, h])) < 3 or len(set(i1[b, j, h])) < 3
nf = inter["joint_nfev"]
print("dup frac", dup.mean())
print("nfev dup: p50/p90/max", np.percentile(nf[dup], [50, 90, 100]))
print("nfev non-dup: p50/p90/p99/p99.9/max", np.percentile(nf[~dup], [50, 90, 99, 99.9, 100]))
print("non-dup with nfev>60:", (nf[~dup] > 60).sum(), "of", (~dup).sum())
sc = inter["joint_scores"]
print("score of slow non-dup vs all: ", np.median(sc[(~dup) & (nf > 60)]) if ((~dup) & (nf > 60)).any() else None, np.median(sc))
