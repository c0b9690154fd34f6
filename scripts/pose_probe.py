"""GPU probe: per-stage time of the pose solver on teacher predictions vs degraded predictions, LM nfev stats."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from articulated_pose_b200 import _lib, synthetic
from articulated_pose_b200.pose import PoseSolver

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
clouds = [synthetic.make_cloud(i) for i in range(B)]
jc = torch.from_numpy(np.stack([c["joint_cls_gt"] for c in clouds]).astype(np.int32)).cuda()
P = torch.from_numpy(np.stack([c["P"] for c in clouds])).cuda()
for label, kw in (("teacher", {}), ("noisy", dict(nocs_sigma=0.08, outlier_frac=0.4)), ("garbage", dict(nocs_sigma=0.5, outlier_frac=0.9))):
    preds = [synthetic.teacher_predictions(c, **kw) for c in clouds]
    nocs = torch.from_numpy(np.stack([p["nocs_per_point"] for p in preds])).cuda()
    W = torch.from_numpy(np.stack([p["W"] for p in preds])).cuda()
    ax = torch.from_numpy(np.stack([p["joint_axis_per_point"] for p in preds])).cuda()
    for hyp in (500, 10000):
        s = PoseSolver(3, niter_single=hyp, niter_joint=200, seed=1)
        ev = _lib.EventList(len(_lib.POSE_STAGES) + 1)
        for _ in range(2):
            s.solve_device(P, nocs, W, ax, jc, stage_events=ev)
        torch.cuda.synchronize()
        ms = {n: round(ev.elapsed_ms(i, i + 1), 3) for i, n in enumerate(_lib.POSE_STAGES)}
        nf = s.intermediates()["joint_nfev"].cpu().numpy().ravel()
        print(label, "hyp", hyp, "B", B, ms, "nfev p50/p99/max", np.percentile(nf, [50, 99, 100]).tolist(), flush=True)
