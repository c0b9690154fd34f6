"""GPU probe: error of the f32 and f16x3 network paths against the CPU oracle, per tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from articulated_pose_b200 import synthetic, weights
from articulated_pose_b200.network import AncshNet
from oracle import pnpp
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 64
P, _ = synthetic.make_batch(range(10, 14))
w = weights.synthetic_weights(3)
tr = {}
ref = pnpp.forward(P, w, 3, nsample=ns, trace=tr)
def err(a, b):
    e = np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1e-2)
    return e.max(), np.percentile(e, 99.9), e.mean()
for prec in ("f32", "f16x3"):
    net = AncshNet(w, 3, nsample=ns, precision=prec)
    out = net.forward(P)
    it = {k: v.cpu().numpy() for k, v in net.intermediates().items()}
    print("==", prec, "TERMS", os.environ.get("ANCSH_TC_TERMS", "3"))
    for name, r in (("l3_points", tr["l3_points"][:, 0]), ("l2_points_fp", tr["l2_points"]), ("l1_points_fp", tr["l1_points"])):
        print("  %-22s max %.2e  p99.9 %.2e  mean %.2e" % ((name,) + err(it[name], r)))
    for k in ref:
        print("  %-22s max %.2e  p99.9 %.2e  mean %.2e" % ((k,) + err(out[k], ref[k])))
