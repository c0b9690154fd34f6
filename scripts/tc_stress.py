"""Development aid: run the same forward many times and report any run-to-run difference (the kernels are
deterministic: a difference is a race)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import torch
from articulated_pose_b200 import synthetic, weights
from articulated_pose_b200.network import AncshNet

B, ns, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
P, _ = synthetic.make_batch(range(10, 10 + B), "eyeglasses")
w = weights.synthetic_weights(3, True, True, seed=7)
net = AncshNet(w, 3, nsample=ns)
Pd = torch.from_numpy(P).cuda()
out = net.alloc_outputs(B, P.shape[1])
ref = None
bad = 0
for it in range(iters):
    net.forward_device(Pd, out)
    torch.cuda.synchronize()
    cur = {k: v.clone() for k, v in net.intermediates().items() if v.dtype == torch.float32}
    cur.update({"out_" + k: v.clone() for k, v in out.items()})
    if ref is None:
        ref = cur
        continue
    for k in ref:
        if not torch.equal(ref[k], cur[k]):
            d = (ref[k] - cur[k]).abs()
            idx = np.unravel_index(int(d.argmax()), d.shape)
            n = int((d > 0).sum())
            print("iter %d: %s differs: %d elements, max %.3e at %s" % (it, k, n, float(d.max()), idx))
            bad += 1
print("B=%d ns=%d: %d iterations, %d differing tensors" % (B, ns, iters, bad))
