"""GPU probe: error of the f16x3 tensor-core path and of the exact-f32 CUDA-core path against the oracle under the
layer-rescaling stress of tests/test_network_gpu.py::test_fp16_split_survives_wide_dynamic_range."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from articulated_pose_b200 import synthetic, weights
from articulated_pose_b200.network import AncshNet
from oracle import pnpp
from tests.test_network_gpu import stress_weights
K, ns = 3, 32
w = stress_weights(K)
P, _ = synthetic.make_batch(range(300, 303))
ref = pnpp.forward(P, w, K, nsample=ns)
for prec in ("f16x3", "f32"):
    got = AncshNet(w, K, nsample=ns, precision=prec).forward(P)
    errs = {k: float((np.abs(got[k].astype(np.float64) - ref[k]) / np.maximum(np.abs(ref[k]), 1e-2)).max()) for k in ref}
    print(prec, {k: "%.2e" % v for k, v in errs.items()})
