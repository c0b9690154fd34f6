"""Generates tests/golden/ops_ref_b200.npz by running the REFERENCE's own CUDA kernels
(oracle/_ref/libref_tfops.so = /root/reference/pointnet_plusplus/utils/tf_ops/{sampling,grouping}/*_g.cu
compiled for sm_100a by oracle/build.py) on a B200.  Run on the GPU box:

    python tests/golden/make_ops_golden.py gpurun_out/ops_ref_b200.npz

then copy the file into tests/golden/.  Inputs are the seeded synthetic eyeglasses clouds.
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from articulated_pose_b200 import synthetic  # noqa: E402
from oracle import build as obuild  # noqa: E402


def main(out_path):
    ref = ctypes.CDLL(obuild.REF_TFOPS_SO)
    P, _ = synthetic.make_batch(range(6))
    P[4, 512:] = P[4, :512]                      # one tiled cloud: exact distance ties
    b, n, _ = P.shape
    m, ns, radius = 512, 64, 0.2
    vp = ctypes.c_void_p
    d = torch.from_numpy(P).cuda()
    temp = torch.empty(32 * n, dtype=torch.float32, device="cuda")
    fps = torch.empty((b, m), dtype=torch.int32, device="cuda")
    assert ref.ref_fps(b, n, m, vp(d.data_ptr()), vp(temp.data_ptr()), vp(fps.data_ptr()), 1) == 0
    new_xyz = torch.empty((b, m, 3), dtype=torch.float32, device="cuda")
    assert ref.ref_gather_point(b, n, m, vp(d.data_ptr()), vp(fps.data_ptr()), vp(new_xyz.data_ptr()), 1) == 0
    idx = torch.zeros((b, m, ns), dtype=torch.int32, device="cuda")
    cnt = torch.zeros((b, m), dtype=torch.int32, device="cuda")
    assert ref.ref_ball_query(b, n, m, ctypes.c_float(radius), ns, vp(d.data_ptr()), vp(new_xyz.data_ptr()),
                              vp(idx.data_ptr()), vp(cnt.data_ptr()), 1) == 0
    np.savez_compressed(out_path, xyz=P, fps_idx=fps.cpu().numpy(), new_xyz=new_xyz.cpu().numpy(),
                        ball_idx=idx.cpu().numpy().astype(np.int16).astype(np.int32), ball_cnt=cnt.cpu().numpy(),
                        radius=np.float32(radius), gpu=torch.cuda.get_device_name(0))
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ops_ref_b200.npz")
