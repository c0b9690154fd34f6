"""GPU probe: timeline (ms offsets) of forward / pose stages of consecutive pipelined steps."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from articulated_pose_b200 import _lib, synthetic
from articulated_pose_b200.network import AncshNet
from articulated_pose_b200.pipeline import AncshPipeline

class A: pass
args = A(); args.category = "eyeglasses"; args.no_baseline_net = False; args.nsample = 32
K, B = 3, 256
dev = torch.device("cuda:0")
def feats(kind, w, Pc):
    return AncshNet(w, K, mixed_pred=(kind == "ancsh"), early_split_nocs=(kind == "ancsh"), nsample=32, device=dev).features(Pc)
w_a, w_n = bench.synthetic_weight_sets(K, args, feats)
pipe = AncshPipeline(w_a, K, weights_npcs=w_n, nsample=32, niter_single=500, niter_joint=200, seed=1, device=dev)
P_host, clouds = synthetic.make_batch(range(B))
P = torch.from_numpy(P_host).to(dev)
jc = torch.from_numpy(np.stack([c["joint_cls_gt"] for c in clouds]).astype(np.int32)).to(dev)
for i in range(8):
    pipe.submit(P, jc, slot=i % pipe.N_SLOTS)
pipe.join(); torch.cuda.synchronize()
nst, npst = len(_lib.NET_STAGES), len(_lib.POSE_STAGES)
S = 8
evs = [(_lib.EventList(nst + 1), _lib.EventList(nst + 1), _lib.EventList(npst + 1)) for _ in range(S)]
base = _lib.EventList(1)
_lib.ancsh_event_record(base.arr[0], torch.cuda.current_stream().cuda_stream)
for i in range(S):
    pipe.submit(P, jc, slot=i % pipe.N_SLOTS, net_events=evs[i][0], net_b_events=evs[i][1], pose_events=evs[i][2])
pipe.join(); torch.cuda.synchronize()
def off(ev, j):
    ms = ctypes.c_float(); _lib.ancsh_event_elapsed_ms(base.arr[0], ev.arr[j], ctypes.byref(ms)); return round(ms.value, 2)
for i in range(S):
    e = evs[i]
    print("   fwdA stages:", " ".join("%s=%.2f" % (n, off(e[0], j + 1) - off(e[0], j)) for j, n in enumerate(_lib.NET_STAGES)))
    print("   fwdN stages:", " ".join("%s=%.2f" % (n, off(e[1], j + 1) - off(e[1], j)) for j, n in enumerate(_lib.NET_STAGES)))
    print("step", i, "fwdA %.2f-%.2f" % (off(e[0], 0), off(e[0], nst)), "fwdN %.2f-%.2f" % (off(e[1], 0), off(e[1], nst)),
          "pose:", " ".join("%s@%.2f" % (n[:6], off(e[2], j)) for j, n in enumerate(_lib.POSE_STAGES)), "end@%.2f" % off(e[2], npst))
