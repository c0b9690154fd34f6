"""Development aid: dump the network intermediates of one configuration (python tc_compare.py dump out.npz B ns) and
compare two dumps (python tc_compare.py cmp a.npz b.npz)."""
import sys
import numpy as np

sys.path.insert(0, ".")


def dump(path, B, ns):
    from articulated_pose_b200 import synthetic, weights
    from articulated_pose_b200.network import AncshNet
    P, _ = synthetic.make_batch(range(10, 10 + B), "eyeglasses")
    if B > 2:
        P[1, P.shape[1] // 2:] = P[1, :P.shape[1] // 2]
    w = weights.synthetic_weights(3, True, True, seed=7)
    net = AncshNet(w, 3, nsample=ns)
    got = net.forward(P)
    inter = {k: v.cpu().numpy() for k, v in net.intermediates().items()}
    inter.update({"out_" + k: v for k, v in got.items()})
    np.savez(path, **inter)


def cmp(a, b):
    A, Bz = np.load(a), np.load(b)
    for k in A.files:
        x, y = A[k].astype(np.float64), Bz[k].astype(np.float64)
        if x.dtype.kind != "f":
            continue
        err = np.abs(x - y) / np.maximum(np.abs(y), 1e-2)
        idx = np.unravel_index(err.argmax(), err.shape)
        bad = (err > 5e-4).sum()
        print("%-28s max rel %.3e at %s  (%d > 5e-4 of %d)  a=%.6g b=%.6g" % (k, err.max(), idx, bad, err.size, x[idx], y[idx]))
        if bad and x.ndim == 3:
            rows = np.unique(np.argwhere(err > 5e-4)[:, 1])
            cols = np.unique(np.argwhere(err > 5e-4)[:, 2])
            print("      bad rows %s ... cols %s ..." % (rows[:12], cols[:16]))


if __name__ == "__main__":
    if sys.argv[1] == "dump":
        dump(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
    else:
        cmp(sys.argv[2], sys.argv[3])
