"""Device-side sampling / normalisation of `create_unit_data_from_hdf5` (lib/dataset.py:290-317, 346-372):
the step between the raw per-cloud arrays of the dataset h5 files and the network input batch.  The reference draws
`perm = np.random.permutation(n_total_points)` per cloud from the unseeded global RNG; here `perm` is an argument (like the
RANSAC samples), so that a CPU checker can replay it.  Reading the h5 files, URDF parsing and the test-group filters stay
with the caller (out of scope).  PyTorch tensors are device-memory containers only."""
import ctypes

import numpy as np
import torch

from . import _lib

_IN_WIDTH = {"pts": 3, "cls": 0, "heatmap": 0, "unitvec": 3, "orient": 3, "joint_cls": 0, "nocs_p": 3, "nocs_g": 3}
_OUT_OF = {"P": "pts", "cls_gt": "cls", "nocs_gt": "nocs_p", "nocs_gt_g": "nocs_g", "heatmap_gt": "heatmap",
           "unitvec_gt": "unitvec", "orient_gt": "orient", "joint_cls_gt": "joint_cls", "joint_cls_mask": "joint_cls"}


def subsample_normalize(clouds, perm, norm_factor, n_parts, num_points=None, rot=None, device="cuda:0"):
    """clouds: list of per-cloud dicts with `pts` (n,3), `cls` (n,) and optionally `heatmap`, `unitvec`, `orient`, `joint_cls`,
    `nocs_p`, `nocs_g` (the arrays create_unit_data_from_hdf5 reads, any n per cloud); perm: (B,num_points) int positions into
    each cloud tiled to >= num_points points; norm_factor: (B,); rot: None or (B,3,3) sapien joint_rpy rotations.
    Returns the batch dict of lib/dataset.py:379-391 ('P', 'cls_gt', 'mask_array', 'nocs_gt', 'nocs_gt_g', 'heatmap_gt',
    'unitvec_gt', 'orient_gt', 'joint_cls_gt', 'joint_cls_mask') as host float32 arrays."""
    dev = torch.device(device)
    B = len(clouds)
    perm = np.ascontiguousarray(perm, np.int32)
    num_points = int(num_points or perm.shape[1])
    if perm.shape != (B, num_points):
        raise ValueError("perm must be (B, num_points)")
    n_total = np.array([len(c["pts"]) for c in clouds], np.int32)
    n_max = int(n_total.max())
    keys = [k for k in _IN_WIDTH if all(k in c and c[k] is not None for c in clouds)]
    if "pts" not in keys or "cls" not in keys:
        raise ValueError("every cloud needs pts and cls")
    din = {}
    for k in keys:
        w = _IN_WIDTH[k]
        a = np.zeros((B, n_max, w) if w else (B, n_max), np.float32)
        for b, c in enumerate(clouds):
            a[b, :n_total[b]] = np.asarray(c[k], np.float32).reshape((n_total[b], w) if w else (n_total[b],))
        din[k] = torch.from_numpy(a).to(dev)
    dout = {"mask_array": torch.empty((B, num_points, int(n_parts)), dtype=torch.float32, device=dev)}
    for ko, ki in _OUT_OF.items():
        if ki in din:
            w = _IN_WIDTH[ki]
            dout[ko] = torch.empty((B, num_points, w) if w else (B, num_points), dtype=torch.float32, device=dev)
    uin, uout = _lib.UnitIn(), _lib.UnitOut()
    for k in _lib.UNIT_IN_FIELDS:
        setattr(uin, k, din[k].data_ptr() if k in din else None)
    for k in _lib.UNIT_OUT_FIELDS:
        setattr(uout, k, dout[k].data_ptr() if k in dout else None)
    d_nt = torch.from_numpy(n_total).to(dev)
    d_perm = torch.from_numpy(perm).to(dev)
    d_nf = torch.from_numpy(np.ascontiguousarray(norm_factor, np.float32).reshape(B)).to(dev)
    d_rot = None if rot is None else torch.from_numpy(np.ascontiguousarray(rot, np.float64).reshape(B, 9)).to(dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.ancsh_unit_data(B, n_max, num_points, int(n_parts), d_nt.data_ptr(), d_perm.data_ptr(), d_nf.data_ptr(),
                                        d_rot.data_ptr() if d_rot is not None else None, ctypes.byref(uin), ctypes.byref(uout),
                                        torch.cuda.current_stream().cuda_stream), "ancsh_unit_data")
    return {k: v.cpu().numpy() for k, v in dout.items()}
