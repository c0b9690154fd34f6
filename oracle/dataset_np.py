"""oracle/dataset_np.py -- TEST INFRASTRUCTURE ONLY.

NumPy restatement of the sampling / normalisation tail of `Dataset.create_unit_data_from_hdf5` (lib/dataset.py:290-317 tiling,
:346-372 permutation gather, :351 norm factor, :356-360 masks, :369-377 sapien rotation, :379-391 result dict, nocs_type 'A').
The function itself needs the dataset's h5 / URDF files (absent here) and is the body of a 180-line method, so it cannot be
imported and run: **parity unpinned** beyond this statement-for-statement restatement.  `perm` replaces
np.random.permutation(n_total_points) (:346)."""
import numpy as np


def unit_data(cloud, perm, norm_factor, n_parts, num_points, rot_mat=None):
    pts, cls = np.asarray(cloud["pts"], np.float32), np.asarray(cloud["cls"])
    arrs = {k: np.asarray(cloud[k], np.float32) for k in ("heatmap", "unitvec", "orient", "joint_cls", "nocs_p", "nocs_g") if k in cloud}
    n_total = pts.shape[0]
    if n_total < num_points:                                              # :290-317
        tile_n = int(num_points / n_total) + 1
        n_total = tile_n * n_total
        pts, cls = np.concatenate([pts] * tile_n, 0), np.concatenate([cls] * tile_n, 0)
        arrs = {k: np.concatenate([v] * tile_n, 0) for k, v in arrs.items()}
    perm = np.asarray(perm)[:num_points]
    cls_arr = cls[perm]                                                   # :347
    pts_arr = pts[perm] * np.float32(norm_factor)                         # :351
    out = {"P": pts_arr, "cls_gt": cls_arr.astype(np.float32)}
    mask_array = np.zeros([num_points, n_parts], dtype=np.float32)        # :327
    mask_array[np.arange(num_points), cls_arr.astype(np.int8)] = 1.00     # :360
    out["mask_array"] = mask_array
    g = {k: v[perm] for k, v in arrs.items()}                             # :352-367
    if "joint_cls" in g:
        m = np.zeros((num_points,), np.float32)                           # :356-358
        m[np.where(g["joint_cls"] > 0)[0]] = 1.00
        out["joint_cls_gt"], out["joint_cls_mask"] = g["joint_cls"].astype(np.float32), m
    if rot_mat is not None:                                               # :369-377
        for k in ("nocs_p", "nocs_g"):
            if k in g:
                g[k] = np.dot(g[k] - 0.5, rot_mat.T) + 0.5
        for k in ("unitvec", "orient"):
            if k in g:
                g[k] = np.dot(g[k], rot_mat.T)
    for ko, ki in (("nocs_gt", "nocs_p"), ("nocs_gt_g", "nocs_g"), ("heatmap_gt", "heatmap"), ("unitvec_gt", "unitvec"),
                   ("orient_gt", "orient")):
        if ki in g:
            out[ko] = g[ki].astype(np.float32)
    return out
