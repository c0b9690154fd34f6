"""oracle/iou_np.py -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

NumPy restatement of the metric kernels behind the pose stage:
    get_3d_bbox      lib/d3_utils.py:7-38
    pts_inside_box   lib/d3_utils.py:40-53
    iou_3d           lib/d3_utils.py:55-69
    amodal extents   evaluation/compute_miou.py:187,196-200
    part boxes / IoU evaluation/compute_miou.py:19-24 (compose_rt, f32), :222-231
Pinned against the reference's own d3_utils.iou_3d imported unmodified (tests/test_iou_oracle_cpu.py, live when
/root/reference is mounted) and against tests/golden/iou_ref.npz minted from it (tests/golden/make_iou_golden.py).
The dot products are written out left to right (the reference calls np.matmul on an (n,3)x(3,1) product whose BLAS
summation order is not specified); the goldens show the counts agree.
"""
import warnings

import numpy as np


def get_3d_bbox(scale, shift=0):
    """(3,8) corners of the axis-aligned box of extents `scale` about `shift` (d3_utils.py:7-38).  Like np.array over
    np.float32 scalars, the half extents keep the dtype of `scale`; adding an f64 `shift` promotes afterwards."""
    s = np.asarray(scale)
    h = (s / 2) if s.ndim else np.full(3, s / 2)
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1])
    sy = np.array([1, 1, 1, 1, -1, -1, -1, -1])
    sz = np.array([1, -1, 1, -1, 1, -1, 1, -1])
    box = np.stack([sx * h[0], sy * h[1], sz * h[2]], 1).astype(h.dtype if h.dtype.kind == "f" else np.float64)
    return (box + shift).transpose()


def _dot3(up, u):
    return (up[:, 0] * u[0] + up[:, 1] * u[1]) + up[:, 2] * u[2]


def pts_inside_box(pts, bbox):
    """(n,) bool: strictly inside the oriented box spanned at corner 4 by corners 5, 7, 0 (d3_utils.py:40-53)."""
    bbox = np.asarray(bbox, np.float64)
    u = [bbox[5] - bbox[4], bbox[7] - bbox[4], bbox[0] - bbox[4]]
    up = pts - bbox[4].reshape(1, 3)
    flag = np.ones(len(pts), bool)
    for uk in u:
        p = _dot3(up, uk)
        flag &= (p > 0) & (p < ((uk[0] * uk[0] + uk[1] * uk[1]) + uk[2] * uk[2]))
    return flag


def iou_counts(bbox1, bbox2, nres=50):
    both = np.concatenate((np.asarray(bbox1, np.float64), np.asarray(bbox2, np.float64)), 0)
    bmin, bmax = both.min(0), both.max(0)
    xs, ys, zs = (np.linspace(bmin[d], bmax[d], nres) for d in range(3))
    g = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), -1).reshape(-1, 3)      # itertools.product order (:60)
    f1, f2 = pts_inside_box(g, bbox1), pts_inside_box(g, bbox2)
    return int(np.sum(f1 & f2)), int(np.sum(f1 | f2))


def iou_3d(bbox1, bbox2, nres=50):
    inter, union = iou_counts(bbox1, bbox2, nres)
    return 1 if union == 0 else inter / float(union)


def amodal_extent(nocs_pred, mask_pred, n_parts):
    """(K,3) f32 extents + (K,) counts (compute_miou.py:187,196-199); NaN for an empty part."""
    cls = np.argmax(mask_pred, axis=1)
    ext = np.full((n_parts, 3), np.nan, np.float32)
    cnt = np.zeros(n_parts, np.int32)
    for j in range(n_parts):
        idx = np.where(cls == j)[0]
        cnt[j] = len(idx)
        if len(idx):
            centered = nocs_pred[idx, 3 * j:3 * (j + 1)] - 0.5
            ext[j] = 2 * np.max(abs(centered), axis=0)
    return ext, cnt


def compose_rt(rotation, translation):
    rt = np.zeros((4, 4), dtype=np.float32)                                       # compute_miou.py:19-24 (f32!)
    rt[:3, :3] = rotation[:3, :3]
    rt[:3, 3] = translation
    rt[3, 3] = 1
    return rt


def part_boxes(extent, s, r, t):
    """World-frame corners (8,3) of one part: get_3d_bbox(extent, shift=.5) * s, rotated and shifted by the f32
    compose_rt(r, t) (compute_miou.py:200,224-229)."""
    bb = get_3d_bbox(extent, shift=np.array([1 / 2, 1 / 2, 1 / 2])).transpose() * s
    rt = compose_rt(np.asarray(r), np.asarray(t))
    return np.dot(bb, rt[:3, :3].T) + rt[:3, 3]


def joint_vote(gocs, mask_pred, unitvec_pred, heatmap_pred, orient_pred, index_per_point, num_parts, thres_r=0.2):
    """evaluation/eval_joint_params.py:139-140,153-166,178-190 for one cloud (the script body is not importable: this is
    a line-by-line restatement of its NumPy calls, "parity unpinned" beyond that).  Returns the list `joints['pred']`:
    [{'l': median axis (3,), 'p': median joint point (3,)} for j in 1..num_parts-1]."""
    joint_cls_pred = np.argmax(index_per_point, axis=1)
    cls_per_pt_pred = np.argmax(mask_pred, axis=1)
    gn_final = np.zeros((gocs.shape[0], 3), gocs.dtype)
    for j in range(num_parts):
        idx = np.where(cls_per_pt_pred == j)[0]
        gn_final[idx, :] = gocs[idx, :3] if gocs.shape[1] == 3 else gocs[idx, j * 3:j * 3 + 3]
    out = []
    with np.errstate(all="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore")                   # np.median of an empty slice warns and returns NaN
        for j in range(1, num_parts):
            offset = unitvec_pred * (1 - heatmap_pred.reshape(-1, 1)) * thres_r
            joint_pts = gn_final + offset
            idx = np.where(joint_cls_pred == j)[0]
            out.append({"l": np.median(orient_pred[idx], axis=0), "p": np.median(joint_pts[idx], axis=0)})
    return out
