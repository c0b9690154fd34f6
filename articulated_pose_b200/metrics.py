"""Host mirror of the metric kernels that follow the pose stage: same names, argument meaning and
return values as the reference's lib/d3_utils.py (get_3d_bbox :7-38, iou_3d :55-69) and the per-cloud loop of
evaluation/compute_miou.py:187-231, with the heavy parts (amodal extents over the predicted parts, the nres^3 grid
IoU) on the GPU through ancsh_amodal_extent / ancsh_box_iou_3d.  No CPU fallback: without the CUDA library the import
of _lib raises."""
import numpy as np
import torch

from . import _lib


def get_3d_bbox(scale, shift=0):
    """d3_utils.py:7-38: (3,8) corners, order (+,+,+) (+,+,-) (-,+,+) (-,+,-) (+,-,+) (+,-,-) (-,-,+) (-,-,-) times
    scale/2, plus shift.  Half extents keep scale's float dtype (np.array over np.float32 scalars) before the shift."""
    s = np.asarray(scale)
    h = (s / 2) if s.ndim else np.full(3, s / 2)
    sign = np.array([[1, 1, 1], [1, 1, -1], [-1, 1, 1], [-1, 1, -1], [1, -1, 1], [1, -1, -1], [-1, -1, 1], [-1, -1, -1]])
    box = (sign * h[None, :]).astype(h.dtype if h.dtype.kind == "f" else np.float64)
    return (box + shift).transpose()


def iou_3d_batch(bbox1, bbox2, nres=50, device="cuda:0", counts=False):
    """iou_3d for (n,8,3) box pairs in one launch -> (n,) f64 [, intersect (n,), union (n,) int32]."""
    dev = torch.device(device)
    a = torch.from_numpy(np.ascontiguousarray(bbox1, np.float64).reshape(-1, 8, 3)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(bbox2, np.float64).reshape(-1, 8, 3)).to(dev)
    if a.shape != b.shape:
        raise ValueError("bbox1 / bbox2 must both be (n,8,3): %s vs %s" % (tuple(a.shape), tuple(b.shape)))
    n = a.shape[0]
    iou = torch.empty(n, dtype=torch.float64, device=dev)
    inter = torch.empty(n, dtype=torch.int32, device=dev)
    uni = torch.empty(n, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.ancsh_box_iou_3d(n, int(nres), a.data_ptr(), b.data_ptr(), iou.data_ptr(), inter.data_ptr(),
                                         uni.data_ptr(), torch.cuda.current_stream().cuda_stream), "ancsh_box_iou_3d")
    if counts:
        return iou.cpu().numpy(), inter.cpu().numpy(), uni.cpu().numpy()
    return iou.cpu().numpy()


def iou_3d(bbox1, bbox2, nres=50, device="cuda:0"):
    """Drop-in for d3_utils.iou_3d(bbox1, bbox2, nres=50): two (8,3) corner arrays -> IoU (1 when the union is empty)."""
    return float(iou_3d_batch(np.asarray(bbox1)[None], np.asarray(bbox2)[None], nres, device)[0])


def amodal_extent(nocs_pred, mask_pred, device="cuda:0"):
    """compute_miou.py:187,196-199 for a batch: nocs_pred (B,N,3K) f32, mask_pred (B,N,K) f32 (instance_per_point) ->
    scale_pred (B,K,3) f32 = 2 max |nocs - 0.5| over each predicted part, part sizes (B,K) int32."""
    dev = torch.device(device)
    q = nocs_pred if torch.is_tensor(nocs_pred) else torch.from_numpy(np.ascontiguousarray(nocs_pred, np.float32))
    w = mask_pred if torch.is_tensor(mask_pred) else torch.from_numpy(np.ascontiguousarray(mask_pred, np.float32))
    q, w = q.to(dev).contiguous(), w.to(dev).contiguous()
    B, N, K = w.shape
    if tuple(q.shape) != (B, N, 3 * K) or q.dtype != torch.float32 or w.dtype != torch.float32:
        raise ValueError("nocs_pred must be f32 (B,N,3K) and mask_pred f32 (B,N,K)")
    ext = torch.empty((B, K, 3), dtype=torch.float32, device=dev)
    cnt = torch.empty((B, K), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.ancsh_amodal_extent(B, N, K, q.data_ptr(), w.data_ptr(), ext.data_ptr(), cnt.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "ancsh_amodal_extent")
    return ext.cpu().numpy(), cnt.cpu().numpy()


def compose_rt(rotation, translation):
    """compute_miou.py:19-24 (note: f32, rotation NOT transposed -- unlike lib/aligning.py:105-110)."""
    rt = np.zeros((4, 4), dtype=np.float32)
    rt[:3, :3] = np.asarray(rotation)[:3, :3]
    rt[:3, 3] = translation
    rt[3, 3] = 1
    return rt


def part_box(extent, s, rt):
    """World-frame (8,3) corners of a part box: get_3d_bbox(extent, shift=.5).T * s through the 4x4 rt
    (compute_miou.py:200,209,224-229)."""
    bb = get_3d_bbox(extent, shift=np.array([1 / 2, 1 / 2, 1 / 2])).transpose() * s
    return np.dot(bb, rt[:3, :3].T) + rt[:3, 3]


def part_ious(nocs_pred, mask_pred, rotation, translation, scale, rt_gt, s_gt, extent_gt, nres=50, device="cuda:0"):
    """The IoU loop of compute_miou.py:196-231 for B clouds at once.
    nocs_pred (B,N,3K), mask_pred (B,N,K): network outputs (h5 'nocs_per_point', 'instance_per_point');
    rotation (B,K,3,3), translation (B,K,3), scale (B,K): the pose stage's result for one key ('baseline' / 'nonlinear');
    rt_gt (B,K,4,4), s_gt (B,K): GT part poses (compute_gt_pose); extent_gt (B,K,3): GT box extents in NOCS (:130-140).
    Returns iou (B,K) f64 (NaN where the predicted part is empty), scale_pred (B,K,3)."""
    ext, cnt = amodal_extent(nocs_pred, mask_pred, device)
    B, K = cnt.shape
    gt, pr = np.zeros((B, K, 8, 3)), np.zeros((B, K, 8, 3))
    for b in range(B):
        for j in range(K):
            gt[b, j] = part_box(np.asarray(extent_gt[b][j]), s_gt[b][j], np.asarray(rt_gt[b][j]))
            pr[b, j] = part_box(ext[b, j], scale[b][j], compose_rt(rotation[b][j], translation[b][j]))
    ok = np.isfinite(pr).all(axis=(2, 3))
    iou = iou_3d_batch(gt.reshape(-1, 8, 3), np.nan_to_num(pr).reshape(-1, 8, 3), nres, device).reshape(B, K)
    iou[~ok] = np.nan
    return iou, ext


def joint_vote(gocs, mask_pred, unitvec_pred, heatmap_pred, orient_pred, index_per_point, thres_r=0.2, device="cuda:0"):
    """Joint parameter voting of evaluation/eval_joint_params.py:178-190 for a batch: gocs (B,N,3K or 3) [h5
    'gocs_per_point'], mask_pred (B,N,K) ['instance_per_point'], unitvec_pred (B,N,3), heatmap_pred (B,N,1) or (B,N),
    orient_pred (B,N,3) ['joint_axis_per_point'], index_per_point (B,N,J).  Returns per cloud the reference's
    `joints['pred']` list: [{'l': median axis, 'p': median joint point} for j in 1..K-1] (f32; NaN where a joint has no
    voters), plus the voter counts (B,K-1)."""
    dev = torch.device(device)

    def dv(x):
        t = x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, np.float32))
        return t.to(dev).contiguous()
    g, m, u, h, o, ix = (dv(x) for x in (gocs, mask_pred, unitvec_pred, heatmap_pred, orient_pred, index_per_point))
    B, N, K = m.shape
    if g.shape[:2] != (B, N) or g.shape[2] not in (3, 3 * K) or tuple(u.shape) != (B, N, 3) or tuple(o.shape) != (B, N, 3) or \
            h.numel() != B * N or ix.shape[:2] != (B, N) or any(t.dtype != torch.float32 for t in (g, m, u, h, o, ix)):
        raise ValueError("joint_vote: f32 arrays gocs (B,N,3K|3), mask (B,N,K), unitvec/orient (B,N,3), heatmap (B,N[,1]), index (B,N,J)")
    axis = torch.empty((B, K - 1, 3), dtype=torch.float32, device=dev)
    pt = torch.empty((B, K - 1, 3), dtype=torch.float32, device=dev)
    cnt = torch.empty((B, K - 1), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.ancsh_joint_vote(B, N, K, g.shape[2], ix.shape[2], g.data_ptr(), m.data_ptr(), u.data_ptr(), h.data_ptr(),
                                         o.data_ptr(), ix.data_ptr(), float(thres_r), axis.data_ptr(), pt.data_ptr(),
                                         cnt.data_ptr(), torch.cuda.current_stream().cuda_stream), "ancsh_joint_vote")
    axis, pt = axis.cpu().numpy(), pt.cpu().numpy()
    joints = [[{"l": axis[b, j], "p": pt[b, j]} for j in range(K - 1)] for b in range(B)]
    return joints, cnt.cpu().numpy()
