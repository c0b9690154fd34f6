"""Host-side mirrors of the reference's seven-op Python surface (same names, argument order and shapes):

  pointnet_plusplus/utils/tf_ops/sampling/tf_sampling.py:29,48      gather_point, farthest_point_sample
  pointnet_plusplus/utils/tf_ops/grouping/tf_grouping.py:8,33       query_ball_point, group_point
  pointnet_plusplus/utils/tf_ops/3d_interpolation/tf_interpolate.py:8,19   three_nn, three_interpolate

Inputs/outputs are CUDA torch tensors (device-memory containers only); each call launches on torch's
current stream through the C ABI.  Shape errors raise ValueError, like the reference's OP_REQUIRES ->
InvalidArgumentError (tf_sampling.cpp:105, tf_grouping.cpp:79-84, tf_interpolate.cpp:170-176).
"""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, name, dtype, last=None, rank=3):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise ValueError("%s must be %s" % (name, dtype))
    if t.dim() != rank or (last is not None and t.shape[-1] != last):
        raise ValueError("%s has wrong shape %s" % (name, tuple(t.shape)))
    return t.contiguous()


def farthest_point_sample(npoint, inp):
    """inp (B,N,3) f32 -> (B,npoint) int32 (tf_sampling.py:48-56)."""
    inp = _chk(inp, "inp", torch.float32, 3)
    b, n, _ = inp.shape
    out = torch.empty((b, npoint), dtype=torch.int32, device=inp.device)
    _lib.check(_lib.ancsh_fps(b, n, npoint, inp.data_ptr(), None, out.data_ptr(), _stream()), "ancsh_fps")
    return out


def farthest_point_sample_two_level(npoint1, npoint2, inp):
    """Both sampling levels of the trunk in one launch (pointnet_util.py:47 for layer1 and layer2,
    architectures.py:62-70): returns (idx1 (B,npoint1), xyz1 (B,npoint1,3), idx2 (B,npoint2), xyz2 (B,npoint2,3)) with
    idx2 = farthest_point_sample(npoint2, xyz1).  Bit-exact with two launches; npoint1 <= 512."""
    inp = _chk(inp, "inp", torch.float32, 3)
    b, n, _ = inp.shape
    if npoint2 > npoint1:
        raise ValueError("npoint2 must not exceed npoint1")
    idx1 = torch.empty((b, npoint1), dtype=torch.int32, device=inp.device)
    xyz1 = torch.empty((b, npoint1, 3), dtype=torch.float32, device=inp.device)
    idx2 = torch.empty((b, npoint2), dtype=torch.int32, device=inp.device)
    xyz2 = torch.empty((b, npoint2, 3), dtype=torch.float32, device=inp.device)
    _lib.check(_lib.ancsh_fps_two_level(b, n, npoint1, npoint2, inp.data_ptr(), idx1.data_ptr(), xyz1.data_ptr(),
                                        idx2.data_ptr(), xyz2.data_ptr(), _stream()), "ancsh_fps_two_level")
    return idx1, xyz1, idx2, xyz2


def gather_point(inp, idx):
    """inp (B,N,3), idx (B,M) int32 -> (B,M,3) (tf_sampling.py:29-37)."""
    inp = _chk(inp, "inp", torch.float32, 3)
    idx = _chk(idx, "idx", torch.int32, rank=2)
    b, n, _ = inp.shape
    if idx.shape[0] != b:
        raise ValueError("GatherPoint expects idx batch to match inp")
    m = idx.shape[1]
    out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
    _lib.check(_lib.ancsh_gather_point(b, n, m, inp.data_ptr(), idx.data_ptr(), out.data_ptr(), _stream()),
               "ancsh_gather_point")
    return out


def query_ball_point(radius, nsample, xyz1, xyz2):
    """xyz1 (B,N,3) dataset, xyz2 (B,M,3) centroids -> idx (B,M,nsample) int32, pts_cnt (B,M) int32
    (tf_grouping.py:8-20)."""
    xyz1 = _chk(xyz1, "xyz1", torch.float32, 3)
    xyz2 = _chk(xyz2, "xyz2", torch.float32, 3)
    b, n, _ = xyz1.shape
    if xyz2.shape[0] != b:
        raise ValueError("QueryBallPoint expects matching batch sizes")
    m = xyz2.shape[1]
    idx = torch.empty((b, m, nsample), dtype=torch.int32, device=xyz1.device)
    cnt = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
    _lib.check(_lib.ancsh_ball_query(b, n, m, float(radius), int(nsample), xyz1.data_ptr(), xyz2.data_ptr(),
                                     idx.data_ptr(), cnt.data_ptr(), _stream()), "ancsh_ball_query")
    return idx, cnt


def group_point(points, idx):
    """points (B,N,C), idx (B,M,S) int32 -> (B,M,S,C) (tf_grouping.py:33-41)."""
    points = _chk(points, "points", torch.float32)
    idx = _chk(idx, "idx", torch.int32)
    b, n, c = points.shape
    if idx.shape[0] != b:
        raise ValueError("GroupPoint expects matching batch sizes")
    _, m, s = idx.shape
    out = torch.empty((b, m, s, c), dtype=torch.float32, device=points.device)
    _lib.check(_lib.ancsh_group_point(b, n, c, m, s, points.data_ptr(), idx.data_ptr(), out.data_ptr(), _stream()),
               "ancsh_group_point")
    return out


def three_nn(xyz1, xyz2):
    """xyz1 (B,N,3) queries, xyz2 (B,M,3) known -> dist (B,N,3) squared f32, idx (B,N,3) int32
    (tf_interpolate.py:8-17)."""
    xyz1 = _chk(xyz1, "xyz1", torch.float32, 3)
    xyz2 = _chk(xyz2, "xyz2", torch.float32, 3)
    b, n, _ = xyz1.shape
    if xyz2.shape[0] != b:
        raise ValueError("ThreeNN expects matching batch sizes")
    m = xyz2.shape[1]
    dist = torch.empty((b, n, 3), dtype=torch.float32, device=xyz1.device)
    idx = torch.empty((b, n, 3), dtype=torch.int32, device=xyz1.device)
    _lib.check(_lib.ancsh_three_nn(b, n, m, xyz1.data_ptr(), xyz2.data_ptr(), dist.data_ptr(), idx.data_ptr(),
                                   _stream()), "ancsh_three_nn")
    return dist, idx


def three_interpolate(points, idx, weight):
    """points (B,M,C), idx (B,N,3) int32, weight (B,N,3) -> (B,N,C) (tf_interpolate.py:19-28)."""
    points = _chk(points, "points", torch.float32)
    idx = _chk(idx, "idx", torch.int32, 3)
    weight = _chk(weight, "weight", torch.float32, 3)
    b, m, c = points.shape
    n = idx.shape[1]
    if idx.shape[0] != b or weight.shape != idx.shape:
        raise ValueError("ThreeInterpolate expects (b,n,3) idx and weight")
    out = torch.empty((b, n, c), dtype=torch.float32, device=points.device)
    _lib.check(_lib.ancsh_three_interpolate(b, m, c, n, points.data_ptr(), idx.data_ptr(), weight.data_ptr(),
                                            out.data_ptr(), _stream()), "ancsh_three_interpolate")
    return out
