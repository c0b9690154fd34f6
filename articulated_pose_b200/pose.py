"""Pose stage -- host-side mirror of the reference's RANSAC / joint-solve interfaces over the C ABI.

  PoseSolver.solve(...)        per-cloud body of solver_ransac_nonlinear     (evaluation/parallel_ancsh_pose.py:214-352)
  ransac_single / ransac_joint ransac(dataset, estimator, verifier, inlier_th, niter) with the single / joint
                               estimator+verifier pairs (:20-54, :106-194); same dataset and model dict keys
  rts_dict                     the result-dict schema written to the sub-pickles (:251-256, 346-352)
  compute_gt_pose              evaluation/compute_gt_pose.py:80-97 (Umeyama per part, lib/aligning.py:580-622)

The reference draws hypothesis samples from the unseeded global np.random (:38,110-111).  Here the samples are
either given explicitly (`idx_*`, positions inside each part's point list) or generated on the device by
Philox4x32-10 keyed with `seed` (and retrievable with `sample_indices` for replay into a CPU checker).
All tensors are torch CUDA tensors used as device-memory containers; compute is in libancsh_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_OUT_SPEC = {  # name -> (dtype, shape(B,K,N,J))
    "single_R": (torch.float64, lambda B, K, N, J: (B, K, 3, 3)),
    "single_s": (torch.float64, lambda B, K, N, J: (B, K)),
    "single_t": (torch.float64, lambda B, K, N, J: (B, K, 3)),
    "single_score": (torch.int32, lambda B, K, N, J: (B, K)),
    "single_inliers": (torch.uint8, lambda B, K, N, J: (B, K, N)),
    "joint_R0": (torch.float64, lambda B, K, N, J: (B, J, 3, 3)),
    "joint_s0": (torch.float64, lambda B, K, N, J: (B, J)),
    "joint_t0": (torch.float64, lambda B, K, N, J: (B, J, 3)),
    "joint_R1": (torch.float64, lambda B, K, N, J: (B, J, 3, 3)),
    "joint_s1": (torch.float64, lambda B, K, N, J: (B, J)),
    "joint_t1": (torch.float64, lambda B, K, N, J: (B, J, 3)),
    "joint_score": (torch.float64, lambda B, K, N, J: (B, J)),
    "joint_inliers0": (torch.uint8, lambda B, K, N, J: (B, J, N)),
    "joint_inliers1": (torch.uint8, lambda B, K, N, J: (B, J, N)),
    "part_count": (torch.int32, lambda B, K, N, J: (B, K)),
    "status": (torch.int32, lambda B, K, N, J: (B, K)),
}

_WS_SPEC = {
    "part_idx": (torch.int32, lambda B, K, N, J, ns, nj: (B, K, N)),
    "part_src": (torch.float32, lambda B, K, N, J, ns, nj: (B, K, N, 3)),
    "part_tgt": (torch.float32, lambda B, K, N, J, ns, nj: (B, K, N, 3)),
    "axis_med": (torch.float64, lambda B, K, N, J, ns, nj: (B, J, 3)),
    "single_scores": (torch.int32, lambda B, K, N, J, ns, nj: (B, K, ns)),
    "joint_scores": (torch.float64, lambda B, K, N, J, ns, nj: (B, J, nj)),
    "single_best": (torch.int32, lambda B, K, N, J, ns, nj: (B, K)),
    "joint_best": (torch.int32, lambda B, K, N, J, ns, nj: (B, J)),
    "joint_nfev": (torch.int32, lambda B, K, N, J, ns, nj: (B, J, nj)),
    "joint_models": (torch.float64, lambda B, K, N, J, ns, nj: (B, J, nj, 26)),
}
_ESIZE = {torch.int32: 4, torch.float32: 4, torch.float64: 8, torch.uint8: 1}


class PoseSolver:
    def __init__(self, n_parts, niter_single=10000, niter_joint=200, inlier_th=0.1, seed=0, device="cuda:0"):
        """Defaults are the reference's: 10000 / 200 hypotheses (parallel_ancsh_pose.py:262,288), threshold 0.1
        (pose_multi_process.py:32)."""
        if not torch.cuda.is_available():
            raise RuntimeError("PoseSolver needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.K = int(n_parts)
        self.cfg = _lib.PoseCfg(self.K, int(niter_single), int(niter_joint), float(inlier_th), int(seed))
        self._ws = {}
        self.last = None

    def _plan(self, B, N, ws_slot=0):
        key = (B, N, ws_slot)
        if key not in self._ws:
            lay = _lib.PoseWs()
            _lib.check(_lib.ancsh_pose_plan(ctypes.byref(self.cfg), B, N, ctypes.byref(lay)), "ancsh_pose_plan")
            self._ws[key] = (torch.empty(lay.total_bytes, dtype=torch.uint8, device=self.device), lay)
        return self._ws[key]

    def alloc_outputs(self, B, N):
        J = max(self.K - 1, 1)
        return {k: torch.empty(shp(B, self.K, N, J), dtype=dt, device=self.device) for k, (dt, shp) in _OUT_SPEC.items()}

    def solve_device(self, P, nocs, mask, joint_axis=None, joint_cls=None, idx_single=None, idx_joint0=None,
                     idx_joint1=None, out=None, stage_events=None, ws_slot=0, seed=None):
        """P (B,N,3) f32, nocs (B,N,3K) f32, mask (B,N,K) f32, joint_axis (B,N,3) f32, joint_cls (B,N) int32;
        optional idx_single (B,K,niter_single,3), idx_joint0/1 (B,K-1,niter_joint,3) int32.  Launches on torch's
        current stream and returns a dict of CUDA tensors (see include/ancsh_b200.h: ancsh_pose_out_t).
        seed: Philox key of this call (default: the solver's), so that successive batches draw different hypotheses."""
        B, N, _ = P.shape
        K = self.K

        def chk(t, name, dtype, shape):
            if t is None:
                return None
            if not t.is_cuda or t.dtype != dtype or tuple(t.shape) != tuple(shape):
                raise ValueError("%s must be a CUDA %s tensor of shape %s, got %s %s" % (name, dtype, shape, t.dtype,
                                                                                         tuple(t.shape)))
            return t.contiguous()
        P = chk(P, "P", torch.float32, (B, N, 3))
        nocs = chk(nocs, "nocs", torch.float32, (B, N, 3 * K))
        mask = chk(mask, "mask", torch.float32, (B, N, K))
        joint_axis = chk(joint_axis, "joint_axis", torch.float32, (B, N, 3))
        joint_cls = chk(joint_cls, "joint_cls", torch.int32, (B, N))
        idx_single = chk(idx_single, "idx_single", torch.int32, (B, K, self.cfg.niter_single, 3))
        idx_joint0 = chk(idx_joint0, "idx_joint0", torch.int32, (B, max(K - 1, 0), self.cfg.niter_joint, 3))
        idx_joint1 = chk(idx_joint1, "idx_joint1", torch.int32, (B, max(K - 1, 0), self.cfg.niter_joint, 3))
        if K > 1 and (joint_axis is None or joint_cls is None):
            raise ValueError("joint_axis and joint_cls are required when n_parts > 1")
        ws, lay = self._plan(B, N, ws_slot)       # one workspace per slot: slots may be in flight concurrently
        if out is None:
            out = self.alloc_outputs(B, N)
        pin = _lib.PoseIn()
        for k, v in zip(_lib.POSE_IN_FIELDS, (P, nocs, mask, joint_axis, joint_cls, idx_single, idx_joint0, idx_joint1)):
            setattr(pin, k, v.data_ptr() if v is not None else None)
        pout = _lib.PoseOut()
        for k in _lib.POSE_OUT_FIELDS:
            setattr(pout, k, out[k].data_ptr())
        cfg = self.cfg if seed is None else _lib.PoseCfg(self.K, self.cfg.niter_single, self.cfg.niter_joint,
                                                         self.cfg.inlier_th, int(seed))
        rc = _lib.ancsh_pose_solve(ctypes.byref(cfg), ctypes.byref(pin), B, N, ws.data_ptr(), lay.total_bytes,
                                   ctypes.byref(pout), stage_events.arr if stage_events is not None else None,
                                   torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "ancsh_pose_solve")
        self.last = (ws, lay, B, N)
        return out

    def intermediates(self):
        ws, lay, B, N = self.last
        K, J = self.K, max(self.K - 1, 1)
        res = {}
        for name, (dt, shp) in _WS_SPEC.items():
            shape = shp(B, K, N, J, self.cfg.niter_single, self.cfg.niter_joint)
            nbytes = int(np.prod(shape)) * _ESIZE[dt]
            off = getattr(lay, name)
            res[name] = ws[off:off + nbytes].view(dt).view(shape)
        cnt = ws[lay.joint_tail:lay.joint_tail + 256].view(torch.int32)
        res["lm_njev_total"], res["lm_lmpar_total"] = cnt[60], cnt[61]      # totals over all LM solves of the call
        return res

    def solve(self, P, nocs, mask, joint_axis=None, joint_cls=None, idx_single=None, idx_joint0=None, idx_joint1=None):
        """Host arrays in, list (one per cloud) of {'baseline': [model per part], 'nonlinear': [model per joint],
        'inliers_single', 'inliers_joint', 'part_count', 'status'} out; models use the reference's dict keys
        (parallel_ancsh_pose.py:42-46, 177-184)."""
        def dev(a, dt):
            return None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(self.device)
        with torch.cuda.device(self.device):
            out = self.solve_device(dev(P, np.float32), dev(nocs, np.float32), dev(mask, np.float32),
                                    dev(joint_axis, np.float32), dev(joint_cls, np.int32), dev(idx_single, np.int32),
                                    dev(idx_joint0, np.int32), dev(idx_joint1, np.int32))
            h = {k: v.cpu().numpy() for k, v in out.items()}
        return unpack_results(h, self.K)

    def sample_indices(self, stream_id, n_per_problem, niter):
        """The positions Philox generates for (seed, problem, hypothesis): (nprob,niter,3) int32 on the host.
        stream_id 0: single RANSAC (problem = b*K+j, n = part_count); 1 / 2: joint RANSAC part 0 / part j
        (problem = b*(K-1)+(j-1))."""
        n_per = torch.from_numpy(np.ascontiguousarray(n_per_problem, dtype=np.int32)).to(self.device)
        nprob = int(n_per.numel())
        idx = torch.empty((nprob, niter, 3), dtype=torch.int32, device=self.device)
        _lib.check(_lib.ancsh_pose_sample_indices(self.cfg.seed, int(stream_id), nprob, int(niter), n_per.data_ptr(),
                                                  idx.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "ancsh_pose_sample_indices")
        return idx.cpu().numpy()


def unpack_results(h, K):
    B = h["single_s"].shape[0]
    res = []
    for b in range(B):
        cnt = h["part_count"][b]
        r = {"baseline": [], "nonlinear": [], "inliers_single": [], "inliers_joint": [], "part_count": cnt.copy(),
             "status": h["status"][b].copy(), "single_score": h["single_score"][b].copy()}
        for j in range(K):
            r["baseline"].append({"rotation": h["single_R"][b, j].copy(), "scale": float(h["single_s"][b, j]),
                                  "translation": h["single_t"][b, j].copy()})
            r["inliers_single"].append(h["single_inliers"][b, j, :cnt[j]].astype(bool))
        for j in range(1, K):
            r["nonlinear"].append({
                "rotation0": h["joint_R0"][b, j - 1].copy(), "scale0": float(h["joint_s0"][b, j - 1]),
                "translation0": h["joint_t0"][b, j - 1].copy(),
                "rotation1": h["joint_R1"][b, j - 1].copy(), "scale1": float(h["joint_s1"][b, j - 1]),
                "translation1": h["joint_t1"][b, j - 1].copy(), "score": float(h["joint_score"][b, j - 1])})
            r["inliers_joint"].append([h["joint_inliers0"][b, j - 1, :cnt[0]].astype(bool),
                                       h["joint_inliers1"][b, j - 1, :cnt[j]].astype(bool)])
        res.append(r)
    return res


# ---------------------------------------------------------------------------------------------------------
# ransac()-compatible entry points (evaluation/parallel_ancsh_pose.py:20)
# ---------------------------------------------------------------------------------------------------------
def ransac_single(dataset, inlier_th, niter=10000, sample_idx=None, seed=0, device="cuda:0"):
    """ransac(dataset, single_transformation_estimator, single_transformation_verifier, inlier_th, niter):
    dataset = {'source': (n,3), 'target': (n,3), 'nsource': n} -> (best_model, best_inliers) with
    best_model = {'rotation','scale','translation'} (:42-46).  An empty dataset raises ValueError like
    np.random.randint(0) does in the reference (:38)."""
    src = np.asarray(dataset["source"], np.float32)
    tgt = np.asarray(dataset["target"], np.float32)
    n = src.shape[0]
    if n == 0:
        raise ValueError("ransac_single: empty dataset (the reference raises in np.random.randint)")
    solver = PoseSolver(1, niter_single=niter, niter_joint=1, inlier_th=inlier_th, seed=seed, device=device)
    idx = None if sample_idx is None else np.asarray(sample_idx, np.int32).reshape(1, 1, niter, 3)
    r = solver.solve(tgt[None], src[None], np.ones((1, n, 1), np.float32), idx_single=idx)[0]
    return r["baseline"][0], r["inliers_single"][0]


def ransac_joint(dataset, inlier_th, niter=200, sample_idx0=None, sample_idx1=None, seed=0, device="cuda:0"):
    """ransac(dataset, joint_transformation_estimator, joint_transformation_verifier, inlier_th, niter):
    dataset = {'source0','target0','nsource0','source1','target1','nsource1','joint_direction'} (:297-304) ->
    (best_model, [inliers0, inliers1]), best_model keys rotation0/scale0/translation0/rotation1/... (:177-184)."""
    s0, t0 = np.asarray(dataset["source0"], np.float32), np.asarray(dataset["target0"], np.float32)
    s1, t1 = np.asarray(dataset["source1"], np.float32), np.asarray(dataset["target1"], np.float32)
    n0, n1 = s0.shape[0], s1.shape[0]
    if n0 == 0 or n1 == 0:
        raise ValueError("ransac_joint: empty part (the reference raises in np.random.randint)")
    n = n0 + n1
    P = np.concatenate([t0, t1])[None]
    nocs = np.zeros((1, n, 6), np.float32)
    nocs[0, :n0, 0:3] = s0
    nocs[0, n0:, 3:6] = s1
    mask = np.zeros((1, n, 2), np.float32)
    mask[0, :n0, 0] = 1
    mask[0, n0:, 1] = 1
    axis = np.tile(np.asarray(dataset["joint_direction"], np.float32).reshape(1, 1, 3), (1, n, 1))
    solver = PoseSolver(2, niter_single=1, niter_joint=niter, inlier_th=inlier_th, seed=seed, device=device)
    i0 = None if sample_idx0 is None else np.asarray(sample_idx0, np.int32).reshape(1, 1, niter, 3)
    i1 = None if sample_idx1 is None else np.asarray(sample_idx1, np.int32).reshape(1, 1, niter, 3)
    r = solver.solve(P, nocs, mask, axis, np.ones((1, n), np.int32), idx_joint0=i0, idx_joint1=i1)[0]
    return r["nonlinear"][0], r["inliers_joint"][0]


# ---------------------------------------------------------------------------------------------------------
# result schema of solver_ransac_nonlinear and GT poses
# ---------------------------------------------------------------------------------------------------------
def rot_diff_degree(rot1, rot2):
    """lib/d3_utils.py:144-148"""
    return (np.arccos((np.trace(np.matmul(rot1, rot2.T)) - 1) / 2) % (2 * np.pi)) / np.pi * 180


def rts_dict(result, rt_gt, scale_gt):
    """One cloud's entry of the pickle solver_ransac_nonlinear writes (parallel_ancsh_pose.py:251-256, 270-352):
    {'scale','rotation','translation','xyz_err','rpy_err','scale_err'} each {'gt','baseline','nonlinear'}.
    rt_gt: list of 4x4 per part, scale_gt: list of (3,) per part (evaluation/compute_gt_pose.py:85-90)."""
    K = len(result["baseline"])
    scale_d = {"gt": [], "baseline": [], "nonlinear": []}
    r_d = {"gt": [], "baseline": [], "nonlinear": []}
    t_d = {"gt": [], "baseline": [], "nonlinear": []}
    xyz_err, rpy_err, scale_err = ({"baseline": [], "nonlinear": []} for _ in range(3))
    for j in range(K):
        m = result["baseline"][j]
        rpy_err["baseline"].append(rot_diff_degree(m["rotation"], rt_gt[j][:3, :3]))
        xyz_err["baseline"].append(np.linalg.norm(m["translation"] - rt_gt[j][:3, 3]))
        scale_err["baseline"].append(np.linalg.norm(m["scale"] - scale_gt[j][0]))
        scale_d["baseline"].append(m["scale"]); r_d["baseline"].append(m["rotation"]); t_d["baseline"].append(m["translation"])
    for j in range(1, K):
        m = result["nonlinear"][j - 1]
        parts = ([(0, "0")] if j == 1 else []) + [(j, "1")]
        for pj, sfx in parts:
            rpy_err["nonlinear"].append(rot_diff_degree(m["rotation" + sfx], rt_gt[pj][:3, :3]))
            xyz_err["nonlinear"].append(np.linalg.norm(m["translation" + sfx] - rt_gt[pj][:3, 3]))
            scale_err["nonlinear"].append(np.linalg.norm(m["scale" + sfx] - scale_gt[pj][0]))
            scale_d["gt"].append(scale_gt[pj][0]); scale_d["nonlinear"].append(m["scale" + sfx])
            r_d["gt"].append(rt_gt[pj][:3, :3]); r_d["nonlinear"].append(m["rotation" + sfx])
            t_d["gt"].append(rt_gt[pj][:3, 3]); t_d["nonlinear"].append(m["translation" + sfx])
    return {"scale": scale_d, "rotation": r_d, "translation": t_d, "xyz_err": xyz_err, "rpy_err": rpy_err,
            "scale_err": scale_err}


def umeyama(src, tgt, cnt, device="cuda:0"):
    """Batched estimateSimilarityUmeyama (lib/aligning.py:580-622): src/tgt (nprob,nmax,3), cnt (nprob) ->
    Scales (nprob,3), Rotation (nprob,3,3) [reference convention (U Vh)^T], Translation (nprob,3),
    OutTransform (nprob,4,4)."""
    dev = torch.device(device)
    s = torch.from_numpy(np.ascontiguousarray(src, np.float32)).to(dev)
    t = torch.from_numpy(np.ascontiguousarray(tgt, np.float32)).to(dev)
    c = torch.from_numpy(np.ascontiguousarray(cnt, np.int32)).to(dev)
    nprob, nmax, _ = s.shape
    scale = torch.empty(nprob, dtype=torch.float64, device=dev)
    R = torch.empty((nprob, 3, 3), dtype=torch.float64, device=dev)
    tr = torch.empty((nprob, 3), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.ancsh_umeyama(nprob, nmax, s.data_ptr(), t.data_ptr(), c.data_ptr(), scale.data_ptr(), R.data_ptr(),
                                      tr.data_ptr(), torch.cuda.current_stream().cuda_stream), "ancsh_umeyama")
    scale, R, tr = scale.cpu().numpy(), R.cpu().numpy(), tr.cpu().numpy()
    out = np.tile(np.eye(4), (nprob, 1, 1))
    out[:, :3, :3] = scale[:, None, None] * np.transpose(R, (0, 2, 1))       # diag(Scales) @ Rotation.T  (:616)
    out[:, :3, 3] = tr
    return np.repeat(scale[:, None], 3, axis=1), R, tr, out


def compute_gt_pose(P, nocs_gt, cls_gt, n_parts, device="cuda:0"):
    """evaluation/compute_gt_pose.py:80-97 for one cloud: per part Umeyama(nocs_gt[part], P[part]) ->
    {'scale': {'gt': [...]}, 'rt': {'gt': [4x4 ...]}} with compose_rt's transposed rotation (:14-19)."""
    N = P.shape[0]
    src = np.zeros((n_parts, N, 3), np.float32)
    tgt = np.zeros((n_parts, N, 3), np.float32)
    cnt = np.zeros(n_parts, np.int32)
    for j in range(n_parts):
        m = np.where(cls_gt == j)[0]
        cnt[j] = len(m)
        src[j, :len(m)] = nocs_gt[m]
        tgt[j, :len(m)] = P[m]
    s, r, t, _ = umeyama(src, tgt, cnt, device)
    rts = []
    for j in range(n_parts):
        rt = np.zeros((4, 4), np.float32)
        rt[:3, :3] = r[j].T                                                  # compose_rt stores rotation.T
        rt[:3, 3] = t[j]
        rt[3, 3] = 1
        rts.append(rt)
    return {"scale": {"gt": [s[j] for j in range(n_parts)]}, "rt": {"gt": rts}}


def similarity_ransac(src, tgt, cnt, sample_idx, device="cuda:0"):
    """Batched estimateSimilarityTransform (lib/aligning.py:17-33) through ancsh_similarity_ransac:
    src/tgt (nprob,nmax,3), cnt (nprob), sample_idx (nprob,niter,5) int (the reference's np.random.randint(n, size=5)
    draws, each < cnt[p]).  Returns a dict of arrays: scale (nprob), rotation (nprob,3,3) [(U Vh)^T], translation (nprob,3),
    inlier_ratio, inliers (nprob,nmax) bool, iters (nprob), status (nprob)."""
    dev = torch.device(device)
    s = torch.from_numpy(np.ascontiguousarray(src, np.float32)).to(dev)
    t = torch.from_numpy(np.ascontiguousarray(tgt, np.float32)).to(dev)
    c_host = np.ascontiguousarray(cnt, np.int32)
    idx_host = np.ascontiguousarray(sample_idx, np.int32)
    nprob, nmax, _ = s.shape
    if t.shape != s.shape or c_host.shape != (nprob,) or idx_host.ndim != 3 or idx_host.shape[0] != nprob or idx_host.shape[2] != 5:
        raise ValueError("similarity_ransac: src/tgt (nprob,nmax,3), cnt (nprob), sample_idx (nprob,niter,5)")
    if (idx_host < 0).any() or (idx_host >= np.maximum(c_host, 1)[:, None, None]).any():
        raise ValueError("similarity_ransac: sample indices must lie in [0, cnt)")
    niter = idx_host.shape[1]
    c = torch.from_numpy(c_host).to(dev)
    idx = torch.from_numpy(idx_host).to(dev)
    scale = torch.empty(nprob, dtype=torch.float64, device=dev)
    R = torch.empty((nprob, 3, 3), dtype=torch.float64, device=dev)
    tr = torch.empty((nprob, 3), dtype=torch.float64, device=dev)
    ratio = torch.empty(nprob, dtype=torch.float64, device=dev)
    inl = torch.empty((nprob, nmax), dtype=torch.uint8, device=dev)
    iters = torch.empty(nprob, dtype=torch.int32, device=dev)
    status = torch.empty(nprob, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.ancsh_similarity_ransac(nprob, nmax, niter, s.data_ptr(), t.data_ptr(), c.data_ptr(), idx.data_ptr(),
                                                scale.data_ptr(), R.data_ptr(), tr.data_ptr(), ratio.data_ptr(), inl.data_ptr(),
                                                iters.data_ptr(), status.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "ancsh_similarity_ransac")
    return {"scale": scale.cpu().numpy(), "rotation": R.cpu().numpy(), "translation": tr.cpu().numpy(),
            "inlier_ratio": ratio.cpu().numpy(), "inliers": inl.cpu().numpy().astype(bool), "iters": iters.cpu().numpy(),
            "status": status.cpu().numpy()}


def estimateSimilarityTransform(source, target, rt_pre=None, verbose=False, sample_idx=None, seed=0, device="cuda:0"):
    """Drop-in for lib/aligning.py:17-33 estimateSimilarityTransform(source, target): (n,3) arrays ->
    (Scales (3,), Rotation (3,3), Translation (3,), OutTransform (4,4)), or four Nones when the best inlier ratio is below
    0.1.  `sample_idx` (100,5) replays given draws; otherwise they come from a seeded generator (the reference uses the
    unseeded global np.random).  rt_pre (an externally supplied rotation, unused by every caller in the reference) is not
    supported."""
    if rt_pre is not None:
        raise NotImplementedError("rt_pre is not supported")
    source, target = np.asarray(source), np.asarray(target)
    n = source.shape[0]
    if sample_idx is None:
        sample_idx = np.random.default_rng(seed).integers(0, n, size=(100, 5))
    r = similarity_ransac(source[None], target[None], np.array([n]), np.asarray(sample_idx)[None], device)
    if r["status"][0] != 0:
        if verbose or r["inlier_ratio"][0] < 0.1:
            print('[ WARN ] - Something is wrong. Small BestInlierRatio: ', r["inlier_ratio"][0])
        return None, None, None, None
    sf, Rot, tr = r["scale"][0], r["rotation"][0], r["translation"][0]
    out = np.identity(4)
    out[:3, :3] = sf * Rot.T                                                 # diag(Scales) @ Rotation.T (:616)
    out[:3, 3] = tr
    return np.array([sf, sf, sf]), Rot, tr, out
