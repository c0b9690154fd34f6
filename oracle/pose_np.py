"""oracle/pose_np.py -- TEST INFRASTRUCTURE ONLY.

NumPy/SciPy restatement of the reference's pose stage, with the random sample indices INJECTED (the reference
draws them from the unseeded global np.random, evaluation/parallel_ancsh_pose.py:38,110-111):

  rotate_pts / scale_pts / transform_pts / rotate_points_with_rotvec     lib/d3_utils.py:206-246, 150-163
  ransac + single_transformation_estimator/_verifier                     evaluation/parallel_ancsh_pose.py:20-54
  objective_eval                                                         evaluation/parallel_ancsh_pose.py:56-68
  joint_transformation_estimator/_verifier                               evaluation/parallel_ancsh_pose.py:106-194
  per-cloud body of solver_ransac_nonlinear                              evaluation/parallel_ancsh_pose.py:214-352
  estimateSimilarityUmeyama                                              lib/aligning.py:580-622

Third-party arithmetic is the reference's own: numpy.linalg.svd (LAPACK), scipy.optimize.least_squares
(method='lm' -> MINPACK lmder, x_scale=1.0 as in the pinned scipy 1.3.1) and scipy Rotation.

dtype contract: all inputs are promoted to float64 first.  (The reference slices float32 h5 datasets, so parts
of ITS arithmetic run in float32 depending on how the h5 was written; parity is defined on float64-typed
inputs holding the f32 values -- see DESIGN.md.)

Pinned against the imported reference by tests/test_pose_oracle_cpu.py (when /root/reference is mounted) and by
tests/golden/pose_ref.npz (always).
"""
import numpy as np
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation as srot


# ---------------------------------------------------------------- lib/d3_utils.py
def rotate_pts(source, target):
    """d3_utils.py:206-220"""
    source = source - np.mean(source, 0, keepdims=True)
    target = target - np.mean(target, 0, keepdims=True)
    M = np.matmul(target.T, source)
    U, D, Vh = np.linalg.svd(M, full_matrices=True)
    d = (np.linalg.det(U) * np.linalg.det(Vh)) < 0.0
    if d:
        D[-1] = -D[-1]
        U[:, -1] = -U[:, -1]
    return np.matmul(U, Vh)


def scale_pts(source, target):
    """d3_utils.py:237-246 (all n^2 ordered pairs)"""
    pdist_s = source.reshape(source.shape[0], 1, 3) - source.reshape(1, source.shape[0], 3)
    A = np.sqrt(np.sum(pdist_s ** 2, 2)).reshape(-1)
    pdist_t = target.reshape(target.shape[0], 1, 3) - target.reshape(1, target.shape[0], 3)
    b = np.sqrt(np.sum(pdist_t ** 2, 2)).reshape(-1)
    return np.dot(A, b) / (np.dot(A, A) + 1e-6)


def transform_pts(source, target):
    """d3_utils.py:223-234"""
    source_centered = source - np.mean(source, 0, keepdims=True)
    target_centered = target - np.mean(target, 0, keepdims=True)
    rotation = rotate_pts(source_centered, target_centered)
    scale = scale_pts(source_centered, target_centered)
    translation = np.mean(target.T - scale * np.matmul(rotation, source.T), 1)
    return rotation, scale, translation


def rotate_points_with_rotvec(points, rot_vecs):
    """d3_utils.py:150-163"""
    theta = np.linalg.norm(rot_vecs, axis=1)[:, np.newaxis]
    with np.errstate(invalid="ignore", divide="ignore"):
        v = rot_vecs / theta
        v = np.nan_to_num(v)
    dot = np.sum(points * v, axis=1)[:, np.newaxis]
    cos_theta, sin_theta = np.cos(theta), np.sin(theta)
    return cos_theta * points + sin_theta * np.cross(v, points) + dot * (1 - cos_theta) * v


# ---------------------------------------------------------------- single-part RANSAC
def single_estimator(dataset, sample_idx):
    """parallel_ancsh_pose.py:35-46; sample_idx: (3,) ints or a boolean inlier mask"""
    rotation, scale, translation = transform_pts(dataset["source"][sample_idx, :], dataset["target"][sample_idx, :])
    return {"rotation": rotation, "scale": scale, "translation": translation}


def single_verifier(dataset, model, inlier_th):
    """parallel_ancsh_pose.py:48-54"""
    res = dataset["target"].T - model["scale"] * np.matmul(model["rotation"], dataset["source"].T) \
        - model["translation"].reshape((3, 1))
    inliers = np.sqrt(np.sum(res ** 2, 0)) < inlier_th
    return np.sum(inliers), inliers


def ransac_single(source, target, inlier_th, sample_idx, return_scores=False):
    """parallel_ancsh_pose.py:20-33 with the estimator/verifier above; sample_idx (niter,3)."""
    ds = {"source": np.asarray(source, np.float64), "target": np.asarray(target, np.float64)}
    best_model, best_score, best_inliers = None, -np.inf, None
    scores = np.zeros(len(sample_idx), np.int64)
    for i, idx in enumerate(np.asarray(sample_idx)):
        cur = single_estimator(ds, idx)
        score, inl = single_verifier(ds, cur, inlier_th)
        scores[i] = score
        if score > best_score:
            best_model, best_inliers, best_score = cur, inl, score
    best_model = single_estimator(ds, best_inliers)
    if return_scores:
        return best_model, best_inliers, scores
    return best_model, best_inliers


# ---------------------------------------------------------------- joint RANSAC
def objective_eval(params, x0, y0, x1, y1, joints, isweight=True):
    """parallel_ancsh_pose.py:56-68"""
    rotvec0 = params[:3].reshape((1, 3))
    rotvec1 = params[3:].reshape((1, 3))
    res0 = y0 - rotate_points_with_rotvec(x0, rotvec0)
    res1 = y1 - rotate_points_with_rotvec(x1, rotvec1)
    res_joint = rotate_points_with_rotvec(joints, rotvec0) - rotate_points_with_rotvec(joints, rotvec1)
    if isweight:
        res0 /= x0.shape[0]
        res1 /= x1.shape[0]
        res_joint /= joints.shape[0]
    return np.concatenate((res0, res1, res_joint), 0).ravel()


def joint_estimator(dataset, sample_idx0, sample_idx1, return_info=False):
    """parallel_ancsh_pose.py:106-184 (joint_type='revolute' -- ransac() never passes another, :26)"""
    source0 = dataset["source0"][sample_idx0, :]
    target0 = dataset["target0"][sample_idx0, :]
    source1 = dataset["source1"][sample_idx1, :]
    target1 = dataset["target1"][sample_idx1, :]
    scale0 = scale_pts(source0, target0)
    scale1 = scale_pts(source1, target1)
    scale0_inv = scale_pts(target0, source0)
    scale1_inv = scale_pts(target1, source1)
    target0_sc = scale0_inv * target0
    target0_sc -= np.mean(target0_sc, 0, keepdims=True)
    source0_c = source0 - np.mean(source0, 0, keepdims=True)
    target1_sc = scale1_inv * target1
    target1_sc -= np.mean(target1_sc, 0, keepdims=True)
    source1_c = source1 - np.mean(source1, 0, keepdims=True)
    nj = min(source0.shape[0], source1.shape[0])
    joint_points0 = np.ones((nj, 1)) * dataset["joint_direction"].reshape((1, 3))           # :134
    R0 = rotate_pts(source0_c, target0_sc)
    R1 = rotate_pts(source1_c, target1_sc)
    rotvec0 = srot.from_matrix(R0).as_rotvec()
    rotvec1 = srot.from_matrix(R1).as_rotvec()
    res = least_squares(objective_eval, np.hstack((rotvec0, rotvec1)), verbose=0, ftol=1e-4, method="lm", x_scale=1.0,
                        args=(source0_c, target0_sc, source1_c, target1_sc, joint_points0, False))
    R0 = srot.from_rotvec(res.x[:3]).as_matrix()
    R1 = srot.from_rotvec(res.x[3:]).as_matrix()
    translation0 = np.mean(target0.T - scale0 * np.matmul(R0, source0.T), 1)
    translation1 = np.mean(target1.T - scale1 * np.matmul(R1, source1.T), 1)
    out = {"rotation0": R0, "scale0": scale0, "translation0": translation0,
           "rotation1": R1, "scale1": scale1, "translation1": translation1}
    if return_info:
        out["_x0"] = np.hstack((rotvec0, rotvec1))
        out["_x"] = res.x.copy()
        out["_nfev"] = res.nfev
        out["_status"] = res.status
        out["_cost"] = res.cost
    return out


def joint_verifier(dataset, model, inlier_th):
    """parallel_ancsh_pose.py:186-194 (res.shape[0] is 3, not n)"""
    res0 = dataset["target0"].T - model["scale0"] * np.matmul(model["rotation0"], dataset["source0"].T) \
        - model["translation0"].reshape((3, 1))
    inliers0 = np.sqrt(np.sum(res0 ** 2, 0)) < inlier_th
    res1 = dataset["target1"].T - model["scale1"] * np.matmul(model["rotation1"], dataset["source1"].T) \
        - model["translation1"].reshape((3, 1))
    inliers1 = np.sqrt(np.sum(res1 ** 2, 0)) < inlier_th
    score = (np.sum(inliers0) / res0.shape[0] + np.sum(inliers1) / res1.shape[0]) / 2
    return score, [inliers0, inliers1]


def ransac_joint(source0, target0, source1, target1, joint_direction, inlier_th, sample_idx0, sample_idx1,
                 return_scores=False):
    ds = {"source0": np.asarray(source0, np.float64), "target0": np.asarray(target0, np.float64),
          "source1": np.asarray(source1, np.float64), "target1": np.asarray(target1, np.float64),
          "joint_direction": np.asarray(joint_direction, np.float64)}
    best_model, best_score, best_inliers = None, -np.inf, None
    scores = np.zeros(len(sample_idx0), np.float64)
    for i, (i0, i1) in enumerate(zip(np.asarray(sample_idx0), np.asarray(sample_idx1))):
        cur = joint_estimator(ds, i0, i1)
        score, inl = joint_verifier(ds, cur, inlier_th)
        scores[i] = score
        if score > best_score:
            best_model, best_inliers, best_score = cur, inl, score
    best_model = joint_estimator(ds, best_inliers[0], best_inliers[1], return_info=True)
    if return_scores:
        return best_model, best_inliers, scores
    return best_model, best_inliers


# ---------------------------------------------------------------- per-cloud body of solver_ransac_nonlinear
def solve_cloud(P, nocs_pred, mask_pred, joint_axis_pred, joint_cls_gt, num_parts, inlier_th, idx_single, idx_joint0,
                idx_joint1):
    """parallel_ancsh_pose.py:237-344 without the h5 / GT-error bookkeeping.

    idx_single[j]: (niter,3) samples for part j; idx_joint0[j-1], idx_joint1[j-1]: (njiter,3) for joint j.
    Sample values are positions inside the part's point list (np.random.randint(nsource), :38,110-111).
    Returns {'baseline': [model per part], 'nonlinear': [joint model per joint], 'partidx': [...]}.
    """
    P = np.asarray(P, np.float64)
    nocs_pred = np.asarray(nocs_pred, np.float64)
    cls = np.argmax(mask_pred, axis=1)                                                     # :238
    partidx = [np.where(cls == j)[0] for j in range(num_parts)]
    out = {"baseline": [], "nonlinear": [], "partidx": partidx, "inliers_single": [], "inliers_joint": []}
    for j in range(num_parts):
        src = nocs_pred[partidx[j], 3 * j:3 * (j + 1)]                                     # :259
        tgt = P[partidx[j], :3]
        m, inl = ransac_single(src, tgt, inlier_th, idx_single[j])
        out["baseline"].append(m)
        out["inliers_single"].append(inl)
    for j in range(1, num_parts):
        src0, tgt0 = nocs_pred[partidx[0], :3], P[partidx[0], :3]                          # :291-294
        src1, tgt1 = nocs_pred[partidx[j], 3 * j:3 * (j + 1)], P[partidx[j], :3]
        jidx = np.where(np.asarray(joint_cls_gt) == j)[0]
        jt_axis = np.median(np.asarray(joint_axis_pred, np.float64)[jidx, :], 0)           # :295
        m, inl = ransac_joint(src0, tgt0, src1, tgt1, jt_axis, inlier_th, idx_joint0[j - 1], idx_joint1[j - 1])
        out["nonlinear"].append(m)
        out["inliers_joint"].append(inl)
    return out


# ---------------------------------------------------------------- lib/aligning.py:580-622
def estimate_similarity_umeyama(SourceHom, TargetHom):
    SourceCentroid = np.mean(SourceHom[:3, :], axis=1)
    TargetCentroid = np.mean(TargetHom[:3, :], axis=1)
    nPoints = SourceHom.shape[1]
    CenteredSource = SourceHom[:3, :] - np.tile(SourceCentroid, (nPoints, 1)).transpose()
    CenteredTarget = TargetHom[:3, :] - np.tile(TargetCentroid, (nPoints, 1)).transpose()
    CovMatrix = np.matmul(CenteredTarget, np.transpose(CenteredSource)) / nPoints
    U, D, Vh = np.linalg.svd(CovMatrix, full_matrices=True)
    d = (np.linalg.det(U) * np.linalg.det(Vh)) < 0.0
    if d:
        D[-1] = -D[-1]
        U[:, -1] = -U[:, -1]
    Rotation = np.matmul(U, Vh).T
    varP = np.var(SourceHom[:3, :], axis=1).sum()
    ScaleFact = 1 / varP * np.sum(D)
    Scales = np.array([ScaleFact, ScaleFact, ScaleFact])
    Translation = TargetHom[:3, :].mean(axis=1) - SourceHom[:3, :].mean(axis=1).dot(ScaleFact * Rotation)
    OutTransform = np.identity(4)
    OutTransform[:3, :3] = np.diag(Scales) @ Rotation.T
    OutTransform[:3, 3] = Translation
    return Scales, Rotation, Translation, OutTransform


# ---------------------------------------------------------------- lib/aligning.py:17-33, 88-103, 485-507, 540-547
# estimateSimilarityTransform: the NOCS-style per-part baseline (5-point Umeyama RANSAC, <= 100 iterations, thresholds from
# the source/target norm ratio, Umeyama refit on the best hypothesis' inliers).  SURVEY 8(a) row a-21.  `sample_idx`
# (MaxIterations, 5) replaces the reference's unseeded np.random.randint(n, size=5) draws.
def set_config(source, target):
    SourceHom = np.transpose(np.hstack([source, np.ones([source.shape[0], 1])]))
    TargetHom = np.transpose(np.hstack([target, np.ones([target.shape[0], 1])]))
    TargetNorm = np.mean(np.linalg.norm(target, axis=1))
    SourceNorm = np.mean(np.linalg.norm(source, axis=1))
    RatioTS = TargetNorm / SourceNorm
    RatioST = SourceNorm / TargetNorm
    PassT = RatioST if RatioST > RatioTS else RatioTS
    return SourceHom, TargetHom, PassT, PassT / 100


def evaluate_model(OutTransform, SourceHom, TargetHom, PassThreshold):
    Diff = TargetHom - np.matmul(OutTransform, SourceHom)
    ResidualVec = np.linalg.norm(Diff[:3, :], axis=0)
    Residual = np.linalg.norm(ResidualVec)
    InlierIdx = np.where(ResidualVec < PassThreshold)
    nInliers = np.count_nonzero(InlierIdx)          # counts the non-zero INDICES (:545): point 0 is never counted
    return Residual, nInliers / SourceHom.shape[1], InlierIdx[0]


def get_ransac_inliers(SourceHom, TargetHom, sample_idx, PassThreshold, StopThreshold):
    BestResidual, BestInlierRatio = 1e10, 0
    BestInlierIdx = np.arange(SourceHom.shape[1])
    iters = 0
    for RandIdx in np.asarray(sample_idx):
        iters += 1
        _, _, _, OutTransform = estimate_similarity_umeyama(SourceHom[:, RandIdx], TargetHom[:, RandIdx])
        Residual, InlierRatio, InlierIdx = evaluate_model(OutTransform, SourceHom, TargetHom, PassThreshold)
        if InlierRatio > BestInlierRatio:
            BestResidual, BestInlierRatio, BestInlierIdx = Residual, InlierRatio, InlierIdx
        if BestResidual < StopThreshold:
            break
    return BestInlierIdx, BestInlierRatio, iters


def estimate_similarity_transform(source, target, sample_idx, return_info=False):
    """(Scales, Rotation, Translation, OutTransform) or four Nones when the best inlier ratio is below 0.1 (:25-27)."""
    source, target = np.asarray(source, np.float64), np.asarray(target, np.float64)
    SourceHom, TargetHom, PassT, StopT = set_config(source, target)
    idx, ratio, iters = get_ransac_inliers(SourceHom, TargetHom, sample_idx, PassT, StopT)
    info = {"inlier_idx": idx, "inlier_ratio": ratio, "iters": iters, "pass_t": PassT}
    if ratio < 0.1:
        out = (None, None, None, None)
    else:
        out = estimate_similarity_umeyama(SourceHom[:, idx], TargetHom[:, idx])
    return out + (info,) if return_info else out
