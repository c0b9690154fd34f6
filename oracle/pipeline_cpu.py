"""oracle/pipeline_cpu.py -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's CPU path for one cloud, end to end, timed by bench.py (cpu_baseline and --impl reference):
network forward(s) via oracle/pnpp.py, pose stage via oracle/pose_np.py.  The reference parallelises the pose
stage over CLOUDS with os.cpu_count()-2 forked workers (evaluation/pose_multi_process.py:54-67); so that a bounded
sample of a few clouds still keeps every core busy, hypotheses are spread over the same number of workers
instead (RANSAC hypotheses are independent; argmax + refit stay serial) -- same arithmetic, same total work.
"""
import multiprocessing as mp
import os

import numpy as np

from . import pnpp, pose_np

_POOL = None


def workers():
    return max(1, (os.cpu_count() or 1) - 2)


def pool():
    global _POOL
    if _POOL is None:
        _POOL = mp.get_context("fork").Pool(workers())
    return _POOL


def close_pool():
    global _POOL
    if _POOL is not None:
        _POOL.terminate()
        _POOL = None


def _single_scores(args):
    src, tgt, th, idx = args
    ds = {"source": src, "target": tgt}
    out = np.zeros(len(idx), np.int64)
    for i, s in enumerate(idx):
        out[i] = pose_np.single_verifier(ds, pose_np.single_estimator(ds, s), th)[0]
    return out


def _joint_scores(args):
    ds, th, i0, i1 = args
    out = np.zeros(len(i0), np.float64)
    for i, (a, b) in enumerate(zip(i0, i1)):
        out[i] = pose_np.joint_verifier(ds, pose_np.joint_estimator(ds, a, b), th)[0]
    return out


def solve_cloud_parallel(P, nocs, mask, joint_axis, joint_cls, K, th, niter_single, niter_joint, rng):
    """pose_np.solve_cloud with the hypothesis loops spread over the worker pool."""
    P = np.asarray(P, np.float64)
    nocs = np.asarray(nocs, np.float64)
    cls = np.argmax(mask, axis=1)
    partidx = [np.where(cls == j)[0] for j in range(K)]
    W = workers()
    res = {"baseline": [], "nonlinear": []}
    for j in range(K):
        if len(partidx[j]) == 0:
            res["baseline"].append(None)
            continue
        src, tgt = nocs[partidx[j], 3 * j:3 * j + 3], P[partidx[j]]
        idx = rng.integers(0, len(partidx[j]), size=(niter_single, 3))
        chunks = np.array_split(idx, W)
        scores = np.concatenate(pool().map(_single_scores, [(src, tgt, th, c) for c in chunks if len(c)]))
        best = int(np.argmax(scores))                                  # first maximum, like ransac()'s strict '>'
        ds = {"source": src, "target": tgt}
        _, inl = pose_np.single_verifier(ds, pose_np.single_estimator(ds, idx[best]), th)
        res["baseline"].append(pose_np.single_estimator(ds, inl) if inl.any() else None)
    for j in range(1, K):
        if len(partidx[0]) == 0 or len(partidx[j]) == 0:
            res["nonlinear"].append(None)
            continue
        jidx = np.where(np.asarray(joint_cls) == j)[0]
        ds = {"source0": nocs[partidx[0], :3], "target0": P[partidx[0]], "source1": nocs[partidx[j], 3 * j:3 * j + 3],
              "target1": P[partidx[j]],
              "joint_direction": np.median(np.asarray(joint_axis, np.float64)[jidx], 0) if len(jidx) else np.zeros(3)}
        i0 = rng.integers(0, len(partidx[0]), size=(niter_joint, 3))
        i1 = rng.integers(0, len(partidx[j]), size=(niter_joint, 3))
        c0, c1 = np.array_split(i0, W), np.array_split(i1, W)
        scores = np.concatenate(pool().map(_joint_scores, [(ds, th, a, b) for a, b in zip(c0, c1) if len(a)]))
        best = int(np.argmax(scores))
        _, inl = pose_np.joint_verifier(ds, pose_np.joint_estimator(ds, i0[best], i1[best]), th)
        ok = inl[0].any() and inl[1].any()
        res["nonlinear"].append(pose_np.joint_estimator(ds, inl[0], inl[1]) if ok else None)
    return res


def _solve_cloud_serial(args):
    P, nocs, mask, axis, jc, K, th, ns, nj, seed = args
    rng = np.random.default_rng(seed)
    cnt = np.bincount(np.argmax(mask, axis=1), minlength=K)
    if (cnt == 0).any():
        return 0                                   # the reference's worker would raise in np.random.randint(0)
    idx_s = [rng.integers(0, cnt[j], size=(ns, 3)) for j in range(K)]
    idx_0 = [rng.integers(0, cnt[0], size=(nj, 3)) for _ in range(1, K)]
    idx_1 = [rng.integers(0, cnt[j], size=(nj, 3)) for j in range(1, K)]
    pose_np.solve_cloud(P, nocs, mask, axis, jc, K, th, idx_s, idx_0, idx_1)
    return 1


def run_clouds_fanout(P, joint_cls, w_ancsh, w_npcs, K, nsample, th, niter_single, niter_joint, seed=0):
    """BASELINE configs[0] as the reference runs it: `main.py --test` over all clouds (both networks; the native ops use all
    cores), then ONE `pose_multi_process.py`: the clouds are cut into contiguous slices, one per forked worker
    (cpu_count - 2 of them, evaluation/pose_multi_process.py:54-67), and every worker solves its clouds serially
    (parallel_ancsh_pose.py:214-352).  Returns #clouds done."""
    pnpp.set_threads(os.cpu_count() or 1)
    B = P.shape[0]
    jobs = []
    for b in range(B):
        pred = pnpp.forward(P[b:b + 1], w_ancsh, K, nsample=nsample)
        src = pred
        if w_npcs is not None:
            src = pnpp.forward(P[b:b + 1], w_npcs, K, nsample=nsample, mixed_pred=False, early_split_nocs=False)
        jobs.append((P[b], src["nocs_per_point"][0], src["W"][0], pred["joint_axis_per_point"][0], joint_cls[b], K, th,
                     niter_single, niter_joint, seed * 100003 + b))
    W = workers()
    num_per = int(B / W) + 1                                       # pose_multi_process.py:55
    slices = [jobs[k * num_per:(k + 1) * num_per] for k in range(W) if jobs[k * num_per:(k + 1) * num_per]]
    pool().map(_run_slice, slices)
    return B


def _run_slice(jobs):
    return sum(_solve_cloud_serial(j) for j in jobs)


def run_clouds(P, joint_cls, w_ancsh, w_npcs, K, nsample, th, niter_single, niter_joint, stages="full", seed=0):
    """Full CPU path for a batch of clouds (sequentially; every stage uses all cores).  Returns #clouds done."""
    rng = np.random.default_rng(seed)
    pnpp.set_threads(os.cpu_count() or 1)
    for b in range(P.shape[0]):
        pred = pnpp.forward(P[b:b + 1], w_ancsh, K, nsample=nsample)
        if stages == "forward":
            continue
        src = pred
        if w_npcs is not None:
            src = pnpp.forward(P[b:b + 1], w_npcs, K, nsample=nsample, mixed_pred=False, early_split_nocs=False)
        solve_cloud_parallel(P[b], src["nocs_per_point"][0], src["W"][0], pred["joint_axis_per_point"][0], joint_cls[b], K,
                             th, niter_single, niter_joint, rng)
    return P.shape[0]
