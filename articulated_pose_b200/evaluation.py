"""Pose-level drop-in: `solver_ransac_nonlinear` and the `pose_multi_process.py` fan-out on the GPU.

Reference: evaluation/parallel_ancsh_pose.py:196-370 (per-cloud loop over the prediction files, result pickle) and
evaluation/pose_multi_process.py:52-68 (contiguous slices of the test group, one sub-pickle per worker).  Here a
"worker" is a GPU rank and the per-cloud loop becomes batched `PoseSolver.solve` calls; names, argument order, the
pickle schema and the sub-pickle file names are the reference's, so `eval_pose_err.py:73-77,125-126` reads the output
unchanged.

Differences that are visible to a caller (all documented in DESIGN.md):
  * hypotheses are drawn by the device Philox generator (`seed`), not from the unseeded global numpy state;
  * clouds with an empty part get NaN models and error entries (the reference raises ValueError inside
    np.random.randint(0) and the worker dies);
  * prediction files are read through prediction_io.load_prediction (h5 when h5py is importable, else npz / in-memory).
"""
import os
import pickle

import numpy as np

from . import prediction_io
from .pose import PoseSolver, rts_dict as _rts_entry


def worker_slices(n_items, n_workers):
    """[(s_ind, e_ind)] of pose_multi_process.py:55-62: num_per_cpu = int(n / cpuCount) + 1, contiguous, clipped."""
    num_per = int(n_items / n_workers) + 1
    return [(min(num_per * k, n_items), min(num_per * (k + 1), n_items)) for k in range(n_workers)]


def result_file_name(directory, baseline_exp, domain, nocs, item, choose_threshold, k=None):
    """pose_multi_process.py:48 (merged file) and :60 (sub file of worker k, inside `<directory>/subs`)."""
    stem = "{}_{}_{}_{}_rt_ours_{}".format(baseline_exp, domain, nocs, item, choose_threshold)
    if k is None:
        return os.path.join(directory, stem + ".pkl")
    return os.path.join(directory, "subs", stem + "_{}.pkl".format(k))


def solver_ransac_nonlinear(s_ind, e_ind, test_exp, baseline_exp, choose_threshold, num_parts, test_group, problem_ins,
                            rts_all, file_name, pred_root=None, use_baseline=True, niter_single=10000, niter_joint=200,
                            batch=256, seed=0, device="cuda:0"):
    """Same positional arguments as the reference (parallel_ancsh_pose.py:196).  `test_exp` / `baseline_exp` are
    directory names under `pred_root` (reference: <base_path>/results/test_pred) or PredictionStore objects.
    Writes {basename: rts_dict} to `file_name` (None: no file) and returns it."""
    def source(exp):
        return exp if isinstance(exp, prediction_io.PredictionStore) else os.path.join(pred_root or ".", str(exp))

    names = []
    for i in range(s_ind, e_ind):
        if test_group[i].split("_")[0] in problem_ins:                     # :217-219
            continue
        names.append(test_group[i].split(".")[0])
    all_rts = {}
    solver = None
    for b0 in range(0, len(names), batch):
        chunk = names[b0:b0 + batch]
        P, nocs, mask, axis, jcls = [], [], [], [], []
        for base in chunk:
            f = prediction_io.load_prediction(source(test_exp), base)
            fb = prediction_io.load_prediction(source(baseline_exp), base) if use_baseline else f
            m = np.asarray(fb["instance_per_point"][()])                   # :227, :232-236 (USE_BASELINE)
            if m.ndim == 1:                                                # W_reduced files hold labels, not scores
                m = np.eye(num_parts, dtype=np.float32)[m.astype(np.int64)]
            P.append(np.asarray(f["P"][()], np.float32)[:, :3])
            nocs.append(np.asarray(fb["nocs_per_point"][()], np.float32))
            mask.append(m.astype(np.float32))
            axis.append(np.asarray(f["joint_axis_per_point"][()], np.float32))
            jcls.append(np.asarray(f["joint_cls_gt"][()]).astype(np.int32))
            f.close()
            if fb is not f:
                fb.close()
        if solver is None:
            solver = PoseSolver(num_parts, niter_single=niter_single, niter_joint=niter_joint,
                                inlier_th=choose_threshold, seed=seed, device=device)
        results = solver.solve(np.stack(P), np.stack(nocs), np.stack(mask), np.stack(axis), np.stack(jcls))
        for base, res in zip(chunk, results):
            gt = rts_all[base]                                             # read only: the reference builds a fresh rts_dict (:346-352)
            all_rts[base] = _rts_entry(res, gt["rt"]["gt"], gt["scale"]["gt"])
    if file_name is not None:
        os.makedirs(os.path.dirname(os.path.abspath(file_name)), exist_ok=True)
        with open(file_name, "wb") as fh:
            pickle.dump(all_rts, fh)
    return all_rts


def pose_multi_process(test_exp, baseline_exp, num_parts, test_group, rts_all, directory, domain="unseen", nocs="ANCSH",
                       item="eyeglasses", choose_threshold=0.1, problem_ins=(), rank=0, world=1, **solver_kw):
    """pose_multi_process.py:52-68 with GPU ranks as the workers: rank k solves slice k of the test group and writes
    `<directory>/subs/<baseline_exp>_<domain>_<nocs>_<item>_rt_ours_<thr>_<k>.pkl`.  Returns (file name, dict)."""
    s_ind, e_ind = worker_slices(len(test_group), world)[rank]
    sub = result_file_name(directory, baseline_exp if not isinstance(baseline_exp, prediction_io.PredictionStore) else "mem",
                           domain, nocs, item, choose_threshold, rank)
    out = solver_ransac_nonlinear(s_ind, e_ind, test_exp, baseline_exp, choose_threshold, num_parts, test_group,
                                  list(problem_ins), rts_all, sub, **solver_kw)
    return sub, out


def merge_sub_pickles(directory, baseline_exp, domain, nocs, item, choose_threshold, n_workers):
    """What eval_pose_err.py:73-77 does when it reads the workers' files back: one dict over all basenames."""
    merged = {}
    for k in range(n_workers):
        fn = result_file_name(directory, baseline_exp, domain, nocs, item, choose_threshold, k)
        if os.path.exists(fn):
            with open(fn, "rb") as fh:
                merged.update(pickle.load(fh))
    return merged
