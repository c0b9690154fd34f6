"""Mixed-category stream (BASELINE.json configs[4]) and the pose-level drop-ins (solver_ransac_nonlinear /
pose_multi_process fan-out) on the GPU."""
import copy
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _pipelines(cats, nsample=32):
    from articulated_pose_b200 import synthetic, weights
    from articulated_pose_b200.pipeline import AncshPipeline
    pipes = {}
    for c in cats:
        K = synthetic.n_parts(c)
        pipes[c] = AncshPipeline(weights.synthetic_weights(K, True, True, seed=7), K,
                                 weights_npcs=weights.synthetic_weights(K, False, False, seed=8), nsample=nsample,
                                 niter_single=64, niter_joint=8, seed=3)
    return pipes


def test_mixed_stream_equals_per_category_runs():
    """Bucketing must not change what a cloud's launch computes: the network-derived partition of every cloud (argmax
    labels -> part_count, independent of the RANSAC samples) equals the one a single-cloud run of its category's
    pipeline gives, results come back in stream order with the category's K, and two ranks cover the stream."""
    from articulated_pose_b200 import stream, synthetic
    items = synthetic.mixed_stream(23, seed=11)
    cats = sorted({c for c, _ in items})
    assert len(cats) >= 4
    pipes = _pipelines(cats)
    cache = {}

    def load(cat, cid):
        if (cat, cid) not in cache:
            c = synthetic.make_cloud(cid, cat)
            cache[(cat, cid)] = (c["P"], c["joint_cls_gt"])
        return cache[(cat, cid)]

    s, e, res = stream.MixedStream(pipes, load, batch=4).run(items)
    assert (s, e) == (0, len(items)) and all(r is not None for r in res)
    for (cat, cid), r in zip(items, res):
        K = synthetic.n_parts(cat)
        P, jc = load(cat, cid)
        assert len(r["baseline"]) == K and len(r["nonlinear"]) == K - 1
        assert int(r["part_count"].sum()) == P.shape[0]
        alone = pipes[cat].run(P[None], jc[None])[0]
        np.testing.assert_array_equal(r["part_count"], alone["part_count"])
    # two ranks: disjoint contiguous slices whose union is the stream; same partitions
    parts = [synthetic.n_parts(c) for c, _ in items]
    rows = []
    for rank in range(2):
        ms = stream.MixedStream(pipes, load, batch=4, rank=rank, world=2)
        s, e, rr = ms.run(items)
        for pos, r in zip(range(s, e), rr):
            np.testing.assert_array_equal(r["part_count"], res[pos]["part_count"])
        rows.append(stream.record_matrix(rr, parts[s:e], 4))
    full = np.concatenate(rows)
    assert full.shape[0] == len(items) and list(full[:, 0].astype(int)) == parts


def test_solver_ransac_nonlinear_drop_in(tmp_path):
    """prediction files (save_batch_nn) -> solver_ransac_nonlinear -> the reference's pickle schema; on teacher
    predictions the recovered poses are close to the ground truth and both workers' sub-pickles merge to the full set."""
    from articulated_pose_b200 import evaluation as ev, prediction_io as pio, synthetic
    from articulated_pose_b200.pose import compute_gt_pose
    K, ids = 3, list(range(70, 75))
    clouds = [synthetic.make_cloud(i) for i in ids]
    preds = [synthetic.teacher_predictions(c) for c in clouds]
    names = ["%04d_0_%d" % (i, i) for i in ids]
    N = clouds[0]["P"].shape[0]
    z = lambda *sh: np.zeros((len(ids),) + sh, np.float32)
    pred = {"W": np.stack([p["W"] for p in preds]), "confi_per_point": z(N, 1),
            "nocs_per_point": np.stack([p["nocs_per_point"] for p in preds]), "gocs_per_point": z(N, 3 * K),
            "heatmap_per_point": z(N, 1), "unitvec_per_point": z(N, 3),
            "joint_axis_per_point": np.stack([p["joint_axis_per_point"] for p in preds]), "index_per_point": z(N, 3)}
    inp = {"P": np.stack([c["P"] for c in clouds]), "cls_gt": np.stack([c["cls_gt"] for c in clouds]),
           "nocs_gt": np.stack([c["nocs_gt"] for c in clouds]), "nocs_gt_g": z(N, 3), "heatmap_gt": z(N), "unitvec_gt": z(N, 3),
           "orient_gt": z(N, 3), "joint_cls_gt": np.stack([c["joint_cls_gt"] for c in clouds])}
    exp_dir = tmp_path / "test_pred" / "3.9"
    os.makedirs(exp_dir)
    pio.save_batch_nn("ancsh", pred, inp, names, str(exp_dir), is_mixed=True, W_reduced=False)      # lib/network.py:297-305
    store = pio.PredictionStore()
    pio.save_batch_nn("npcs", pred, inp, names, store, W_reduced=False)
    rts_all = {}
    for n, c in zip(names, clouds):
        rts_all[n] = compute_gt_pose(c["P"], c["nocs_gt"], c["cls_gt"], K)                           # compute_gt_pose.py:80-97
        rts_all[n]["nocs_err"] = [0.0] * K
    rts_fresh = copy.deepcopy(rts_all)           # rts_all is read only: the reference builds a fresh rts_dict per cloud (:346-352)
    test_group = pio.list_predictions(str(exp_dir))
    assert test_group == sorted(n + ".h5" for n in names)
    out_dir = str(tmp_path / "pickle" / "3.9")
    merged = {}
    for rank in range(2):
        fn, out = ev.pose_multi_process("3.9", store, K, test_group, rts_all, out_dir, item="eyeglasses", rank=rank, world=2,
                                        pred_root=str(tmp_path / "test_pred"), niter_single=500, niter_joint=50, seed=9)
        assert os.path.basename(fn) == "mem_unseen_ANCSH_eyeglasses_rt_ours_0.1_%d.pkl" % rank and os.path.exists(fn)
        with open(fn, "rb") as fh:
            assert set(pickle.load(fh)) == set(out)
        merged.update(out)
    assert set(merged) == set(names)
    for n in names:
        e = merged[n]
        assert set(e) == {"scale", "rotation", "translation", "xyz_err", "rpy_err", "scale_err"}     # the reference's pickle keys
        for k in ("scale", "rotation", "translation"):
            assert set(e[k]) == {"gt", "baseline", "nonlinear"}
        assert len(e["rpy_err"]["baseline"]) == K and len(e["rpy_err"]["nonlinear"]) == K       # part 0 + parts 1..K-1
        assert max(e["rpy_err"]["baseline"]) < 10.0 and max(e["xyz_err"]["baseline"]) < 0.1, (n, e["rpy_err"], e["xyz_err"])
        assert max(e["rpy_err"]["nonlinear"]) < 10.0 and max(e["xyz_err"]["nonlinear"]) < 0.1, (n, e["rpy_err"], e["xyz_err"])
    # rts_all is left untouched, so a second pass (another threshold, a retry) over the same dict works
    assert all(set(rts_all[n]) == set(rts_fresh[n]) for n in names)
    assert np.shape(rts_all[names[0]]["scale"]["gt"][0]) == np.shape(rts_fresh[names[0]]["scale"]["gt"][0])
    # problem instances are skipped (parallel_ancsh_pose.py:217-219)
    out = ev.solver_ransac_nonlinear(0, len(test_group), "3.9", store, 0.1, K, test_group, [names[0].split("_")[0]], rts_fresh,
                                     None, pred_root=str(tmp_path / "test_pred"), niter_single=64, niter_joint=8)
    assert set(out) == set(names[1:])
