"""End-to-end plumbing of AncshPipeline (forward(s) -> pose, predictions never leaving HBM):
the pipeline's poses equal PoseSolver run on the network's own host-copied predictions, the fitted synthetic heads
give a realistic partition, and the whole chain agrees with the CPU oracle chain on the same clouds."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _weights(K, nsample):
    from articulated_pose_b200 import synthetic, weights
    from articulated_pose_b200.network import AncshNet
    Pc, cc = synthetic.make_batch(range(900000, 900008))
    cls = np.stack([c["cls_gt"] for c in cc]); nocs = np.stack([c["nocs_gt"] for c in cc])
    w_a = weights.synthetic_weights(K, True, True, seed=7)
    w_n = weights.synthetic_weights(K, False, False, seed=8)
    w_a = weights.fit_heads(w_a, AncshNet(w_a, K, nsample=nsample).features(Pc), cls, nocs, K, True)
    w_n = weights.fit_heads(w_n, AncshNet(w_n, K, mixed_pred=False, early_split_nocs=False, nsample=nsample).features(Pc),
                            cls, nocs, K, False)
    return w_a, w_n


def test_pipeline_matches_stagewise_and_oracle():
    from articulated_pose_b200 import synthetic
    from articulated_pose_b200.network import AncshNet
    from articulated_pose_b200.pipeline import AncshPipeline
    from articulated_pose_b200.pose import PoseSolver
    from oracle import pnpp, pose_np
    K, ns, B = 3, 32, 3
    w_a, w_n = _weights(K, ns)
    P, clouds = synthetic.make_batch(range(60, 60 + B))
    jc = np.stack([c["joint_cls_gt"] for c in clouds]).astype(np.int32)
    pipe = AncshPipeline(w_a, K, weights_npcs=w_n, nsample=ns, niter_single=128, niter_joint=16, seed=5)
    res = pipe.run(P, jc)
    # (1) stage-wise on host copies of the same predictions
    pa = AncshNet(w_a, K, nsample=ns).forward(P)
    pn = AncshNet(w_n, K, mixed_pred=False, early_split_nocs=False, nsample=ns).forward(P)
    solver = PoseSolver(K, niter_single=128, niter_joint=16, seed=5)
    res2 = solver.solve(P, pn["nocs_per_point"], pn["W"], pa["joint_axis_per_point"], jc)
    for b in range(B):
        assert res[b]["part_count"].min() > 10, res[b]["part_count"]          # fitted heads: no empty parts
        np.testing.assert_array_equal(res[b]["part_count"], res2[b]["part_count"])
        for j in range(K):
            np.testing.assert_array_equal(res[b]["baseline"][j]["rotation"], res2[b]["baseline"][j]["rotation"])
        for j in range(K - 1):
            np.testing.assert_array_equal(res[b]["nonlinear"][j]["rotation1"], res2[b]["nonlinear"][j]["rotation1"])
    # (2) whole chain vs the CPU oracle chain (network outputs agree to 1e-4, so partitions can differ only for
    #     points whose two best class scores are within that margin)
    oa = pnpp.forward(P, w_a, K, nsample=ns)
    on = pnpp.forward(P, w_n, K, nsample=ns, mixed_pred=False, early_split_nocs=False)
    cnt = np.stack([r["part_count"] for r in res])
    idx_s = solver.sample_indices(0, cnt.reshape(-1), 128).reshape(B, K, 128, 3)
    idx_0 = solver.sample_indices(1, np.repeat(cnt[:, :1], K - 1, 1).reshape(-1), 16).reshape(B, K - 1, 16, 3)
    idx_1 = solver.sample_indices(2, cnt[:, 1:].reshape(-1), 16).reshape(B, K - 1, 16, 3)
    checked, checked_joint, clouds_compared = 0, 0, 0
    for b in range(B):
        if not np.array_equal(np.argmax(on["W"][b], 1), np.argmax(pn["W"][b], 1)):
            continue
        clouds_compared += 1
        ref = pose_np.solve_cloud(P[b], on["nocs_per_point"][b], on["W"][b], oa["joint_axis_per_point"][b], jc[b], K, 0.1,
                                  idx_s[b], idx_0[b], idx_1[b])
        for j in range(K):
            # EVERY part whose inlier set coincides must agree: the refit then differs only by the <= 1e-4 differences of
            # the network outputs it is fed with
            if np.array_equal(res[b]["inliers_single"][j], ref["inliers_single"][j]):
                m, r = res[b]["baseline"][j], ref["baseline"][j]
                assert np.abs(m["rotation"] - r["rotation"]).max() < 1e-3, (b, j)
                assert abs(m["scale"] - r["scale"]) < 1e-3 and np.abs(m["translation"] - r["translation"]).max() < 1e-3, (b, j)
                checked += 1
        for j in range(1, K):
            gi, ri = res[b]["inliers_joint"][j - 1], ref["inliers_joint"][j - 1]
            if np.array_equal(gi[0], ri[0]) and np.array_equal(gi[1], ri[1]):
                m, r = res[b]["nonlinear"][j - 1], ref["nonlinear"][j - 1]
                for f in ("rotation0", "rotation1", "translation0", "translation1"):
                    assert np.abs(np.asarray(m[f]) - np.asarray(r[f])).max() < 1e-3, (b, j, f)
                checked_joint += 1
    # same partition on at least two of the three clouds, and most of their parts pick the same inlier set
    assert clouds_compared >= 2 and checked >= 2 * clouds_compared and checked_joint >= 1, (clouds_compared, checked, checked_joint)


def test_shared_geometry_forward_is_bit_identical():
    """ancsh_net_forward_shared: the second network reuses the first network's FPS / ball-query results (same xyz,
    pointnet_util.py:47-49) -- outputs must equal a stand-alone forward bit for bit."""
    import torch
    from articulated_pose_b200 import synthetic, weights
    from articulated_pose_b200.network import AncshNet
    P, _ = synthetic.make_batch(range(40, 44))
    a = AncshNet(weights.synthetic_weights(3, True, True, seed=7), 3, nsample=32)
    b = AncshNet(weights.synthetic_weights(3, False, False, seed=8), 3, mixed_pred=False, early_split_nocs=False, nsample=32)
    Pd = torch.from_numpy(P).cuda()
    alone = {k: v.clone() for k, v in b.forward_device(Pd).items()}
    a.forward_device(Pd)
    shared = b.forward_device(Pd, geometry_from=a)
    torch.cuda.synchronize()
    for k in alone:
        assert torch.equal(alone[k], shared[k]), k


def test_run_many_begin_overlaps_pipelines_without_changing_results():
    """run_many_begin / finish (the mixed stream enqueues every category's pipeline before it collects any result) with
    more batches than buffer slots and two pipelines interleaved: bit-identical to one run() per batch; prepare() only
    pre-allocates."""
    from articulated_pose_b200 import synthetic
    from articulated_pose_b200.pipeline import AncshPipeline
    K, ns, B = 3, 32, 2
    w_a, w_n = _weights(K, ns)
    pipes = [AncshPipeline(w_a, K, weights_npcs=w_n, nsample=ns, niter_single=64, niter_joint=8, seed=11 + i) for i in range(2)]
    n_batches = AncshPipeline.N_SLOTS + 3                       # forces slot reuse inside the enqueue loop
    work = []
    for i in range(n_batches):
        P, clouds = synthetic.make_batch(range(300 + B * i, 300 + B * (i + 1)))
        work.append((P, np.stack([c["joint_cls_gt"] for c in clouds]).astype(np.int32)))
    seeds = list(range(1000, 1000 + n_batches))
    pipes[0].prepare(B, work[0][0].shape[1])
    fins = [p.run_many_begin(work, seeds=seeds) for p in pipes]  # both pipelines enqueued before either is collected
    outs = [f() for f in fins]
    for p, out in zip(pipes, outs):
        assert len(out) == n_batches
        for i in (0, AncshPipeline.N_SLOTS, n_batches - 1):
            Pd, jd = torch.from_numpy(work[i][0]).cuda(), torch.from_numpy(work[i][1]).cuda()
            ref = {k: v.cpu().numpy() for k, v in p.run_device(Pd, jd, seed=seeds[i]).items()}
            assert set(ref) == set(out[i])
            for k in ref:
                np.testing.assert_array_equal(out[i][k], ref[k], err_msg="batch %d %s" % (i, k))
