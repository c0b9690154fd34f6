"""N > 1 host logic with world_size-2 gloo on CPU: contiguous cloud sharding (pose_multi_process.py:54-63) and the single
end-of-run all-gather of the per-cloud pose records."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_results(ids, K):
    out = []
    for i in ids:
        rng = np.random.default_rng(i)
        out.append({"baseline": [{"rotation": rng.normal(size=(3, 3)), "scale": float(i), "translation": rng.normal(size=3)}
                                 for _ in range(K)],
                    "nonlinear": [{"rotation0": rng.normal(size=(3, 3)), "scale0": 1.0, "translation0": rng.normal(size=3),
                                   "rotation1": rng.normal(size=(3, 3)), "scale1": 2.0, "translation1": rng.normal(size=3),
                                   "score": float(i) / 3} for _ in range(K - 1)]})
    return out


def _worker(rank, world, port, n_total, K, q):
    sys.path.insert(0, ROOT)
    from articulated_pose_b200 import dist as adist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, e = adist.shard_range(n_total, rank, world)
    local = adist.pack_records(_fake_results(range(s, e), K), K)
    full = adist.gather_records(local)
    q.put((rank, s, e, full.numpy()))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything_like_pose_multi_process():
    sys.path.insert(0, ROOT)
    from articulated_pose_b200 import dist as adist
    for n, w in ((100000, 8), (10, 3), (7, 8), (64, 1), (0, 2)):
        spans = [adist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_gather_records_world2_gloo():
    sys.path.insert(0, ROOT)
    from articulated_pose_b200 import dist as adist
    n_total, K, world = 11, 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = adist.pack_records(_fake_results(range(n_total), K), K)
    assert ref.shape == (n_total, adist.record_width(K))
    for rank, s, e, full in got:
        np.testing.assert_array_equal(full, ref)


def test_records_from_arrays_equals_pack_records():
    """The vectorised record builder of the mixed stream equals the per-cloud packer (same column order) and the padded
    stream layout of stream.record_matrix."""
    import numpy as np
    from articulated_pose_b200 import dist as adist, stream
    from articulated_pose_b200.pose import unpack_results
    rng = np.random.default_rng(3)
    for K in (2, 3, 4):
        B, N, J = 5, 16, K - 1
        h = {"single_R": rng.normal(size=(B, K, 3, 3)), "single_s": rng.normal(size=(B, K)), "single_t": rng.normal(size=(B, K, 3)),
             "single_score": rng.integers(0, 9, size=(B, K)).astype(np.int32), "single_inliers": np.zeros((B, K, N), np.uint8),
             "joint_R0": rng.normal(size=(B, J, 3, 3)), "joint_s0": rng.normal(size=(B, J)), "joint_t0": rng.normal(size=(B, J, 3)),
             "joint_R1": rng.normal(size=(B, J, 3, 3)), "joint_s1": rng.normal(size=(B, J)), "joint_t1": rng.normal(size=(B, J, 3)),
             "joint_score": rng.normal(size=(B, J)), "joint_inliers0": np.zeros((B, J, N), np.uint8),
             "joint_inliers1": np.zeros((B, J, N), np.uint8), "part_count": np.full((B, K), 4, np.int32),
             "status": np.zeros((B, K), np.int32)}
        res = unpack_results(h, K)
        np.testing.assert_array_equal(adist.records_from_arrays(h, K), adist.pack_records(res, K))
        np.testing.assert_array_equal(adist.records_from_arrays(h, K, 4), stream.record_matrix(res, [K] * B, 4))
