"""Host-side wire formats and fan-out logic (no GPU): prediction files (lib/prediction_io.py:65-95), the
pose_multi_process.py:52-68 slicing / sub-pickle names, and the bucketing of mixed-category streams."""
import os
import pickle

import numpy as np


def _fake_batch(B=3, N=16, K=3, seed=0):
    rng = np.random.default_rng(seed)
    pred = {"W": rng.uniform(size=(B, N, K)).astype(np.float32),
            "confi_per_point": rng.uniform(size=(B, N, 1)).astype(np.float32),
            "nocs_per_point": rng.uniform(size=(B, N, 3 * K)).astype(np.float32),
            "gocs_per_point": rng.uniform(size=(B, N, 3 * K)).astype(np.float32),
            "heatmap_per_point": rng.uniform(size=(B, N, 1)).astype(np.float32),
            "unitvec_per_point": rng.uniform(size=(B, N, 3)).astype(np.float32),
            "joint_axis_per_point": rng.uniform(size=(B, N, 3)).astype(np.float32),
            "index_per_point": rng.uniform(size=(B, N, 3)).astype(np.float32)}
    inp = {"P": rng.normal(size=(B, N, 3)).astype(np.float32), "cls_gt": rng.integers(0, K, size=(B, N)).astype(np.float32),
           "nocs_gt": rng.uniform(size=(B, N, 3)).astype(np.float32), "nocs_gt_g": rng.uniform(size=(B, N, 3)).astype(np.float32),
           "heatmap_gt": rng.uniform(size=(B, N)).astype(np.float32), "unitvec_gt": rng.uniform(size=(B, N, 3)).astype(np.float32),
           "orient_gt": rng.uniform(size=(B, N, 3)).astype(np.float32), "joint_cls_gt": rng.integers(0, K, size=(B, N)).astype(np.float32)}
    return pred, inp, ["0001_%d_%d" % (b, b + 7) for b in range(B)]


def test_save_batch_nn_files_and_store(tmp_path):
    from articulated_pose_b200 import prediction_io as pio
    pred, inp, names = _fake_batch()
    store = pio.PredictionStore()
    for target in (str(tmp_path), store):
        pio.save_batch_nn("ancsh", pred, inp, names, target, is_mixed=True, W_reduced=False)
        assert sorted(pio.list_predictions(target)) == sorted(n + ".h5" for n in names)
        for b, n in enumerate(names):
            f = pio.load_prediction(target, n)
            assert set(pio.DATASETS) <= set(f.keys())
            np.testing.assert_array_equal(f["instance_per_point"][()], pred["W"][b])          # W_reduced=False: (N,K) scores
            np.testing.assert_array_equal(f["nocs_per_point"][[1, 5], 3:6], pred["nocs_per_point"][b][[1, 5], 3:6])
            np.testing.assert_array_equal(f["P"][:4, :3], inp["P"][b][:4])
            np.testing.assert_array_equal(f["joint_axis_gt"][()], inp["orient_gt"][b])
            np.testing.assert_array_equal(f["confidence_per_point"][()], pred["confi_per_point"][b])
            assert f.attrs["basename"] == n and f.attrs["method_name"] == "ancsh"
            f.close()
    # default W_reduced=True stores argmax labels; is_mixed=False omits gocs (prediction_io.py:70-71, 80-81)
    pio.save_batch_nn("ancsh", pred, inp, names, store)
    f = pio.load_prediction(store, names[1])
    np.testing.assert_array_equal(f["instance_per_point"][()], np.argmax(pred["W"][1], axis=1))
    assert "gocs_per_point" not in f


def test_worker_slices_and_file_names_follow_pose_multi_process():
    from articulated_pose_b200 import evaluation as ev
    for n, cpu in ((100000, 8), (1417, 14), (10, 3), (7, 8), (64, 1)):
        num_per_cpu = int(n / cpu) + 1                                       # pose_multi_process.py:55
        ref = [(num_per_cpu * k, min(num_per_cpu * (k + 1), n)) for k in range(cpu)]
        got = ev.worker_slices(n, cpu)
        covered = []
        for (s, e), (rs, re_) in zip(got, ref):
            assert e == re_ and (s == rs or s == e == n)                     # workers past the end get empty slices
            covered += list(range(s, e))
        assert covered == list(range(n))
    d = "/x/results/pickle/3.9"
    assert ev.result_file_name(d, "3.91", "unseen", "ANCSH", "eyeglasses", 0.1, 5) == \
        d + "/subs" + "/{}_{}_{}_{}_rt_ours_{}_{}.pkl".format("3.91", "unseen", "ANCSH", "eyeglasses", 0.1, 5)   # :60
    assert ev.result_file_name(d, "3.91", "unseen", "ANCSH", "eyeglasses", 0.1) == \
        d + "/{}_{}_{}_{}_rt_ours_{}.pkl".format("3.91", "unseen", "ANCSH", "eyeglasses", 0.1)                   # :48


def test_merge_sub_pickles(tmp_path):
    from articulated_pose_b200 import evaluation as ev
    d = str(tmp_path)
    os.makedirs(os.path.join(d, "subs"))
    for k in (0, 2):
        with open(ev.result_file_name(d, "3.91", "unseen", "ANCSH", "oven", 0.1, k), "wb") as fh:
            pickle.dump({"c%d" % k: {"scale": k}}, fh)
    assert ev.merge_sub_pickles(d, "3.91", "unseen", "ANCSH", "oven", 0.1, 4) == {"c0": {"scale": 0}, "c2": {"scale": 2}}


class _FakePipe:
    def __init__(self, K):
        self.K, self.calls = K, []

    def run_many(self, batches, unpack=True):
        self.calls.append([b[0].shape[0] for b in batches])
        return [[{"id": float(P[i, 0, 0]), "K": self.K} for i in range(P.shape[0])] for P, _ in batches]


def test_mixed_stream_buckets_batches_and_restores_order():
    from articulated_pose_b200 import stream, synthetic
    items = synthetic.mixed_stream(41, seed=5)
    assert {c for c, _ in items} <= set(synthetic.ALL_CATEGORIES) and [i for _, i in items] == list(range(41))
    assert items == synthetic.mixed_stream(41, seed=5)
    b = stream.bucket_by_category(items)
    assert sorted(p for v in b.values() for p in v) == list(range(41))
    assert stream.batches(list(range(10)), 4) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9]]

    def load(cat, cid):
        return np.full((8, 3), cid, np.float32), np.zeros(8, np.int32)
    seen = []
    for world in (1, 2, 3):
        got = {}
        for rank in range(world):
            pipes = {c: _FakePipe(synthetic.n_parts(c)) for c in synthetic.ALL_CATEGORIES}
            ms = stream.MixedStream(pipes, load, batch=4, rank=rank, world=world)
            s, e, res = ms.run(items)
            assert len(res) == e - s
            for pos, r in zip(range(s, e), res):
                assert r["id"] == items[pos][1] and r["K"] == synthetic.n_parts(items[pos][0])
                got[pos] = r
            for p in pipes.values():
                assert all(c == 4 for call in p.calls for c in call)          # ragged batches are padded to one shape
        assert sorted(got) == list(range(41))
        seen.append(got)


def test_record_matrix_pads_to_widest_category():
    from articulated_pose_b200 import dist as adist, stream
    from tests.test_dist_cpu import _fake_results
    r2, r4 = _fake_results([3], 2)[0], _fake_results([4], 4)[0]
    m = stream.record_matrix([r2, r4], [2, 4], 4)
    assert m.shape == (2, 1 + adist.record_width(4)) and m[0, 0] == 2 and m[1, 0] == 4
    np.testing.assert_array_equal(m[0, 1:1 + adist.record_width(2)], adist.pack_records([r2], 2)[0])
    assert not m[0, 1 + adist.record_width(2):].any()
    np.testing.assert_array_equal(m[1, 1:], adist.pack_records([r4], 4)[0])
