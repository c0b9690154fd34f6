"""Cloud sharding over the GPUs of one box.

Every cloud is independent through all three stages (the native ops index by batch only, BN is inference mode, the
pose loop is per basename), so the path shards embarrassingly: contiguous slices of the test group per rank -- the same
rule `evaluation/pose_multi_process.py:54-63` uses for its forked workers (`num_per_cpu = int(len/cpu)+1`) -- weights
replicated, and ONE collective at the end that gathers the fixed-size per-cloud pose records.  Works with the `nccl`
backend on GPUs and with `gloo` on CPU tensors (tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """[start, end) of rank's contiguous slice, pose_multi_process.py:55-62: num_per = int(n/world)+1."""
    num_per = int(n_items / world) + 1
    start = min(num_per * rank, n_items)
    return start, min(num_per * (rank + 1), n_items)


def pack_records(results, K):
    """List of per-cloud result dicts (pose.unpack_results) -> float64 (n, K*13 + (K-1)*27) record matrix:
    per part [R(9), s, t(3)], per joint [R0(9), s0, t0(3), R1(9), s1, t1(3), score]."""
    rows = []
    for r in results:
        v = []
        for m in r["baseline"]:
            v += list(np.asarray(m["rotation"]).ravel()) + [m["scale"]] + list(m["translation"])
        for m in r["nonlinear"]:
            v += list(np.asarray(m["rotation0"]).ravel()) + [m["scale0"]] + list(m["translation0"])
            v += list(np.asarray(m["rotation1"]).ravel()) + [m["scale1"]] + list(m["translation1"]) + [m.get("score", 0.0)]
        rows.append(v)
    width = K * 13 + (K - 1) * 27
    return np.asarray(rows, np.float64).reshape(len(rows), width)


def record_width(K):
    return K * 13 + (K - 1) * 27


def records_from_arrays(h, K, k_max=None):
    """The rows pack_records builds, straight from a batch's pose arrays (ancsh_pose_out_t fields as host ndarrays, e.g. one
    entry of AncshPipeline.run_many(..., unpack=False)) without a Python loop over clouds.  With k_max the rows get the
    stream layout of stream.record_matrix: [K, record zero-padded to the width of k_max parts]."""
    B = h["single_s"].shape[0]
    parts = np.concatenate([h["single_R"].reshape(B, K, 9), h["single_s"].reshape(B, K, 1), h["single_t"].reshape(B, K, 3)], 2)
    cols = [parts.reshape(B, K * 13)]
    if K > 1:
        J = K - 1
        joints = np.concatenate([h["joint_R0"].reshape(B, J, 9), h["joint_s0"].reshape(B, J, 1), h["joint_t0"].reshape(B, J, 3),
                                 h["joint_R1"].reshape(B, J, 9), h["joint_s1"].reshape(B, J, 1), h["joint_t1"].reshape(B, J, 3),
                                 h["joint_score"].reshape(B, J, 1)], 2)
        cols.append(joints.reshape(B, J * 27))
    rec = np.concatenate(cols, 1).astype(np.float64)
    if k_max is None:
        return rec
    out = np.zeros((B, 1 + record_width(k_max)), np.float64)
    out[:, 0] = K
    out[:, 1:1 + rec.shape[1]] = rec
    return out


def gather_records(local, device=None):
    """All-gather the per-rank record matrices (ragged: the contiguous slices differ in length) into the full matrix in
    rank (= cloud) order on every rank.  One small all-gather of the row counts, one of the padded records."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    loc = torch.as_tensor(local, dtype=torch.float64, device=device)
    if world == 1:
        return loc
    cnt = torch.tensor([loc.shape[0]], dtype=torch.int64, device=device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt)
    counts = [int(c.item()) for c in cnts]
    buf = torch.zeros((max(counts), loc.shape[1]), dtype=torch.float64, device=device)
    buf[:loc.shape[0]] = loc
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([out[r][:counts[r]] for r in range(world)], 0)
