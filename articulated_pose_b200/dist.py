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
