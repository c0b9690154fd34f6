#!/usr/bin/env python
"""BASELINE.json configs[4]: a mixed stream of the five categories sharded over the GPUs of one box.

    python scripts/bench_mixed.py --clouds 4096                      # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_mixed.py --clouds 100000

Every rank takes its contiguous slice of the stream (pose_multi_process.py:54-63 rule), buckets it by category
(stream.MixedStream), runs each bucket through that category's AncshPipeline from HOST buffers (pinned H2D of the clouds,
D2H of the pose records inside the timed region) and the ranks all-gather the per-cloud records once at the end.
Prints one JSON line on rank 0 (wall-clock clouds/s, max over ranks; a secondary figure -- bench.py is the contract)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=4096)
    ap.add_argument("--unique", type=int, default=64, help="distinct synthetic clouds generated per category (the stream cycles through them)")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--nsample", type=int, default=32)
    ap.add_argument("--hyp", type=int, default=500)
    ap.add_argument("--joint-hyp", type=int, default=200)
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    import torch
    import torch.distributed as dist
    import bench
    from articulated_pose_b200 import stream, synthetic
    from articulated_pose_b200.network import AncshNet
    from articulated_pose_b200.pipeline import AncshPipeline
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    class A:
        pass
    pipes, pool = {}, {}
    for cat in synthetic.ALL_CATEGORIES:
        K = synthetic.n_parts(cat)
        a = A(); a.category = cat; a.no_baseline_net = False; a.nsample = args.nsample

        def feats(kind, w, Pc, K=K):
            return AncshNet(w, K, mixed_pred=(kind == "ancsh"), early_split_nocs=(kind == "ancsh"), nsample=args.nsample,
                            device=dev).features(Pc)
        w_a, w_n = bench.synthetic_weight_sets(K, a, feats)
        pipes[cat] = AncshPipeline(w_a, K, weights_npcs=w_n, nsample=args.nsample, niter_single=args.hyp,
                                   niter_joint=args.joint_hyp, seed=1234 + rank, device=dev)
        pool[cat] = [synthetic.make_cloud(i, cat) for i in range(args.unique)]
    items = synthetic.mixed_stream(args.clouds)

    def load(cat, cid):
        c = pool[cat][cid % args.unique]
        return c["P"], c["joint_cls_gt"]
    ms = stream.MixedStream(pipes, load, batch=args.batch, rank=rank, world=world)
    s, e = ms.my_slice(len(items))
    warm = [it for it in items[s:e]][:min(e - s, 5 * args.batch)]
    stream.MixedStream(pipes, load, batch=args.batch).run(warm, unpack=False)           # buffers, streams, page-in
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    s, e, res = ms.run(items, unpack=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    parts = [synthetic.n_parts(c) for c, _ in items[s:e]]
    # the single collective: fixed-width records (R, s, t per part / joint) padded to the 4-part drawer
    rec = np.zeros((len(res), 1 + 4 * 13 + 3 * 27))
    for i, (r, K) in enumerate(zip(res, parts)):
        v = np.concatenate([np.concatenate([r["single_R"][j].ravel(), [r["single_s"][j]], r["single_t"][j]]) for j in range(K)] +
                           [np.concatenate([r["joint_R0"][j].ravel(), [r["joint_s0"][j]], r["joint_t0"][j], r["joint_R1"][j].ravel(),
                                            [r["joint_s1"][j]], r["joint_t1"][j], [r["joint_score"][j]]]) for j in range(K - 1)])
        rec[i, 0] = K
        rec[i, 1:1 + v.shape[0]] = v
    from articulated_pose_b200 import dist as adist
    t1 = time.perf_counter()
    full = adist.gather_records(rec, device=dev)
    torch.cuda.synchronize()
    t_gather = time.perf_counter() - t1
    tt = torch.tensor([dt + t_gather], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        hist = {c: sum(1 for x, _ in items if x == c) for c in synthetic.ALL_CATEGORIES}
        print(json.dumps({"metric": bench.METRIC, "value": len(items) / float(tt[0]), "unit": bench.UNIT, "n_gpus": world,
                          "config": {"workload": "all 5 categories mixed stream, %d clouds sharded over %d GPU(s), batch %d, "
                                                 "nsample %d, %d/%d hypotheses, host buffers" % (len(items), world, args.batch,
                                                                                               args.nsample, args.hyp, args.joint_hyp),
                                     "category_histogram": hist, "gathered_records": list(full.shape),
                                     "gather_ms": round(1e3 * t_gather, 2)},
                          "seconds": float(tt[0]), "scaling": "strong", "data": "synthetic"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
