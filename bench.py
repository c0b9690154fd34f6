#!/usr/bin/env python
"""bench.py -- clouds/sec of the ANCSH hot path on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs[0..4]; `full` is the configuration the metric is quoted on and the default):
  cpu64    configs[0]  eyeglasses, 64 clouds per step, full pipeline (the reference's own CPU-runnable case; its CPU arm
                       fans the clouds out over forked workers exactly like pose_multi_process.py)
  forward  configs[1]  eyeglasses N=1024 nsample 32, device batches of 256 clouds, PN++ forward + heads only
  full     configs[2]  eyeglasses N=1024 nsample 32, ANCSH + NPCS-baseline forwards + 3 x RANSAC(500) + 2 x joint solve(200, LM)
  drawer   configs[3]  drawer (prismatic, 4 parts) N=2048 nsample 64, full pipeline
  mixed    configs[4]  all five categories as one shuffled stream of --clouds (100 000) clouds sharded over the ranks
                       (strong scaling), host buffers, one timed all-gather of the pose records

One "step" = one pass of the hot path over one job batch of `chunks x batch` synthetic clouds per rank (16 x 256 = 4096 for the
default workload: 20 steps ~ 2.7 s timed), run as `chunks` device batches of `batch` clouds pipelined over the buffer slots of AncshPipeline (weak scaling: every rank owns its
own clouds; clouds are independent).  Every device batch of a step holds different clouds and every batch of
the run draws its RANSAC hypotheses from its own Philox key.  Prints ONE JSON line on rank 0.

  value     device-timed clouds/s, inputs resident in HBM (CUDA events around the K steps on the launch stream, L2
            flushed between steps, max over ranks)
  e2e       the same steps through the public host API (AncshPipeline.run_many) with HOST buffers: pinned H2D of the
            clouds and D2H of the pose results inside the timed region
  roofline  the stage with the largest share of a serialized step (per-stage CUDA events of a profiling pass in the same
            run): the joint LM stage against the FP64 peak measured in this process, or a grouped-MLP stage against the
            measured bf16 tensor peak (MEASURED_PEAKS.json); `roofline_tensor` always reports the slowest grouped MLP
  cpu_baseline / --impl reference   the CPU oracle (oracle/: line-for-line restatement of the reference's network ops --
            TF1 cannot be installed and the reference has no CPU kernels for FPS / ball query -- and of its numpy/scipy
            pose code) on this box's host cores
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clouds/sec end-to-end (PN++ fwd + RANSAC + joint solve)"
UNIT = "clouds/s"
CALIB_CLOUDS = 16

# name -> (category, nsample, stages, device batch, chunks per step, BASELINE.json config index)
WORKLOADS = {
    "cpu64": ("eyeglasses", 32, "full", 64, 1, 0),
    "forward": ("eyeglasses", 32, "forward", 256, 16, 1),
    "full": ("eyeglasses", 32, "full", 256, 16, 2),
    "drawer": ("drawer", 64, "full", 128, 8, 3),
    "mixed": (None, 32, "full", 256, 0, 4),
}

# FLOP model of the joint LM stage (FMA = 2): residual evaluation of the 3 + 3 points and the joint
# row, Jacobian evaluation with its normal equations, the post-Jacobian factorisation (6x6 Cholesky + three triangular
# solves), one lmpar iteration (Cholesky + three solves), and the step bookkeeping.
LM_FLOPS = {"cost": 342.0, "jacobian": 2020.0, "jac_phase": 250.0, "lmpar_iter": 240.0, "finish": 80.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="full", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="clouds per device batch (0 = the workload's)")
    ap.add_argument("--chunks", type=int, default=0, help="device batches per step and rank (0 = the workload's)")
    ap.add_argument("--category", default=None)
    ap.add_argument("--nsample", type=int, default=0, help="0 = the workload's (BASELINE K=32; reference default: 64)")
    ap.add_argument("--hyp", type=int, default=500, help="RANSAC hypotheses per part (BASELINE config: 500; reference 10000)")
    ap.add_argument("--joint-hyp", type=int, default=200)
    ap.add_argument("--stages", default=None, choices=["forward", "full"])
    ap.add_argument("--clouds", type=int, default=100000, help="mixed workload: length of the whole stream")
    ap.add_argument("--unique", type=int, default=64, help="mixed workload: distinct synthetic clouds per category")
    ap.add_argument("--no-baseline-net", action="store_true", help="USE_BASELINE=False: one forward per cloud")
    ap.add_argument("--cpu-sample", type=int, default=0, help="clouds in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "f32"],
                    help="f16x3: grouped MLPs on tcgen05 (fp16 hi/lo split, f32 accumulate); f32: CUDA-core FMA kernels")
    a = ap.parse_args()
    cat, ns, stages, batch, chunks, cfg = WORKLOADS[a.workload]
    a.category = a.category or cat or "eyeglasses"
    a.nsample = a.nsample or ns
    a.stages = a.stages or stages
    a.batch = a.batch or batch
    a.chunks = a.chunks or chunks
    a.config_index = cfg
    return a


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_median": statistics.median(pw) if pw else None}


def stage_flops(net, B, N):
    """Algorithmic FLOPs (2*MAC, dense, unpadded shapes) per forward stage for a batch of B clouds."""
    def chain(rows, layers):
        return 2.0 * rows * sum(l.cin * l.cout for l in layers)
    L = net.layers
    m1, m2, s1, s2 = net.npoint1, net.npoint2, net.nsample1, net.nsample2
    f = {
        "sa1": chain(m1 * s1, [L["sa1[0]"], L["sa1[1]"], L["sa1[2]"]]),
        "sa2": chain(m2 * s2, [L["sa2[0]"], L["sa2[1]"], L["sa2[2]"]]),
        "sa3": chain(m2, [L["sa3[0]"], L["sa3[1]"], L["sa3[2]"]]),
        "fp1": chain(1, [L["fp1_global"]]) + chain(m2, [L["fp1[0]"], L["fp1[1]"]]),
        "fp2": chain(m1, [L["fp2[0]"], L["fp2[1]"]]),
        "fp3_heads": chain(N, [L["fp3[0]"], L["fp3[1]"], L["fp3[2]"], L["fc1"], L["nocs_heads"], L["fc3[0]"],
                               L["fc3[1]"], L["joint_heads"]]),
    }
    return {k: v * B for k, v in f.items()}


def workload_name(args, world=1):
    cat = args.category
    n = 1024 if cat == "eyeglasses" else 2048
    if args.workload == "mixed":
        return ("all 5 categories mixed stream, %d clouds sharded over %d GPU(s), device batches of %d, nsample %d, "
                "RANSAC(%d hyp/part) + joint solve(%d hyp/joint, LM), host buffers") % (args.clouds, world, args.batch,
                                                                                      args.nsample, args.hyp, args.joint_hyp)
    if args.stages == "forward":
        return "%s ANCSH, N=%d nsample=%d, device batch=%d clouds, PN++ forward + heads" % (cat, n, args.nsample, args.batch)
    return ("%s ANCSH, N=%d nsample=%d, device batch=%d clouds, full pipeline: %s + RANSAC(%d hyp/part) + "
            "joint solve(%d hyp/joint, LM)") % (cat, n, args.nsample, args.batch,
                                                "ANCSH forward" if args.no_baseline_net else "ANCSH + NPCS-baseline forwards",
                                                args.hyp, args.joint_hyp)


def synthetic_weight_sets(K, args, feature_fn=None, category=None):
    """Seeded random trunks; when feature_fn(net_kind, weights, P) -> (M,128) is given, the linear segmentation /
    NOCS heads are fitted on CALIB_CLOUDS synthetic clouds (weights.fit_heads) so that the part partition -- and
    with it the pose-stage workload -- is realistic."""
    from articulated_pose_b200 import synthetic, weights
    w_a = weights.synthetic_weights(K, True, True, seed=7)
    w_n = None if args.no_baseline_net else weights.synthetic_weights(K, False, False, seed=8)
    if feature_fn is not None:
        Pc, cc = synthetic.make_batch(range(900000, 900000 + CALIB_CLOUDS), category or args.category)
        cls = np.stack([c["cls_gt"] for c in cc])
        nocs = np.stack([c["nocs_gt"] for c in cc])
        w_a = weights.fit_heads(w_a, feature_fn("ancsh", w_a, Pc), cls, nocs, K, early_split_nocs=True)
        if w_n is not None:
            w_n = weights.fit_heads(w_n, feature_fn("npcs", w_n, Pc), cls, nocs, K, early_split_nocs=False)
    return w_a, w_n


def cpu_reference_run(args, K, P, jc, w_a, w_n, n_clouds, steps):
    """Times the CPU oracle on `n_clouds` clouds per step.  Returns (clouds/s, seconds, cores, description)."""
    from oracle import pipeline_cpu
    cores = os.cpu_count() or 1
    fanout = args.workload == "cpu64" and args.stages == "full"
    t0 = time.perf_counter()
    done = 0
    for s in range(steps):
        lo = (s * n_clouds) % max(1, P.shape[0] - n_clouds + 1)
        if fanout:
            done += pipeline_cpu.run_clouds_fanout(P[lo:lo + n_clouds], jc[lo:lo + n_clouds], w_a, w_n, K, args.nsample, 0.1,
                                                   args.hyp, args.joint_hyp, seed=s)
        else:
            done += pipeline_cpu.run_clouds(P[lo:lo + n_clouds], jc[lo:lo + n_clouds], w_a, w_n, K, args.nsample, 0.1, args.hyp,
                                            args.joint_hyp, stages=args.stages, seed=s)
    dt = time.perf_counter() - t0
    if fanout:
        desc = ("%d cloud(s)/step x %d step(s): both forwards per cloud (native ops OpenMP over %d threads), then ONE "
                "pose_multi_process.py-style fan-out: contiguous cloud slices over %d forked workers (cpu_count-2), each "
                "solving its clouds serially") % (n_clouds, steps, cores, pipeline_cpu.workers())
    else:
        desc = ("%d cloud(s)/step x %d step(s) of the same workload; network ops OpenMP over %d threads, RANSAC hypotheses "
                "over %d forked workers (pose_multi_process.py uses cpu_count-2)") % (n_clouds, steps, cores,
                                                                                     pipeline_cpu.workers())
    return done / dt, dt, cores, desc


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; see module docstring)
    with all host threads, bounded sample per step.  Rank 0 only."""
    if rank != 0:
        return
    from articulated_pose_b200 import synthetic
    from oracle import pipeline_cpu, pnpp
    K = len(synthetic.CATEGORIES[args.category]["boxes"])

    def feats(kind, w, Pc):      # trunk feature through the oracle (this arm never touches the GPU)
        tr = {}
        pnpp.set_threads(os.cpu_count() or 1)
        pnpp.forward(Pc, w, K, nsample=args.nsample, mixed_pred=(kind == "ancsh"), early_split_nocs=(kind == "ancsh"), trace=tr)
        return tr["net"]
    w_a, w_n = synthetic_weight_sets(K, args, feats if args.stages == "full" else None)
    # bounded sample per step: cpu64 runs its 64 clouds (the configuration IS CPU sized); the others 1 cloud (2 forward only)
    n_s = args.cpu_sample or (args.batch if args.workload == "cpu64" else (2 if args.stages == "forward" else 1))
    P, clouds = synthetic.make_batch(range(max(n_s, 4)), args.category)
    jc = np.stack([c["joint_cls_gt"] for c in clouds])
    steps = args.steps if args.workload != "cpu64" else min(args.steps, 2)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(args, K, P, jc, w_a, w_n, 1, 1)
    val, dt, cores, desc = cpu_reference_run(args, K, P, jc, w_a, w_n, n_s, steps)
    pipeline_cpu.close_pool()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 network / f64 pose", "data": "synthetic",
            "config": {"workload": workload_name(args), "baseline_config": args.config_index, "stages": args.stages,
                       "sample": desc},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def fp64_peak_tflops(dev):
    """DFMA peak of this GPU measured in-process (ancsh_diag_fp64_fma: 8 independent chains per thread, 8 blocks of 256
    threads per SM), best of 5."""
    import torch
    from articulated_pose_b200 import _lib
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks, iters = sms * 8, 4096
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.ancsh_diag_fp64_fma(blocks, iters, sink.data_ptr(), st), "ancsh_diag_fp64_fma")
        e1.record()
        torch.cuda.synchronize()
        best = max(best, blocks * 256 * 8 * iters * 2.0 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_mixed(args, rank, local_rank, world):
    """BASELINE configs[4]: one shuffled five-category stream, sharded over the ranks (pose_multi_process.py:54-63 slice
    rule), every rank buckets its slice by category and pushes the buckets through that category's pipeline from HOST
    buffers; the per-cloud pose records are all-gathered once per step inside the timed region."""
    import torch
    import torch.distributed as dist
    from articulated_pose_b200 import _lib, stream, synthetic
    from articulated_pose_b200 import dist as adist
    from articulated_pose_b200.network import AncshNet
    from articulated_pose_b200.pipeline import AncshPipeline
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pipes, pool = {}, {}
    for cat in synthetic.ALL_CATEGORIES:
        K = synthetic.n_parts(cat)

        def feats(kind, w, Pc, K=K):
            return AncshNet(w, K, mixed_pred=(kind == "ancsh"), early_split_nocs=(kind == "ancsh"), nsample=args.nsample,
                            device=dev, precision=args.precision).features(Pc)
        w_a, w_n = synthetic_weight_sets(K, args, feats, category=cat)
        pipes[cat] = AncshPipeline(w_a, K, weights_npcs=w_n, nsample=args.nsample, niter_single=args.hyp,
                                   niter_joint=args.joint_hyp, seed=1234 + rank, device=dev, precision=args.precision)
        cl = [synthetic.make_cloud(i, cat) for i in range(args.unique)]
        pool[cat] = (np.stack([c["P"] for c in cl]).astype(np.float32), np.stack([c["joint_cls_gt"] for c in cl]).astype(np.int32))
        pipes[cat].prepare(args.batch, pool[cat][0].shape[1])     # all slots' buffers / workspaces before the stream starts
    items = synthetic.mixed_stream(args.clouds)
    steps = max(1, args.steps)
    seg = [(len(items) * k // steps, len(items) * (k + 1) // steps) for k in range(steps)]

    def load_batch(cat, cids):
        ids = np.asarray(cids) % args.unique
        return pool[cat][0][ids], pool[cat][1][ids]
    ms = stream.MixedStream(pipes, load_batch=load_batch, batch=args.batch, rank=rank, world=world)
    k_max = max(synthetic.n_parts(c) for c in synthetic.ALL_CATEGORIES)

    def one_step(part):
        s, e, rec = ms.run(part, records_k_max=k_max)          # gather-ready records, no per-cloud Python work
        t1 = time.perf_counter()
        full = adist.gather_records(rec, device=dev)
        torch.cuda.synchronize()
        return int(full.shape[0]), int(full.shape[1]), time.perf_counter() - t1

    warm = items[:min(len(items), world * 3 * args.batch)]
    for _ in range(max(1, min(args.warmup, 2))):
        one_step(warm)
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.start()
    l0 = _lib.ancsh_launch_count()
    t0 = time.perf_counter()
    gathered, width, t_gather = 0, 0, 0.0
    for a, b in seg:
        n, width, tg = one_step(items[a:b])
        gathered += n
        t_gather += tg
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = _lib.ancsh_launch_count() - l0
    clocks = sampler.stop()
    tt = torch.tensor([dt, t_gather], dtype=torch.float64, device=dev)
    per_rank = [dt]
    if world > 1:
        allt = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allt, tt)
        per_rank = [float(t[0]) for t in allt]
        t_gather = max(float(t[1]) for t in allt)
    dt = max(per_rank)
    if rank == 0:
        hist = {c: sum(1 for x, _ in items if x == c) for c in synthetic.ALL_CATEGORIES}
        line = {"metric": METRIC, "value": len(items) / dt, "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f16x3->f32 network / f64 pose", "data": "synthetic",
                "config": {"workload": workload_name(args, world), "baseline_config": 4, "stages": "full",
                           "category_histogram": hist, "all_gathered_records": gathered, "record_width_f64": width,
                           "gather_ms_total": round(1e3 * t_gather, 2), "gather_bytes_total": gathered * width * 8,
                           "timing": "wall clock around the K segments incl. host-side bucketing, H2D, D2H and the gather; max over ranks",
                           "seconds_per_rank": {"min": round(min(per_rank), 4), "max": round(max(per_rank), 4)},
                           "l2_flush_between_steps": False, "inputs_larger_than_l2": True},
                "clocks": clocks,
                "e2e": {"value": len(items) / dt, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
                "gpu_launches": int(launches)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload == "mixed":
        run_mixed(args, rank, local_rank, world)
        return

    import torch
    import torch.distributed as dist
    from articulated_pose_b200 import _lib, synthetic
    from articulated_pose_b200.network import AncshNet
    from articulated_pose_b200.pipeline import AncshPipeline

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    K = len(synthetic.CATEGORIES[args.category]["boxes"])
    B, C = args.batch, args.chunks
    full = args.stages == "full"

    def feats(kind, w, Pc):
        n = AncshNet(w, K, mixed_pred=(kind == "ancsh"), early_split_nocs=(kind == "ancsh"), nsample=args.nsample, device=dev,
                     precision=args.precision)
        return n.features(Pc)
    w_a, w_n = synthetic_weight_sets(K, args, feats if full else None)
    pipe = AncshPipeline(w_a, K, weights_npcs=w_n if full else None, use_baseline=full and not args.no_baseline_net,
                         nsample=args.nsample, niter_single=args.hyp, niter_joint=args.joint_hyp, seed=1234 + rank,
                         device=dev, precision=args.precision)
    # the job batch of a step: C device batches of B DIFFERENT clouds each (ids disjoint across ranks)
    P_host, clouds = synthetic.make_batch(range(rank * B * C, (rank + 1) * B * C), args.category)
    jc_host = np.stack([c["joint_cls_gt"] for c in clouds]).astype(np.int32)
    N = P_host.shape[1]
    P_host = P_host.reshape(C, B, N, 3)
    jc_host = jc_host.reshape(C, B, N)
    P_dev = torch.from_numpy(P_host).to(dev)
    jc_dev = torch.from_numpy(jc_host).to(dev)
    out_fwd = pipe.net.alloc_outputs(B, N)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    two_nets = pipe.net_npcs is not None

    nst, npst = len(_lib.NET_STAGES), len(_lib.POSE_STAGES)
    seed_of = lambda step, c: 1000003 * (rank + 1) + 131 * step + c      # one Philox key per device batch of the run

    def run_chunk(c, step=0, ev=None):
        if full:
            return pipe.run_device(P_dev[c], jc_dev[c], net_events=ev[0] if ev else None, net_b_events=ev[1] if ev else None,
                                   pose_events=ev[2] if ev else None, seed=seed_of(step, c))
        return pipe.net.forward_device(P_dev[c], out_fwd, stage_events=ev[0] if ev else None)

    def run_step(step):
        """One step on the device: C batches through the slot pipeline (the pose stage of a batch overlaps later forwards)."""
        res = None
        for c in range(C):
            if full:
                res = pipe.submit(P_dev[c], jc_dev[c], slot=(step * C + c) % pipe.N_SLOTS, seed=seed_of(step, c))
            else:
                res = run_chunk(c)
        return res

    W = max(args.warmup, 3)
    res = run_chunk(0)
    for s in range(W):
        res = run_step(-1 - s)
    if full:
        pipe.join()
    torch.cuda.synchronize()
    part_hist = res["part_count"].float().mean(0).tolist() if full else None

    # ---- serialized profiling pass: per-stage CUDA events (stage durations without cross-batch overlap) --------
    n_prof = min(5, C * args.steps)
    evs = [(_lib.EventList(nst + 1), _lib.EventList(nst + 1), _lib.EventList(npst + 1)) for _ in range(n_prof)]
    lm_tot = {"nfev": 0.0, "njev": 0.0, "nlm": 0.0}
    for i in range(n_prof):
        flush.zero_()
        run_chunk(i % C, step=10 ** 6 + i, ev=evs[i])
        if full and K > 1:
            inter = pipe.pose.intermediates()
            lm_tot["nfev"] += float(inter["joint_nfev"].sum())
            lm_tot["njev"] += float(inter["lm_njev_total"])
            lm_tot["nlm"] += float(inter["lm_lmpar_total"])
    torch.cuda.synchronize()

    def ctypes_elapsed(ea, i, eb, j):
        import ctypes
        ms = ctypes.c_float()
        _lib.check(_lib.ancsh_event_elapsed_ms(ea.arr[i], eb.arr[j], ctypes.byref(ms)), "elapsed")
        return float(ms.value)

    serial_ms = sum(ctypes_elapsed(e[0], 0, e[2], npst) if full else e[0].elapsed_ms(0, nst) for e in evs) / n_prof
    stage_ms = {nm: sum(e[0].elapsed_ms(i, i + 1) for e in evs) / n_prof for i, nm in enumerate(_lib.NET_STAGES)}
    extra_ms = {}
    if full:
        if two_nets:
            for i, nm in enumerate(_lib.NET_STAGES):
                extra_ms["npcs_" + nm] = sum(e[1].elapsed_ms(i, i + 1) for e in evs) / n_prof
        for i, nm in enumerate(_lib.POSE_STAGES):
            extra_ms["pose_" + nm] = sum(e[2].elapsed_ms(i, i + 1) for e in evs) / n_prof

    fp64_peak = fp64_peak_tflops(dev) if full else None

    # ---- timed region: EXACTLY K steps ------------------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    l0 = _lib.ancsh_launch_count()
    t_wall0 = time.perf_counter()
    ev0.record()
    host_enqueue_s = []
    for s in range(args.steps):
        flush.zero_()                       # L2 flush between timed steps (256 MB memset, ~0.1 ms, counted)
        th = time.perf_counter()
        res = run_step(s)
        host_enqueue_s.append(time.perf_counter() - th)     # host time to enqueue a step: the first step starts on an empty
                                                            # launch queue (pure host cost), later ones are throttled by the device
    if full:
        pipe.join()
    ev1.record()
    torch.cuda.synchronize()
    launches = _lib.ancsh_launch_count() - l0
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    nf = pipe.pose.intermediates()["joint_nfev"].float() if full and K > 1 else None
    lm_stats = None
    if nf is not None:
        q = torch.quantile(nf.flatten(), torch.tensor([0.5, 0.99, 1.0], device=nf.device)).tolist()
        lm_stats = {"nfev_p50": q[0], "nfev_p99": q[1], "nfev_max": q[2], "nfev_mean": float(nf.mean())}

    # ---- end to end through the public host API (host buffers, H2D + D2H inside the timed region) ----
    host_batches = [(P_host[c], jc_host[c]) for c in range(C)]

    def e2e_steps(first, count):
        """`count` steps through the host API as ONE stream of device batches (AncshPipeline.run_many keeps the slot pipeline
        full across step boundaries; every batch is copied H2D from pinned memory and its pose records are copied back)."""
        if full:
            seeds = [seed_of(first + s, c) for s in range(count) for c in range(C)]
            return pipe.run_many(host_batches * count, seeds=seeds)[-1]
        r = None
        for _ in range(count):
            for c in range(C):
                r = pipe.net.forward(P_host[c], copy=False)
        return r
    e2e_steps(-W, W)   # warm-up of the timed call: run_many pins its staging buffers and creates its copy stream on first use
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    r = e2e_steps(0, args.steps)
    e2e_s = time.perf_counter() - t0
    h2d = C * (P_host[0].nbytes + (jc_host[0].nbytes if full else 0))
    d2h = C * sum(v.nbytes for v in r.values())

    gathered, gather_ms, gather_bytes, per_rank_ms = None, None, None, [total_ms]
    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x[0]) for x in allt]
        total_ms, e2e_s = max(per_rank_ms), max(float(x[1]) for x in allt)
        if full:   # the one collective of the path: gather the per-cloud pose records of a step, timed
            from articulated_pose_b200 import dist as adist
            rec = torch.cat([res[k].reshape(B, -1).double() for k in
                             ("single_R", "single_s", "single_t", "joint_R0", "joint_s0", "joint_t0", "joint_R1",
                              "joint_s1", "joint_t1", "joint_score")], 1).contiguous()
            rec = rec.repeat(C, 1)                     # records of all C device batches of a step
            adist.gather_records(rec, device=dev)      # warm (NCCL communicator set-up)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            g0.record()
            out = adist.gather_records(rec, device=dev)
            g1.record()
            torch.cuda.synchronize()
            gathered, gather_ms, gather_bytes = int(out.shape[0]), g0.elapsed_time(g1), int(out.numel() * 8)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    fl = stage_flops(pipe.net, B, N)
    tensor_peak = peaks["bf16_tflops_sustained"]
    n_fwd = 2 if (full and two_nets) else 1
    fwd_ms = sum(stage_ms.values())
    all_ms = {**stage_ms, **extra_ms}

    def tensor_roofline(dom):
        ach = fl[dom] / (stage_ms[dom] * 1e-3) / 1e12
        kern = {"sa1": "sa_lean_kernel<0,64,64,128>", "sa2": "sa_lean_kernel<128,128,128,256>"}.get(dom, "chain2_kernel<rows>")
        if args.precision != "f16x3":
            kern = "sa_kernel / fp_kernel (f32 CUDA cores)"
        return {"bound": "tensor", "kernel": "%s (%s, ANCSH net)" % (kern, dom), "achieved": ach, "peak": tensor_peak,
                "unit": "TFLOP/s", "frac": ach / tensor_peak, "traffic": None,
                "peak_source": peaks["source"] + " bf16 sustained (MEASURED_PEAKS.json)",
                "avg_launch_ms": stage_ms[dom], "flops_per_launch": fl[dom],
                "note": "algorithmic FLOPs (2*MAC, unpadded) of the stage for one device batch / its CUDA-event duration in a "
                        "serialized pass of this run; the fp16 hi/lo split issues 3 MMAs per algorithmic product"}
    tensor_stages = ("sa1", "sa2", "sa3", "fp1", "fp2", "fp3_heads")
    dom_t = max(tensor_stages, key=lambda k: stage_ms[k])
    roofline_tensor = tensor_roofline(dom_t)
    roofline_tensor["frac_by_stage"] = {k: round(fl[k] / (stage_ms[k] * 1e-3) / 1e12 / tensor_peak, 4) for k in tensor_stages}
    dom = max(all_ms, key=lambda k: all_ms[k])
    if dom == "pose_joint_score" and lm_tot["nfev"] > 0:
        flops = (lm_tot["nfev"] * (LM_FLOPS["cost"] + LM_FLOPS["finish"]) + lm_tot["njev"] * (LM_FLOPS["jacobian"] + LM_FLOPS["jac_phase"]) +
                 lm_tot["nlm"] * LM_FLOPS["lmpar_iter"]) / n_prof
        ach = flops / (all_ms[dom] * 1e-3) / 1e12
        roofline = {"bound": "fp64", "kernel": "joint_init + 3 x joint_lm_kernel + joint_model + joint_verify (pose stage joint_score)",
                    "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": None,
                    "peak_source": "DFMA peak measured in this process (ancsh_diag_fp64_fma, 8 chains x 8 blocks/SM)",
                    "avg_launch_ms": all_ms[dom], "flops_per_launch": flops,
                    "model": {"flops": LM_FLOPS, "evaluations": lm_tot["nfev"] / n_prof, "jacobians": lm_tot["njev"] / n_prof,
                              "lmpar_iterations": lm_tot["nlm"] / n_prof},
                    "note": "latency bound: MINPACK's lmder is a strictly sequential chain per solve (profiles/r02_lm_experiments.md); "
                            "the stage is kept narrow on purpose (lm_grid) so that the forwards of the overlapped batches keep the "
                            "other SMs: more lanes shorten this stage and lower the pipelined throughput"}
        import ctypes
        lm_thr, lm_blk = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(_lib.ancsh_pose_lm_shape(ctypes.byref(lm_thr), ctypes.byref(lm_blk)), "ancsh_pose_lm_shape")
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        roofline["lm_grid"] = {"threads_per_block": lm_thr.value, "max_blocks": lm_blk.value, "sms": sms}
        roofline["frac_of_occupied_sms"] = ach / (fp64_peak * min(lm_blk.value, sms) / sms)
    elif dom in fl:
        roofline = tensor_roofline(dom)
    else:
        roofline = dict(roofline_tensor, note="largest stage %s (%.3f ms) has no FLOP/byte model; tensor stage reported" % (dom, all_ms[dom]))
    roofline["stage_ms"] = {k: round(v, 4) for k, v in all_ms.items()}
    roofline["dominant_stage"] = dom
    roofline["forward_tflops"] = sum(fl.values()) / (fwd_ms * 1e-3) / 1e12

    clouds_per_step = B * C
    line = {"metric": METRIC, "value": world * clouds_per_step * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f16x3->f32" if args.precision == "f16x3" else "f32") + (" network / f64 pose" if full else ""),
            "data": "synthetic",
            "config": {"workload": workload_name(args), "baseline_config": args.config_index, "stages": args.stages,
                       "clouds_per_step_per_gpu": clouds_per_step, "device_batch": B, "device_batches_per_step": C,
                       "distinct_clouds_per_step": clouds_per_step, "ransac_seed": "one Philox key per device batch of the run",
                       "l2_flush_between_steps": True,
                       "weights": "seeded random trunk, linear seg/NOCS heads ridge-fitted on %d synthetic clouds "
                                  "(no checkpoint ships with the reference)" % CALIB_CLOUDS if full else "seeded random",
                       "forwards_per_cloud": n_fwd, "mean_part_sizes": part_hist, "joint_lm": lm_stats,
                       "serialized_ms_per_device_batch": round(serial_ms, 3),
                       "streams": "pose stage of device batch i (side stream) overlaps forwards of later batches (%d buffer slots)" % pipe.N_SLOTS if full else "single",
                       "wall_s_timed_region": round(t_wall, 4),
                       "host_enqueue_ms_first_step": round(1e3 * host_enqueue_s[0], 3),
                       "host_enqueue_ms_per_step_mean": round(1e3 * sum(host_enqueue_s) / args.steps, 3),
                       "ms_per_step_per_rank": {"min": round(min(per_rank_ms) / args.steps, 4), "max": round(max(per_rank_ms) / args.steps, 4)},
                       "all_gathered_records": gathered, "gather_ms": gather_ms, "gather_bytes": gather_bytes},
            "clocks": clocks,
            "e2e": {"value": world * clouds_per_step * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            # kernels of libancsh_b200.so launched inside the timed region, from the library's own launch counter
            # (ancsh_launch_count; memsets, copies and the L2 flush are not counted)
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_tensor": roofline_tensor}

    if not args.no_cpu_baseline:
        try:
            from oracle import pipeline_cpu
            fan = args.workload == "cpu64" and full
            n_s = args.cpu_sample or (B if fan else (6 if full else 16))
            Pc, jcc = P_host.reshape(-1, N, 3), jc_host.reshape(-1, N)
            cpu_reference_run(args, K, Pc, jcc, w_a, w_n if full else None, 1, 1)       # warm (pool fork, page-in)
            val, dt, cores, desc = cpu_reference_run(args, K, Pc, jcc, w_a, w_n if full else None, n_s, 1)
            pipeline_cpu.close_pool()
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                                    "seconds": round(dt, 2)}
        except Exception as e:  # the baseline is reporting only; never lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
