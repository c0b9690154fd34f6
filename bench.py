#!/usr/bin/env python
"""bench.py -- clouds/sec of the ANCSH hot path on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic clouds per rank (weak scaling: every rank
owns its own batch; clouds are independent; one all-gather of the pose records at the end).
Workload (BASELINE.json configs[2], the configuration the metric is quoted on): eyeglasses, N=1024, nsample=32,
full pipeline = ANCSH forward + NPCS-baseline forward (the reference's solver reads NOCS/mask from the baseline
experiment, parallel_ancsh_pose.py:197,232-236) + 3 single RANSACs (500 hyp) + 2 joint RANSACs (200 hyp, LM).
Prints ONE JSON line on rank 0.

  value     device-timed clouds/s, inputs resident in HBM (CUDA events per step on the launch stream, L2 flushed
            between steps, summed over the K steps, max over ranks)
  e2e       the same pass through the public host API (AncshPipeline.run) with HOST buffers: pinned H2D of the
            clouds and D2H of the pose results inside the timed region
  roofline  the dominant kernel (grouped-MLP set-abstraction stage) from per-stage CUDA events recorded during
            the timed steps, against the measured bf16 tensor peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference   the CPU oracle (oracle/: line-for-line restatement of the reference's network
            ops -- TF1 cannot be installed and the reference has no CPU kernels for FPS / ball query -- and of its
            numpy/scipy pose code) on this box's host cores
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=48)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clouds per rank per step")
    ap.add_argument("--category", default="eyeglasses")
    ap.add_argument("--nsample", type=int, default=32, help="BASELINE config K=32 (reference default: 64)")
    ap.add_argument("--hyp", type=int, default=500, help="RANSAC hypotheses per part (BASELINE config: 500; reference 10000)")
    ap.add_argument("--joint-hyp", type=int, default=200)
    ap.add_argument("--stages", default="full", choices=["forward", "full"])
    ap.add_argument("--no-baseline-net", action="store_true", help="USE_BASELINE=False: one forward per cloud")
    ap.add_argument("--cpu-sample", type=int, default=0, help="clouds in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "f32"],
                    help="f16x3: grouped MLPs on tcgen05 (fp16 hi/lo split, f32 accumulate); f32: CUDA-core FMA kernels")
    return ap.parse_args()


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
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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


def workload_name(args):
    cat = args.category
    n = 1024 if cat == "eyeglasses" else 2048
    if args.stages == "forward":
        return "%s ANCSH, N=%d nsample=%d, batch=%d clouds/GPU, PN++ forward + heads" % (cat, n, args.nsample, args.batch)
    return ("%s ANCSH, N=%d nsample=%d, batch=%d clouds/GPU, full pipeline: %s + RANSAC(%d hyp/part) + "
            "joint solve(%d hyp/joint, LM)") % (cat, n, args.nsample, args.batch,
                                                "ANCSH forward" if args.no_baseline_net else "ANCSH + NPCS-baseline forwards",
                                                args.hyp, args.joint_hyp)


def synthetic_weight_sets(K, args, feature_fn=None):
    """Seeded random trunks; when feature_fn(net_kind, weights, P) -> (M,128) is given, the linear segmentation /
    NOCS heads are fitted on CALIB_CLOUDS synthetic clouds (weights.fit_heads) so that the part partition -- and
    with it the pose-stage workload -- is realistic."""
    from articulated_pose_b200 import synthetic, weights
    w_a = weights.synthetic_weights(K, True, True, seed=7)
    w_n = None if args.no_baseline_net else weights.synthetic_weights(K, False, False, seed=8)
    if feature_fn is not None:
        Pc, cc = synthetic.make_batch(range(900000, 900000 + CALIB_CLOUDS), args.category)
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
    t0 = time.perf_counter()
    done = 0
    for s in range(steps):
        lo = (s * n_clouds) % max(1, P.shape[0] - n_clouds + 1)
        done += pipeline_cpu.run_clouds(P[lo:lo + n_clouds], jc[lo:lo + n_clouds], w_a, w_n, K, args.nsample, 0.1, args.hyp,
                                        args.joint_hyp, stages=args.stages, seed=s)
    dt = time.perf_counter() - t0
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
    n_s = args.cpu_sample or 1
    P, clouds = synthetic.make_batch(range(max(n_s, 4)), args.category)
    jc = np.stack([c["joint_cls_gt"] for c in clouds])
    for _ in range(min(args.warmup, 1)):
        cpu_reference_run(args, K, P, jc, w_a, w_n, 1, 1)
    val, dt, cores, desc = cpu_reference_run(args, K, P, jc, w_a, w_n, n_s, args.steps)
    pipeline_cpu.close_pool()
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 network / f64 pose", "data": "synthetic",
            "config": {"workload": workload_name(args), "stages": args.stages, "sample": desc},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
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
    B = args.batch
    full = args.stages == "full"

    def feats(kind, w, Pc):
        n = AncshNet(w, K, mixed_pred=(kind == "ancsh"), early_split_nocs=(kind == "ancsh"), nsample=args.nsample, device=dev,
                     precision=args.precision)
        return n.features(Pc)
    w_a, w_n = synthetic_weight_sets(K, args, feats if full else None)
    pipe = AncshPipeline(w_a, K, weights_npcs=w_n if full else None, use_baseline=full and not args.no_baseline_net,
                         nsample=args.nsample, niter_single=args.hyp, niter_joint=args.joint_hyp, seed=1234 + rank,
                         device=dev, precision=args.precision)
    P_host, clouds = synthetic.make_batch(range(rank * B, rank * B + B), args.category)
    jc_host = np.stack([c["joint_cls_gt"] for c in clouds]).astype(np.int32)
    N = P_host.shape[1]
    P_dev = torch.from_numpy(P_host).to(dev)
    jc_dev = torch.from_numpy(jc_host).to(dev)
    out_fwd = pipe.net.alloc_outputs(B, N)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    two_nets = pipe.net_npcs is not None

    nst, npst = len(_lib.NET_STAGES), len(_lib.POSE_STAGES)

    def step(ev=None):
        if full:
            return pipe.run_device(P_dev, jc_dev, net_events=ev[0] if ev else None, net_b_events=ev[1] if ev else None,
                                   pose_events=ev[2] if ev else None)
        return pipe.net.forward_device(P_dev, out_fwd, stage_events=ev[0] if ev else None)

    for _ in range(max(args.warmup, 3)):
        res = step()
    if full:
        for i in range(pipe.N_SLOTS):      # warm the pipelined path too (slot buffers, side streams)
            pipe.submit(P_dev, jc_dev, slot=i)
        pipe.join()
    torch.cuda.synchronize()
    part_hist = res["part_count"].float().mean(0).tolist() if full else None

    # ---- serialized profiling pass: per-stage CUDA events (stage durations without cross-batch overlap) --------
    n_prof = min(5, args.steps)
    evs = [(_lib.EventList(nst + 1), _lib.EventList(nst + 1), _lib.EventList(npst + 1)) for _ in range(n_prof)]
    for i in range(n_prof):
        flush.zero_()
        step(evs[i])
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
            extra_ms["npcs_forward"] = sum(e[1].elapsed_ms(0, nst) for e in evs) / n_prof
        for i, nm in enumerate(_lib.POSE_STAGES):
            extra_ms["pose_" + nm] = sum(e[2].elapsed_ms(i, i + 1) for e in evs) / n_prof

    # ---- timed region: EXACTLY K steps.  Full pipeline: batches are submitted over two buffer slots, the pose stage
    # of step i runs on a side stream and overlaps the forwards of step i+1 (production configuration). ----------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (256 MB memset, ~0.1 ms, counted)
        if full:
            res = pipe.submit(P_dev, jc_dev, slot=i % pipe.N_SLOTS)
        else:
            step()
    if full:
        pipe.join()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    total_ms = ev0.elapsed_time(ev1)
    nf = pipe.pose.intermediates()["joint_nfev"].float() if full else None
    lm_stats = None
    if full:
        q = torch.quantile(nf.flatten(), torch.tensor([0.5, 0.99, 1.0], device=nf.device)).tolist()
        lm_stats = {"nfev_p50": q[0], "nfev_p99": q[1], "nfev_max": q[2]}

    # ---- end to end through the public host API (host buffers, H2D + D2H inside the timed region) ----
    def e2e_call():
        if full:
            return pipe.run(P_host, jc_host, unpack=False)
        return pipe.net.forward(P_host, copy=False)
    for _ in range(2):
        e2e_call()
    if full:   # warm-up of the call that is timed below: run_many pins its staging buffers and creates its copy stream on first use
        pipe.run_many([(P_host, jc_host)] * max(args.warmup, 3))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if full:
        rr = pipe.run_many([(P_host, jc_host)] * args.steps)
        r = rr[-1]
    else:
        for _ in range(args.steps):
            r = e2e_call()
    e2e_s = time.perf_counter() - t0
    h2d = P_host.nbytes + (jc_host.nbytes if full else 0)
    d2h = sum(v.nbytes for v in r.values())

    gathered = None
    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])
        if full:   # the one collective of the path: gather the per-cloud pose records
            from articulated_pose_b200 import dist as adist
            rec = torch.cat([res[k].reshape(B, -1).double() for k in
                             ("single_R", "single_s", "single_t", "joint_R0", "joint_s0", "joint_t0", "joint_R1",
                              "joint_s1", "joint_t1", "joint_score")], 1).contiguous()
            gathered = int(adist.gather_records(rec, device=dev).shape[0])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    fl = stage_flops(pipe.net, B, N)
    dom = max(("sa1", "sa2"), key=lambda k: stage_ms[k])
    ach = fl[dom] / (stage_ms[dom] * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    n_fwd = 2 if (full and two_nets) else 1
    fwd_ms = sum(stage_ms.values())
    # dram__bytes_read.sum + dram__bytes_write.sum of that launch from the committed ncu capture of this same command
    # (profiles/r01u_chain_traffic.csv: batch 256, nsample 32, eyeglasses); null for any other configuration
    NCU_TRAFFIC = {"sa1": 19673088 + 18748160, "sa2": 24419584 + 352256}
    ncu_cfg = args.precision == "f16x3" and B == 256 and N == 1024 and args.nsample == 32 and args.category == "eyeglasses"
    roofline = {"bound": "tensor", "kernel": ("chain2_kernel<SA>" if args.precision == "f16x3" else "sa_kernel<128>") + " (%s, ANCSH net)" % dom,
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": NCU_TRAFFIC[dom] if ncu_cfg else None,
                "traffic_source": "profiles/r01u_chain_traffic.csv (ncu dram bytes of this launch, same command)" if ncu_cfg else None,
                "peak_source": peaks["source"] + " bf16 sustained; kernel duration from per-stage CUDA events of a serialized pass in the same run",
                "avg_launch_ms": stage_ms[dom], "flops_per_launch": fl[dom],
                "stage_ms": {k: round(v, 4) for k, v in {**stage_ms, **extra_ms}.items()},
                "forward_tflops": sum(fl.values()) / (fwd_ms * 1e-3) / 1e12}

    line = {"metric": METRIC, "value": world * B * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f16x3->f32" if args.precision == "f16x3" else "f32") + (" network / f64 pose" if full else ""),
            "data": "synthetic",
            "config": {"workload": workload_name(args), "stages": args.stages, "l2_flush_between_steps": True,
                       "weights": "seeded random trunk, linear seg/NOCS heads ridge-fitted on %d synthetic clouds "
                                  "(no checkpoint ships with the reference)" % CALIB_CLOUDS if full else "seeded random",
                       "forwards_per_cloud": n_fwd, "mean_part_sizes": part_hist, "joint_lm": lm_stats,
                       "serialized_ms_per_step": round(serial_ms, 3),
                       "streams": "pose stage of step i (side stream) overlaps forwards of later steps (%d buffer slots)" % pipe.N_SLOTS if full else "single",
                       "wall_s_timed_region": round(t_wall, 4), "all_gathered_records": gathered},
            "clocks": clocks,
            "e2e": {"value": world * B * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            # kernels of this library per step, counted in profiles/r01w_launches.csv: 15 per forward (2 fps, 2 ball query,
            # 6 chain2, 1 gemm_tc, cloud_bias, 2 fp_interp, heads_act), 11 for the second network (no fps / ball query,
            # fp_blend instead of fp_interp), 10 per pose stage (partition, single score + refit, joint init, 3 LM phases,
            # model, verify, refit); memsets and the L2 flush are not counted
            "gpu_launches": (15 + (11 if n_fwd == 2 else 0) + (10 if full else 0)) * args.steps,
            "roofline": roofline}

    if not args.no_cpu_baseline:
        try:
            from oracle import pipeline_cpu
            n_s = args.cpu_sample or (6 if full else 16)
            cpu_reference_run(args, K, P_host, jc_host, w_a, w_n if full else None, 1, 1)       # warm (pool fork, page-in)
            val, dt, cores, desc = cpu_reference_run(args, K, P_host, jc_host, w_a, w_n if full else None, n_s, 1)
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
