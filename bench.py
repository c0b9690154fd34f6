#!/usr/bin/env python
"""bench.py -- clouds/sec of the ANCSH hot path on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic clouds per rank (weak scaling: every rank
owns its own batch; clouds are independent).  Prints ONE JSON line on rank 0.

  value     device-timed clouds/s, inputs resident in HBM (CUDA events per step on the launch stream, L2 flushed
            between steps, summed over the K steps, max over ranks)
  e2e       the same pass through the public host API (AncshNet.forward / solve) with HOST buffers: pinned H2D of
            the clouds and D2H of the results inside the timed region
  roofline  the dominant kernel (grouped-MLP set-abstraction stage) from per-stage CUDA events recorded during
            the timed steps, against the measured bf16 tensor peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference   the CPU oracle (oracle/, a line-for-line restatement of the reference; TF1
            cannot be installed and the reference has no CPU kernels for FPS / ball query) on this box's host cores
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clouds per rank per step")
    ap.add_argument("--category", default="eyeglasses")
    ap.add_argument("--nsample", type=int, default=32, help="BASELINE config K=32 (reference default: 64)")
    ap.add_argument("--hyp", type=int, default=500, help="RANSAC hypotheses per part (BASELINE config: 500)")
    ap.add_argument("--joint-hyp", type=int, default=200)
    ap.add_argument("--stages", default="auto", help="forward | full | auto")
    ap.add_argument("--cpu-sample", type=int, default=0, help="clouds in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (oracle/), all host
    threads.  Rank 0 only."""
    if rank != 0:
        return
    from articulated_pose_b200 import synthetic, weights
    from oracle import pnpp
    K = synthetic.CATEGORIES[args.category]["boxes"].__len__()
    w = weights.synthetic_weights(K)
    cores = os.cpu_count() or 1
    pnpp.set_threads(cores)
    n_s = args.cpu_sample or 4
    P, _ = synthetic.make_batch(range(n_s), args.category)
    for _ in range(max(1, min(args.warmup, 1))):
        pnpp.forward(P[:1], w, K, nsample=args.nsample)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pnpp.forward(P, w, K, nsample=args.nsample)
    dt = time.perf_counter() - t0
    val = n_s * args.steps / dt
    sample = "%d clouds/step x %d steps, network forward only (pose stage port pending)" % (n_s, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, "forward"), "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(args, stages):
    cat = args.category
    n = 1024 if cat == "eyeglasses" else 2048
    if stages == "forward":
        return "%s ANCSH, N=%d nsample=%d, batch=%d clouds/GPU, PN++ forward + heads" % (cat, n, args.nsample, args.batch)
    return "%s ANCSH, N=%d nsample=%d, batch=%d clouds/GPU, full pipeline: PN++ forward + RANSAC(%d hyp/part) + joint solve(%d hyp/joint)" % (
        cat, n, args.nsample, args.batch, args.hyp, args.joint_hyp)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from articulated_pose_b200 import _lib, synthetic, weights
    from articulated_pose_b200.network import AncshNet

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    K = len(synthetic.CATEGORIES[args.category]["boxes"])
    B = args.batch
    w = weights.synthetic_weights(K)
    net = AncshNet(w, K, nsample=args.nsample, device=dev)
    P_host, clouds = synthetic.make_batch(range(rank * B, rank * B + B), args.category)
    N = P_host.shape[1]
    P_dev = torch.from_numpy(P_host).to(dev)
    out = net.alloc_outputs(B, N)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stages = "forward"

    def step(ev=None):
        net.forward_device(P_dev, out, stage_events=ev)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    nst = len(_lib.NET_STAGES)
    evs = [_lib.EventList(nst + 1) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.zero_()                       # L2 flush between timed iterations (not inside the event pair)
        step(evs[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    step_ms = [e.elapsed_ms(0, nst) for e in evs]
    total_ms = sum(step_ms)
    stage_ms = {nm: sum(e.elapsed_ms(i, i + 1) for e in evs) / args.steps for i, nm in enumerate(_lib.NET_STAGES)}

    # ---- end to end through the public host API (host buffers, H2D + D2H inside the timed region) ----
    for _ in range(2):
        net.forward(P_host, copy=False)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = net.forward(P_host, copy=False)
    e2e_s = time.perf_counter() - t0
    h2d = P_host.nbytes
    d2h = sum(v.nbytes for v in res.values())

    if world > 1:
        t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    fl = stage_flops(net, B, N)
    dom = max(fl, key=lambda k: stage_ms[k])
    ach = fl[dom] / (stage_ms[dom] * 1e-3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    roofline = {"bound": "tensor", "kernel": {"sa1": "sa_kernel<128> (layer1)", "sa2": "sa_kernel<128> (layer2)"}.get(dom, dom),
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "peak_source": peaks["source"] + " bf16 sustained (kernel timed inside the step loop)",
                "avg_launch_ms": stage_ms[dom], "flops_per_launch": fl[dom],
                "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
                "whole_forward_tflops": sum(fl.values()) / (total_ms / args.steps * 1e-3) / 1e12}

    line = {"metric": METRIC, "value": world * B * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args, stages), "stages": stages, "l2_flush_between_steps": True,
                       "weights": "seeded random (no checkpoint ships with the reference)",
                       "wall_s_timed_region": round(t_wall, 4)},
            "clocks": clocks,
            "e2e": {"value": world * B * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": 11 * args.steps,
            "roofline": roofline}

    if not args.no_cpu_baseline:
        try:
            from oracle import pnpp
            cores = os.cpu_count() or 1
            pnpp.set_threads(cores)
            n_s = args.cpu_sample or 8
            pnpp.forward(P_host[:1], w, K, nsample=args.nsample)
            t0 = time.perf_counter()
            pnpp.forward(P_host[:n_s], w, K, nsample=args.nsample)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n_s / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d clouds of the same batch, network forward (oracle/pnpp.py, OpenMP over rows)" % n_s}
        except Exception as e:  # the baseline is reporting only; never lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
