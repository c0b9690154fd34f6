"""End-to-end hot path: PointNet++ forward(s) -> per-part RANSAC -> joint-constrained solve, all on the device.

Mirrors what the reference does across two programs and a directory of h5 files:
  main.py --test  (Network.predict_and_save, lib/network.py:257-305)  ->  results/test_pred/<exp>/<basename>.h5
  pose_multi_process.py -> solver_ransac_nonlinear (evaluation/parallel_ancsh_pose.py:196-352)
The reference's solver reads NOCS + segmentation from the NPCS *baseline* experiment (USE_BASELINE = True,
parallel_ancsh_pose.py:197,232-236) and the joint axis from the ANCSH experiment, i.e. two network forwards per
cloud; `use_baseline` keeps that default.  Here the predictions never leave HBM between the stages.
"""
import numpy as np
import torch

from . import _lib
from .network import AncshNet
from .pose import PoseSolver, unpack_results


class AncshPipeline:
    # batches in flight in submit()/run_many(): the pose tails of several batches overlap later forwards.  Measured on B200
    # (256 clouds per batch, narrow joint-LM grid of csrc/pose.cu): 4 slots 33.1k clouds/s, 6 slots 36.8k, 8 slots 37.2k
    N_SLOTS = int(__import__('os').environ.get('ANCSH_SLOTS', '8'))

    def __init__(self, weights_ancsh, n_parts, weights_npcs=None, use_baseline=True, nsample=64, niter_single=10000,
                 niter_joint=200, inlier_th=0.1, seed=0, device="cuda:0", precision="f16x3"):
        self.device = torch.device(device)
        self.K = int(n_parts)
        self.use_baseline = bool(use_baseline and weights_npcs is not None)
        self.net = AncshNet(weights_ancsh, n_parts, mixed_pred=True, early_split_nocs=True, nsample=nsample, device=device,
                            precision=precision)
        self.net_npcs = None
        if self.use_baseline:
            self.net_npcs = AncshNet(weights_npcs, n_parts, mixed_pred=False, early_split_nocs=False, nsample=nsample,
                                     device=device, precision=precision)
        self.pose = PoseSolver(n_parts, niter_single, niter_joint, inlier_th, seed, device)
        self._buf = {}
        self._slots = {}
        self._pose_streams = {}

    def _buffers(self, B, N):
        key = (B, N)
        if key not in self._buf:
            d = {"pred": self.net.alloc_outputs(B, N), "pose": self.pose.alloc_outputs(B, N),
                 "P": torch.empty((B, N, 3), dtype=torch.float32, device=self.device),
                 "jc": torch.empty((B, N), dtype=torch.int32, device=self.device),
                 "hP": torch.empty((B, N, 3), dtype=torch.float32).pin_memory(),
                 "hjc": torch.empty((B, N), dtype=torch.int32).pin_memory()}
            if self.net_npcs is not None:
                d["pred_b"] = self._npcs_outputs(B, N)
            d["hpose"] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in d["pose"].items()}
            self._buf[key] = d
        return self._buf[key]

    def _npcs_outputs(self, B, N):
        """The pose stage reads only the segmentation and the NOCS of the baseline network (parallel_ancsh_pose.py:232-236):
        its joint branch is not requested, so the forward skips it (csrc/net.cu)."""
        full = self.net_npcs.alloc_outputs(B, N)
        return {k: full[k] for k in ("W", "nocs_per_point", "confi_per_point")}

    def run_device(self, P, joint_cls, net_events=None, net_b_events=None, pose_events=None, seed=None):
        """P (B,N,3) f32 CUDA, joint_cls (B,N) int32 CUDA -> dict of CUDA pose tensors (ancsh_pose_out_t).
        seed: Philox key of this batch's RANSAC draws (default: the pipeline's)."""
        B, N, _ = P.shape
        buf = self._buffers(B, N)
        pred = self.net.forward_device(P, buf["pred"], stage_events=net_events)
        src = pred
        if self.net_npcs is not None:
            src = self.net_npcs.forward_device(P, buf["pred_b"], stage_events=net_b_events, geometry_from=self.net)
        return self.pose.solve_device(P, src["nocs_per_point"], src["W"], pred["joint_axis_per_point"], joint_cls,
                                      out=buf["pose"], stage_events=pose_events, seed=seed)

    # ---- stream-pipelined submission: the pose stage of batch i overlaps the forwards of batch i+1 ----------
    def _slot(self, B, N, slot):
        key = (B, N, slot)
        if key not in self._slots:
            d = {"pred": self.net.alloc_outputs(B, N), "pose": self.pose.alloc_outputs(B, N), "fwd_done": torch.cuda.Event(),
                 "pose_done": torch.cuda.Event(), "used": False}
            if self.net_npcs is not None:
                d["pred_b"] = self._npcs_outputs(B, N)
            self._slots[key] = d
        return self._slots[key]

    def prepare(self, B, N):
        """Allocate everything submit() / run_many() create lazily for (B, N) batches -- the per-slot prediction and pose
        buffers, pose workspaces and side streams, the network workspaces, the pinned staging buffers -- so that the first
        pass over all N_SLOTS slots does not pay for cudaMalloc / cudaHostAlloc in the middle of a stream."""
        with torch.cuda.device(self.device):
            for slot in range(self.N_SLOTS):
                self._slot(B, N, slot)
                self._pose_stream(slot)
                self.pose._plan(B, N, slot)
            self.net._workspace(B, N)
            if self.net_npcs is not None:
                self.net_npcs._workspace(B, N)
            self._many_buffers(B, N)
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)

    def _pose_stream(self, slot):
        if slot not in self._pose_streams:
            # same priority as the caller's stream: measured on B200 (256 clouds/step) a high-priority pose stream gives
            # 24.8-24.9k clouds/s, equal priority 25.4k -- the pose kernels are latency-bound and fill the gaps anyway,
            # while pre-empting the block scheduler only stretches the forwards.  ANCSH_POSE_PRIORITY overrides (-1 = high).
            import os
            self._pose_streams[slot] = torch.cuda.Stream(device=self.device, priority=int(os.environ.get("ANCSH_POSE_PRIORITY", "0")))
        return self._pose_streams[slot]

    def _many_buffers(self, B, N):
        key = ("many", B, N)
        if key not in self._buf:
            self._buf[key] = [{"hP": torch.empty((B, N, 3), dtype=torch.float32).pin_memory(),
                               "hjc": torch.empty((B, N), dtype=torch.int32).pin_memory(),
                               "P": torch.empty((B, N, 3), dtype=torch.float32, device=self.device),
                               "jc": torch.empty((B, N), dtype=torch.int32, device=self.device),
                               "hpose": {k: torch.empty(v.shape, dtype=v.dtype).pin_memory()
                                         for k, v in self.pose.alloc_outputs(B, N).items()},
                               "copied": torch.cuda.Event()} for _ in range(self.N_SLOTS)]
        return self._buf[key]

    def submit(self, P, joint_cls, slot=0, net_events=None, net_b_events=None, pose_events=None, seed=None):
        """Asynchronous run_device: the forwards are enqueued on the current stream, the pose stage on an
        internal side stream (one per slot) that waits for them; buffers are per `slot` (cycle through
        N_SLOTS slots).  The slow tail
        of the joint LM solves then overlaps the next batch's forwards.  Returns the slot's pose tensors; call
        `join()` (or wait on the returned dict's "done" event) before reading them."""
        B, N, _ = P.shape
        sl = self._slot(B, N, slot)
        main = torch.cuda.current_stream()
        ps = self._pose_stream(slot)
        if sl["used"]:
            main.wait_event(sl["pose_done"])          # the slot's prediction buffers are still being read
        pred = self.net.forward_device(P, sl["pred"], stage_events=net_events)
        src = pred
        if self.net_npcs is not None:
            src = self.net_npcs.forward_device(P, sl["pred_b"], stage_events=net_b_events, geometry_from=self.net)
        sl["fwd_done"].record(main)
        with torch.cuda.stream(ps):
            ps.wait_event(sl["fwd_done"])
            out = self.pose.solve_device(P, src["nocs_per_point"], src["W"], pred["joint_axis_per_point"], joint_cls,
                                         out=sl["pose"], stage_events=pose_events, ws_slot=slot, seed=seed)
            sl["pose_done"].record(ps)
        sl["used"] = True
        return out

    def join(self):
        """Make the current stream wait for every submitted pose stage."""
        main = torch.cuda.current_stream()
        for sl in self._slots.values():
            if sl["used"]:
                main.wait_event(sl["pose_done"])

    def run_many(self, batches, unpack=False, seeds=None):
        """Host API for a stream of batches [(P, joint_cls), ...] (all the same shape): pinned H2D, forwards and
        pose stages pipelined over N_SLOTS slots, D2H of the pose records.  Returns one result per batch.
        seeds: optional Philox key per batch."""
        return self.run_many_begin(batches, unpack=unpack, seeds=seeds)()

    def run_many_begin(self, batches, unpack=False, seeds=None):
        """run_many in two halves: enqueues every batch (collecting older results only where a slot has to be reused) and
        returns `finish`, which waits for the rest and returns the per-batch results.  Work enqueued by the caller in
        between -- e.g. another category's pipeline (stream.MixedStream) -- overlaps this stream's pose tails."""
        if not batches:
            return lambda: []
        B, N, _ = batches[0][0].shape
        bufs = self._many_buffers(B, N)
        results = [None] * len(batches)
        pending = [None] * self.N_SLOTS
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        copy_stream = self._copy_stream

        def drain(slot):
            if pending[slot] is None:
                return
            i, out = pending[slot]
            # D2H on its own stream behind the slot's pose stage only: on the main stream the copy would queue behind
            # the forwards of the newer batches and the host would fall a whole pipeline depth behind the device
            copy_stream.wait_event(self._slot(B, N, slot)["pose_done"])
            with torch.cuda.stream(copy_stream):
                for k, v in out.items():
                    bufs[slot]["hpose"][k].copy_(v, non_blocking=True)
                bufs[slot]["copied"].record(copy_stream)
            bufs[slot]["copied"].synchronize()
            h = {k: v.numpy().copy() for k, v in bufs[slot]["hpose"].items()}
            results[i] = unpack_results(h, self.K) if unpack else h
            pending[slot] = None

        with torch.cuda.device(self.device):
            for i, (P, jc) in enumerate(batches):
                slot = i % self.N_SLOTS
                drain(slot)                              # results of the batch that used this slot last
                bufs[slot]["hP"].numpy()[...] = P
                bufs[slot]["hjc"].numpy()[...] = jc
                bufs[slot]["P"].copy_(bufs[slot]["hP"], non_blocking=True)
                bufs[slot]["jc"].copy_(bufs[slot]["hjc"], non_blocking=True)
                pending[slot] = (i, self.submit(bufs[slot]["P"], bufs[slot]["jc"], slot=slot,
                                                seed=None if seeds is None else seeds[i]))

        def finish():
            with torch.cuda.device(self.device):
                for k in range(self.N_SLOTS):              # oldest first
                    drain((len(batches) + k) % self.N_SLOTS)
            return results
        return finish

    def run(self, P, joint_cls, unpack=True):
        """Host arrays in, host results out: pinned H2D of the clouds, all stages on the device, D2H of the poses
        (rotations, scales, translations, scores, inlier masks)."""
        P = np.ascontiguousarray(P, np.float32)
        B, N, _ = P.shape
        buf = self._buffers(B, N)
        buf["hP"].numpy()[...] = P
        buf["hjc"].numpy()[...] = joint_cls
        with torch.cuda.device(self.device):
            buf["P"].copy_(buf["hP"], non_blocking=True)
            buf["jc"].copy_(buf["hjc"], non_blocking=True)
            out = self.run_device(buf["P"], buf["jc"])
            for k, v in out.items():
                buf["hpose"][k].copy_(v, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h = {k: v.numpy() for k, v in buf["hpose"].items()}
        return unpack_results(h, self.K) if unpack else h
