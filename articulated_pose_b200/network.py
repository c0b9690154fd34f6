"""AncshNet -- host-side mirror of the reference's inference call

    pred_result = sess.run(self.pred_dict, feed_dict={P: batch})          (lib/network.py:292)

for the graph of lib/architecture.py:86-161 (get_per_point_model_new) over
pointnet_plusplus/architectures.py:56-95 (build_pointnet2_shared).  `forward(P)` returns a dict with exactly
the keys / shapes / dtypes of `pred_dict` (lib/architecture.py:141-159) so that the reference's unmodified
`prediction_io.save_batch_nn` (lib/prediction_io.py:65-95) can consume it.

PyTorch tensors are device-memory containers only; all compute is in libancsh_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .weights import flatten_packed, flatten_tc_images, pack_network

_PRED_WIDTH = {
    "W": lambda K: K, "nocs_per_point": lambda K: 3 * K, "confi_per_point": lambda K: 1,
    "heatmap_per_point": lambda K: 1, "unitvec_per_point": lambda K: 3, "joint_axis_per_point": lambda K: 3,
    "index_per_point": lambda K: 3, "gocs_per_point": lambda K: 3 * K, "global_scale": lambda K: K,
    "global_translation": lambda K: 3 * K,
}
_MIXED_ONLY = ("gocs_per_point", "global_scale", "global_translation")

_WS_VIEWS = {  # name -> (dtype, shape builder)
    "fps_idx1": (torch.int32, lambda s, B: (B, s.npoint1)),
    "l1_xyz": (torch.float32, lambda s, B: (B, s.npoint1, 3)),
    "fps_idx2": (torch.int32, lambda s, B: (B, s.npoint2)),
    "l2_xyz": (torch.float32, lambda s, B: (B, s.npoint2, 3)),
    "ball_idx1": (torch.int32, lambda s, B: (B, s.npoint1, s.nsample1)),
    "ball_cnt1": (torch.int32, lambda s, B: (B, s.npoint1)),
    "ball_idx2": (torch.int32, lambda s, B: (B, s.npoint2, s.nsample2)),
    "ball_cnt2": (torch.int32, lambda s, B: (B, s.npoint2)),
    "l1_points": (torch.float32, lambda s, B: (B, s.npoint1, 128)),
    "l2_points": (torch.float32, lambda s, B: (B, s.npoint2, 256)),
    "l3_points": (torch.float32, lambda s, B: (B, 1024)),
    "l2_points_fp": (torch.float32, lambda s, B: (B, s.npoint2, 256)),
    "l1_points_fp": (torch.float32, lambda s, B: (B, s.npoint1, 128)),
}


class AncshNet:
    def __init__(self, weights, n_parts, mixed_pred=True, early_split_nocs=True, nsample=64, npoint1=512, npoint2=128,
                 radius1=0.2, radius2=0.4, device="cuda:0", prefix="SPFN", precision="f16x3", packer="python"):
        """weights: dict TF-variable-name -> ndarray (see weights.variable_shapes).
        ANCSH (exp 3.9): mixed_pred=True, early_split_nocs=True (main.py:42-49);
        NPCS baseline (exp 3.91): mixed_pred=False, early_split_nocs=False.
        precision: "f16x3" = grouped MLPs on the tcgen05 tensor cores with the fp16 hi/lo split (3 products, f32-class accuracy,
        default); "f32" = exact f32 FMA kernels on the CUDA cores.  The f16x3 path needs |activation| < 65504 (larger
        values are clipped, csrc/tc_common.cuh); weights of any magnitude are handled by per-layer power-of-two scales.
        packer: "python" = weights.py builds the device buffers; "c" = the library's own import (ancsh_weights_pack +
        ancsh_net_create, what a non-Python host uses) -- same buffers bit for bit (tests/test_weights_c_cpu.py)."""
        if precision not in ("f16x3", "f32"):
            raise ValueError("precision must be 'f16x3' or 'f32'")
        self.precision = precision
        if not torch.cuda.is_available():
            raise RuntimeError("AncshNet needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.n_parts, self.mixed_pred = int(n_parts), bool(mixed_pred)
        self.npoint1, self.npoint2 = int(npoint1), int(npoint2)
        self.nsample1 = self.nsample2 = int(nsample)
        self.layers = pack_network(weights, n_parts, mixed_pred, early_split_nocs, prefix)
        self._c_handles = None
        if packer == "c":
            if (npoint1, npoint2, float(radius1), float(radius2)) != (512, 128, 0.2, 0.4):
                raise ValueError("the C packer builds the reference's level settings only")
            flatten_tc_images(self.layers)                      # sets tc_exp on the metadata copies
            self._net = self._create_in_c(weights, early_split_nocs, prefix, int(nsample), precision == "f16x3")
            self._ws, self._host, self.last_workspace = {}, {}, None
            return
        if packer != "python":
            raise ValueError("packer must be 'python' or 'c'")
        flat, offs = flatten_packed(self.layers)
        self._wbuf = torch.from_numpy(flat).to(self.device)
        base = self._wbuf.data_ptr()
        tc_flat, tc_offs = flatten_tc_images(self.layers)
        self._tcbuf = torch.from_numpy(tc_flat.view(np.int16)).to(self.device)
        tc_base = self._tcbuf.data_ptr()

        net = _lib.Net()
        net.use_tensor_cores = int(precision == "f16x3")
        net.n_parts, net.mixed_pred = self.n_parts, int(self.mixed_pred)
        net.npoint1, net.nsample1, net.radius1 = self.npoint1, self.nsample1, float(radius1)
        net.npoint2, net.nsample2, net.radius2 = self.npoint2, self.nsample2, float(radius2)
        for slot, pl in self.layers.items():
            wo, bo = offs[slot]
            if "[" in slot:
                name, i = slot[:-1].split("[")
                dst = getattr(net, name)[int(i)]
            else:
                dst = getattr(net, slot)
            dst.W, dst.b = base + 4 * wo, base + 4 * bo
            dst.W_tc = tc_base + 2 * tc_offs[slot] if slot in tc_offs else None
            dst.cin, dst.cout, dst.cin_pad, dst.cout_pad, dst.relu = pl.cin, pl.cout, pl.cin_pad, pl.cout_pad, pl.relu
            dst.tc_descale = float(np.ldexp(1.0, -pl.tc_exp))
        net.tc_bias_step = 1
        l0 = self.layers["sa1[0]"]
        self._sa1_conv0 = np.ascontiguousarray(np.concatenate([l0.W[:3, :64].ravel(), l0.b[:64]]).astype(np.float32)) \
            if l0.cout_pad == 64 and l0.cin == 3 else None
        net.sa1_conv0_host = self._sa1_conv0.ctypes.data if self._sa1_conv0 is not None else None
        self._net = net
        self._ws = {}       # (B,N) -> (workspace tensor, layout)
        self._host = {}     # (B,N) -> pinned staging buffers
        self.last_workspace = None

    def _create_in_c(self, weights, early_split_nocs, prefix, nsample, use_tc):
        names = list(weights)
        arrs = [np.ascontiguousarray(weights[k], np.float32) for k in names]
        n = len(names)
        c_names = (ctypes.c_char_p * n)(*[k.encode() for k in names])
        c_data = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
        c_cnt = (ctypes.c_size_t * n)(*[a.size for a in arrs])
        packed, handle = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(_lib.ancsh_weights_pack(n, c_names, c_data, c_cnt, self.n_parts, int(self.mixed_pred), int(early_split_nocs),
                                           prefix.encode(), ctypes.byref(packed)), "ancsh_weights_pack")
        try:
            with torch.cuda.device(self.device):
                _lib.check(_lib.ancsh_net_create(packed, nsample, int(use_tc), ctypes.byref(handle)), "ancsh_net_create")
        finally:
            _lib.ancsh_packed_destroy(packed)
        self._c_handles = handle
        return ctypes.cast(_lib.ancsh_net_get(handle), ctypes.POINTER(_lib.Net)).contents

    def __del__(self):
        h = getattr(self, "_c_handles", None)
        if h:
            _lib.ancsh_net_destroy(h)
            self._c_handles = None

    # ------------------------------------------------------------------------------------------
    def plan(self, B, N):
        lay = _lib.WsLayout()
        _lib.check(_lib.ancsh_net_plan(ctypes.byref(self._net), B, N, ctypes.byref(lay)), "ancsh_net_plan")
        return lay

    def _workspace(self, B, N):
        key = (B, N)
        if key not in self._ws:
            lay = self.plan(B, N)
            self._ws[key] = (torch.empty(lay.total_bytes, dtype=torch.uint8, device=self.device), lay)
        return self._ws[key]

    def alloc_outputs(self, B, N):
        K = self.n_parts
        out = {}
        for k in _lib.PRED_FIELDS:
            if k in _MIXED_ONLY and not self.mixed_pred:
                continue
            out[k] = torch.empty((B, N, _PRED_WIDTH[k](K)), dtype=torch.float32, device=self.device)
        return out

    def forward_device(self, P, out=None, stage_events=None, net_out=None, geometry_from=None):
        """P: CUDA f32 (B,N,3).  Launches on torch's current stream; returns dict of CUDA tensors.
        stage_events: optional _lib.EventList(len(_lib.NET_STAGES)+1) recorded around each stage.
        geometry_from: another AncshNet whose forward over the same P was enqueued before on the same stream: its
        FPS / ball-query results are reused instead of being recomputed (ancsh_net_forward_shared)."""
        if P.dtype != torch.float32 or P.dim() != 3 or P.shape[2] != 3 or not P.is_cuda:
            raise ValueError("P must be a CUDA float32 tensor of shape (B,N,3)")
        P = P.contiguous()
        B, N, _ = P.shape
        ws, lay = self._workspace(B, N)
        if out is None:
            out = self.alloc_outputs(B, N)
        pred = _lib.Pred()
        for k in _lib.PRED_FIELDS:
            setattr(pred, k, out[k].data_ptr() if k in out else None)
        pred.net = net_out.data_ptr() if net_out is not None else None     # optional (B,N,128) trunk feature
        ev = stage_events.arr if stage_events is not None else None
        if geometry_from is not None:
            if geometry_from.last_workspace is None:
                raise ValueError("geometry_from has not run a forward yet")
            gws, _, gB, gN, gP = geometry_from.last_workspace
            if (gB, gN) != (B, N) or gP != P.data_ptr():
                raise ValueError("geometry_from's last forward ran on a different batch (shape %r / another tensor): its FPS, "
                                 "ball-query and three_nn tables do not belong to this P" % ((gB, gN),))
            rc = _lib.ancsh_net_forward_shared(ctypes.byref(self._net), B, N, P.data_ptr(), ws.data_ptr(), lay.total_bytes,
                                               ctypes.byref(geometry_from._net), gws.data_ptr(), ctypes.byref(pred), ev,
                                               torch.cuda.current_stream().cuda_stream)
        else:
            rc = _lib.ancsh_net_forward(ctypes.byref(self._net), B, N, P.data_ptr(), ws.data_ptr(), lay.total_bytes,
                                        ctypes.byref(pred), ev, torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "ancsh_net_forward")
        self.last_workspace = (ws, lay, B, N, P.data_ptr())
        return out

    def forward(self, P, copy=True):
        """P: host ndarray (B,N,3) -> dict of host ndarrays (f32), like sess.run(pred_dict).
        copy=False returns views of the pinned staging buffers (overwritten by the next call)."""
        P = np.ascontiguousarray(P, dtype=np.float32)
        B, N, _ = P.shape
        key = (B, N)
        if key not in self._host:
            hin = torch.empty((B, N, 3), dtype=torch.float32).pin_memory()
            dev_in = torch.empty((B, N, 3), dtype=torch.float32, device=self.device)
            dev_out = self.alloc_outputs(B, N)
            hout = {k: torch.empty(v.shape, dtype=torch.float32).pin_memory() for k, v in dev_out.items()}
            self._host[key] = (hin, dev_in, dev_out, hout)
        hin, dev_in, dev_out, hout = self._host[key]
        hin.numpy()[...] = P
        with torch.cuda.device(self.device):
            dev_in.copy_(hin, non_blocking=True)
            self.forward_device(dev_in, dev_out)
            for k, v in dev_out.items():
                hout[k].copy_(v, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return {k: (v.numpy().copy() if copy else v.numpy()) for k, v in hout.items()}

    def features(self, P):
        """Host (B,N,3) -> host (B,N,128) trunk feature `net` (input of all heads); used to fit synthetic heads."""
        Pd = torch.from_numpy(np.ascontiguousarray(P, dtype=np.float32)).to(self.device)
        B, N, _ = Pd.shape
        net = torch.empty((B, N, 128), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self.forward_device(Pd, net_out=net)
            torch.cuda.current_stream().synchronize()
        return net.cpu().numpy()

    def intermediates(self):
        """Views of the last forward's workspace (indices and per-level features), for parity tests."""
        ws, lay, B = self.last_workspace[:3]
        out = {}
        for name, (dt, shp) in _WS_VIEWS.items():
            shape = shp(self, B)
            nbytes = int(np.prod(shape)) * 4
            off = getattr(lay, name)
            out[name] = ws[off:off + nbytes].view(dt).view(shape)
        return out
