"""ctypes binding of libancsh_b200.so (include/ancsh_b200.h).

There is NO fallback: if the CUDA library has not been built, importing this module raises.
Build it with `python -m articulated_pose_b200.build` (or __graft_entry__.build()).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ANCSH_B200_LIB: load another build of the same library (A/B runs of kernel variants); no fallback either way
LIB_PATH = os.environ.get("ANCSH_B200_LIB") or os.path.join(_HERE, "libancsh_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "articulated_pose_b200: %s is missing. This package has no CPU/PyTorch fallback; build the sm_100a "
        "library first: python -m articulated_pose_b200.build" % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

c_float_p = ctypes.c_void_p  # device pointers are passed as raw addresses
c_int = ctypes.c_int
c_size_t = ctypes.c_size_t

OK = 0
ERRORS = {-1: "ANCSH_ERR_INVALID_ARG", -2: "ANCSH_ERR_CUDA", -3: "ANCSH_ERR_UNSUPPORTED", -4: "ANCSH_ERR_WORKSPACE"}


class AncshError(RuntimeError):
    pass


def check(rc, what):
    if rc != OK:
        raise AncshError("%s failed: %s (%d)" % (what, ERRORS.get(rc, "unknown"), rc))


class Layer(ctypes.Structure):
    _fields_ = [("W", ctypes.c_void_p), ("b", ctypes.c_void_p), ("W_tc", ctypes.c_void_p), ("cin", c_int), ("cout", c_int),
                ("cin_pad", c_int), ("cout_pad", c_int), ("relu", c_int), ("tc_descale", ctypes.c_float)]


class Net(ctypes.Structure):
    _fields_ = [("use_tensor_cores", c_int), ("n_parts", c_int), ("mixed_pred", c_int),
                ("npoint1", c_int), ("nsample1", c_int), ("radius1", ctypes.c_float),
                ("npoint2", c_int), ("nsample2", c_int), ("radius2", ctypes.c_float),
                ("sa1", Layer * 3), ("sa2", Layer * 3), ("sa3", Layer * 3),
                ("fp1_global", Layer), ("fp1", Layer * 2), ("fp2", Layer * 2), ("fp3", Layer * 3),
                ("fc1", Layer), ("nocs_heads", Layer), ("fc3", Layer * 2), ("joint_heads", Layer),
                ("tc_bias_step", c_int), ("sa1_conv0_host", ctypes.c_void_p)]


PRED_FIELDS = ("W", "nocs_per_point", "confi_per_point", "heatmap_per_point", "unitvec_per_point",
               "joint_axis_per_point", "index_per_point", "gocs_per_point", "global_scale", "global_translation")


class Pred(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in PRED_FIELDS + ("net",)]


WS_FIELDS = ("fps_idx1", "l1_xyz", "fps_idx2", "l2_xyz", "ball_idx1", "ball_cnt1", "ball_idx2", "ball_cnt2",
             "l1_points", "l2_points", "l3_points", "fp1_bias", "l2_points_fp", "l1_points_fp", "interp3", "raw_heads",
             "nn_idx2", "nn_w2", "nn_idx3", "nn_w3", "total_bytes")


class WsLayout(ctypes.Structure):
    _fields_ = [(k, c_size_t) for k in WS_FIELDS]


def _sig(name, argtypes, restype=c_int):
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = restype
    return fn


vp = ctypes.c_void_p
ancsh_version = _sig("ancsh_version", [], ctypes.c_char_p)
ancsh_launch_count = _sig("ancsh_launch_count", [], ctypes.c_ulonglong)
ancsh_diag_fp64_fma = _sig("ancsh_diag_fp64_fma", [c_int, c_int, vp, vp])
ancsh_pose_lm_shape = _sig("ancsh_pose_lm_shape", [ctypes.POINTER(c_int), ctypes.POINTER(c_int)])
ancsh_weights_pack = _sig("ancsh_weights_pack", [c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(vp), ctypes.POINTER(c_size_t),
                                                 c_int, c_int, c_int, ctypes.c_char_p, ctypes.POINTER(vp)])
ancsh_packed_destroy = _sig("ancsh_packed_destroy", [vp], None)
ancsh_packed_flat = _sig("ancsh_packed_flat", [vp, ctypes.POINTER(c_size_t)], vp)
ancsh_packed_tc = _sig("ancsh_packed_tc", [vp, ctypes.POINTER(c_size_t)], vp)
ancsh_packed_layer = _sig("ancsh_packed_layer", [vp, ctypes.c_char_p, ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t),
                                                 ctypes.POINTER(c_size_t), ctypes.POINTER(c_int), ctypes.POINTER(c_int)])
ancsh_net_create = _sig("ancsh_net_create", [vp, c_int, c_int, ctypes.POINTER(vp)])
ancsh_net_get = _sig("ancsh_net_get", [vp], vp)
ancsh_net_destroy = _sig("ancsh_net_destroy", [vp], None)
ancsh_ransac_workspace_bytes = _sig("ancsh_ransac_workspace_bytes", [c_int, c_int, c_int, ctypes.POINTER(c_size_t)])
ancsh_ransac_single = _sig("ancsh_ransac_single", [c_int, vp, vp, ctypes.c_double, c_int, vp, ctypes.c_ulonglong, vp, c_size_t,
                                                   vp, vp, vp, vp, vp, vp, vp])
ancsh_ransac_joint = _sig("ancsh_ransac_joint", [c_int, vp, vp, c_int, vp, vp, ctypes.POINTER(ctypes.c_double), ctypes.c_double,
                                                 c_int, vp, vp, ctypes.c_ulonglong, vp, c_size_t, vp, vp, vp, vp, vp, vp, vp, vp,
                                                 vp, vp, vp])
UNIT_IN_FIELDS = ("pts", "cls", "heatmap", "unitvec", "orient", "joint_cls", "nocs_p", "nocs_g")
UNIT_OUT_FIELDS = ("P", "cls_gt", "mask_array", "nocs_gt", "nocs_gt_g", "heatmap_gt", "unitvec_gt", "orient_gt", "joint_cls_gt",
                   "joint_cls_mask")


class UnitIn(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in UNIT_IN_FIELDS]


class UnitOut(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in UNIT_OUT_FIELDS]


ancsh_unit_data = _sig("ancsh_unit_data", [c_int, c_int, c_int, c_int, vp, vp, vp, vp, ctypes.POINTER(UnitIn),
                                           ctypes.POINTER(UnitOut), vp])
ancsh_fps = _sig("ancsh_fps", [c_int, c_int, c_int, vp, vp, vp, vp])
ancsh_fps_two_level = _sig("ancsh_fps_two_level", [c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, vp])
ancsh_gather_point = _sig("ancsh_gather_point", [c_int, c_int, c_int, vp, vp, vp, vp])
ancsh_ball_query = _sig("ancsh_ball_query", [c_int, c_int, c_int, ctypes.c_float, c_int, vp, vp, vp, vp, vp])
ancsh_group_point = _sig("ancsh_group_point", [c_int, c_int, c_int, c_int, c_int, vp, vp, vp, vp])
ancsh_three_nn = _sig("ancsh_three_nn", [c_int, c_int, c_int, vp, vp, vp, vp, vp])
ancsh_three_interpolate = _sig("ancsh_three_interpolate", [c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp])
ancsh_net_plan = _sig("ancsh_net_plan", [ctypes.POINTER(Net), c_int, c_int, ctypes.POINTER(WsLayout)])
ancsh_net_forward = _sig("ancsh_net_forward", [ctypes.POINTER(Net), c_int, c_int, vp, vp, c_size_t,
                                               ctypes.POINTER(Pred), ctypes.POINTER(vp), vp])
ancsh_net_forward_shared = _sig("ancsh_net_forward_shared", [ctypes.POINTER(Net), c_int, c_int, vp, vp, c_size_t,
                                                             ctypes.POINTER(Net), vp, ctypes.POINTER(Pred),
                                                             ctypes.POINTER(vp), vp])
ancsh_event_create = _sig("ancsh_event_create", [ctypes.POINTER(vp)])
ancsh_event_record = _sig("ancsh_event_record", [vp, vp])
ancsh_event_elapsed_ms = _sig("ancsh_event_elapsed_ms", [vp, vp, ctypes.POINTER(ctypes.c_float)])
ancsh_event_destroy = _sig("ancsh_event_destroy", [vp])



class PoseCfg(ctypes.Structure):
    _fields_ = [("n_parts", c_int), ("niter_single", c_int), ("niter_joint", c_int), ("inlier_th", ctypes.c_double),
                ("seed", ctypes.c_ulonglong)]


POSE_IN_FIELDS = ("P", "nocs", "mask", "joint_axis", "joint_cls", "idx_single", "idx_joint0", "idx_joint1")
POSE_OUT_FIELDS = ("single_R", "single_s", "single_t", "single_score", "single_inliers", "joint_R0", "joint_s0",
                   "joint_t0", "joint_R1", "joint_s1", "joint_t1", "joint_score", "joint_inliers0", "joint_inliers1",
                   "part_count", "status")
POSE_WS_FIELDS = ("part_idx", "part_src", "part_tgt", "axis_med", "single_scores", "joint_scores", "single_best",
                  "joint_best", "joint_nfev", "joint_models", "joint_tail", "total_bytes")


class PoseIn(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in POSE_IN_FIELDS]


class PoseOut(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in POSE_OUT_FIELDS]


class PoseWs(ctypes.Structure):
    _fields_ = [(k, c_size_t) for k in POSE_WS_FIELDS]


ancsh_pose_plan = _sig("ancsh_pose_plan", [ctypes.POINTER(PoseCfg), c_int, c_int, ctypes.POINTER(PoseWs)])
ancsh_pose_solve = _sig("ancsh_pose_solve", [ctypes.POINTER(PoseCfg), ctypes.POINTER(PoseIn), c_int, c_int, vp, c_size_t,
                                             ctypes.POINTER(PoseOut), ctypes.POINTER(vp), vp])
POSE_STAGES = ("partition", "single_score", "single_refit", "joint_score", "joint_refit")
ancsh_pose_sample_indices = _sig("ancsh_pose_sample_indices", [ctypes.c_ulonglong, c_int, c_int, c_int, vp, vp, vp])
ancsh_umeyama = _sig("ancsh_umeyama", [c_int, c_int, vp, vp, vp, vp, vp, vp, vp])

ancsh_similarity_ransac = _sig("ancsh_similarity_ransac", [c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp])
ancsh_amodal_extent = _sig("ancsh_amodal_extent", [c_int, c_int, c_int, vp, vp, vp, vp, vp])
ancsh_joint_vote = _sig("ancsh_joint_vote", [c_int, c_int, c_int, c_int, c_int, vp, vp, vp, vp, vp, vp, ctypes.c_float, vp, vp, vp, vp])
ancsh_box_iou_3d = _sig("ancsh_box_iou_3d", [c_int, c_int, vp, vp, vp, vp, vp, vp])

NET_STAGES = ("fps1", "fps2", "ball1", "sa1", "ball2", "sa2", "sa3", "fp1", "fp2", "fp3_heads")


class EventList:
    """n timing-enabled CUDA events owned through the C ABI."""

    def __init__(self, n):
        self.n = n
        self.arr = (vp * n)()
        for i in range(n):
            e = vp()
            check(ancsh_event_create(ctypes.byref(e)), "ancsh_event_create")
            self.arr[i] = e.value

    def elapsed_ms(self, i, j):
        ms = ctypes.c_float()
        check(ancsh_event_elapsed_ms(self.arr[i], self.arr[j], ctypes.byref(ms)), "ancsh_event_elapsed_ms")
        return float(ms.value)

    def __del__(self):
        try:
            for i in range(self.n):
                ancsh_event_destroy(self.arr[i])
        except Exception:
            pass
