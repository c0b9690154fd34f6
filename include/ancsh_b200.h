/*
 * ancsh_b200.h -- C ABI of libancsh_b200.so: the B200 (sm_100a) implementation of the ANCSH
 * per-point-cloud hot path (PointNet++ forward -> per-part RANSAC -> joint-constrained solve).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); no call synchronises;
 *   - dense row-major layouts identical to the reference's TF tensors, f32 / int32;
 *   - return value: ANCSH_OK or a negative ANCSH_ERR_* (the reference's ops raise InvalidArgument via
 *     OP_REQUIRES, tf_sampling.cpp:105, tf_grouping.cpp:79-84; its launchers never check CUDA errors);
 *   - no global state: every entry point is re-entrant and may be called from several host threads on
 *     different streams.
 *
 * Citations are relative to /root/reference (dragonlong/articulated-pose @ 267b70a4).
 */
#ifndef ANCSH_B200_H
#define ANCSH_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANCSH_OK 0
#define ANCSH_ERR_INVALID_ARG (-1) /* shape / size constraint violated (reference: InvalidArgument) */
#define ANCSH_ERR_CUDA (-2)        /* a CUDA runtime call or launch failed */
#define ANCSH_ERR_UNSUPPORTED (-3) /* legal in the reference, outside what this build covers */
#define ANCSH_ERR_WORKSPACE (-4)   /* workspace too small */

/* library / build identification: returns e.g. "ancsh_b200 0.1 sm_100a" */
const char *ancsh_version(void);

/* Diagnostic: number of kernels this library has launched in this process so far (monotonic, all threads and streams;
 * memsets / copies are not counted).  bench.py reports the difference over its timed region as `gpu_launches`. */
unsigned long long ancsh_launch_count(void);

/* Diagnostic: FP64 peak probe.  Launches `blocks` blocks of 256 threads that each run 8 independent chains of `iters`
 * DFMAs (= blocks * 256 * 8 * iters * 2 FLOP); time it with events on `stream`.  bench.py uses the result as the roofline
 * denominator of the f64 pose kernels.  sink: one device double (never written in practice). */
int ancsh_diag_fp64_fma(int blocks, int iters, double *sink, void *stream);

/* Diagnostic: launch shape of the joint LM solve (joint_lm_kernel) on the current device: threads per block and the most
 * blocks one of its phases uses.  The stage is deliberately narrow (the lanes are latency bound; SMs without LM blocks keep
 * two forward CTAs of the overlapped batches); bench.py reports the SMs it occupies next to its roofline fraction. */
int ancsh_pose_lm_shape(int *threads_per_block, int *max_blocks);

/* ------------------------------------------------------------------------------------------------
 * Op level -- one entry point per native op on the path.
 * ---------------------------------------------------------------------------------------------- */

/* Replaces farthestpointsamplingLauncher(int b,int n,int m,const float*inp,float*temp,int*out)
 * (pointnet_plusplus/utils/tf_ops/sampling/tf_sampling_g.cu:203; Python: tf_sampling.py:48
 * farthest_point_sample(npoint, inp)).  inp (b,n,3) -> out (b,m) int32.  `temp` is the reference's
 * 32*n-float scratch; it is accepted for signature compatibility and ignored (may be NULL).
 * Bit-exact with the reference kernel, including its tie rule.  n <= 8192. */
int ancsh_fps(int b, int n, int m, const float *inp, float *temp, int *out, void *stream);

/* Both sampling levels of the PointNet++ trunk in one launch: out1 (b,m1) = farthest_point_sample(m1, inp) and
 * xyz1 (b,m1,3) = gather_point(inp, out1) (layer1, pointnet_util.py:47), out2 (b,m2) = farthest_point_sample(m2, xyz1)
 * and xyz2 (b,m2,3) = gather_point(xyz1, out2) (layer2).  The second level is the prefix of the first (greedy farthest
 * point order), so it costs nothing; bit-exact with two launches of the reference kernel while m1 <= 512 (its tie rule,
 * tf_sampling_g.cu:146-161, orders equal distances by index only below 512 points) -- ANCSH_ERR_UNSUPPORTED otherwise. */
int ancsh_fps_two_level(int b, int n, int m1, int m2, const float *inp, int *out1, float *xyz1, int *out2, float *xyz2,
                        void *stream);

/* Replaces gatherpointLauncher (tf_sampling_g.cu:206; tf_sampling.py:29 gather_point(inp, idx)).
 * inp (b,n,3), idx (b,m) -> out (b,m,3). */
int ancsh_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out, void *stream);

/* Replaces queryBallPointLauncher (tf_grouping_g.cu:125; tf_grouping.py:8
 * query_ball_point(radius, nsample, xyz1, xyz2)).  xyz1 (b,n,3) dataset, xyz2 (b,m,3) centroids ->
 * idx (b,m,nsample), pts_cnt (b,m).  First-nsample-in-index-order semantics, bit-exact.  Rows with no
 * point in the ball (the reference leaves them uninitialised) are filled with 0. */
int ancsh_ball_query(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                     int *idx, int *pts_cnt, void *stream);

/* Replaces groupPointLauncher (tf_grouping_g.cu:133; tf_grouping.py:33 group_point(points, idx)).
 * points (b,n,c), idx (b,m,nsample) -> out (b,m,nsample,c). */
int ancsh_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out,
                      void *stream);

/* Replaces threenn_cpu (3d_interpolation/tf_interpolate.cpp:60; tf_interpolate.py:8 three_nn(xyz1,xyz2)).
 * xyz1 (b,n,3) queries, xyz2 (b,m,3) known -> dist (b,n,3) squared f32, idx (b,n,3).  Bit-exact
 * (un-fused f32 distance, lowest index wins ties, missing neighbours = (inf, 0)). */
int ancsh_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, void *stream);

/* Replaces threeinterpolate_cpu (tf_interpolate.cpp:107; tf_interpolate.py:19
 * three_interpolate(points, idx, weight)).  points (b,m,c), idx (b,n,3), weight (b,n,3) -> out (b,n,c). */
int ancsh_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight,
                            float *out, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Network level -- replaces sess.run(pred_dict) of lib/network.py:292 for the graph built by
 * lib/architecture.py:86-161 (get_per_point_model_new) on pointnet_plusplus/architectures.py:56-95.
 * ---------------------------------------------------------------------------------------------- */

/* One 1x1-conv layer with bias, inference batch-norm already folded in, zero padded:
 *   W  [cin_pad][cout_pad] row-major (TF weights [1,1,cin,cout] are already [cin][cout]),
 *   b  [cout_pad] (or [B][cout_pad] when bias_stride != 0: per-cloud bias, used by fa_layer1).
 * cin_pad % 16 == 0, cout_pad == 64 or cout_pad % 128 == 0.  The host packer
 * (articulated_pose_b200/weights.py) produces these from the TF variable names. */
typedef struct {
    const float *W;
    const float *b;
    const void *W_tc; /* NULL, or the tensor-core operand image of W: fp16 [cin_pad/8][2 (hi,lo)][cout_pad][8] with
                         W^T split as hi = fp16(w), lo = fp16(w - hi) (weights.tc_image); used when use_tensor_cores */
    int cin, cout, cin_pad, cout_pad;
    int relu;
    float tc_descale; /* W_tc holds W * 2^s (s = the layer's power-of-two scale, weights.tc_scale_exp: the largest |w| or |b|
                         lands in [2^13, 2^14), so that BN-folded weights of any magnitude keep both fp16 pieces in the
                         normal range); the kernels multiply the accumulators by tc_descale = 2^-s (exact).  1 if unscaled. */
} ancsh_layer_t;

typedef struct {
    int use_tensor_cores; /* 1: grouped MLPs on tcgen05 (fp16 hi/lo split, 3 products, f32 accumulate in TMEM); 0: exact-f32 CUDA cores */
    int n_parts;    /* K = n_max_parts (eyeglasses 3, drawer 4) */
    int mixed_pred; /* 1: ANCSH heads [K,3K,K,3K,1] (architecture.py:98-102); 0: NPCS heads [K,3K,1] */
    int npoint1, nsample1;
    float radius1; /* layer1: 512, 64, 0.2 (architectures.py:62-65) */
    int npoint2, nsample2;
    float radius2; /* layer2: 128, 64, 0.4 (architectures.py:67-70) */
    /* channel order inside each first layer is [features..., xyz] (rows of W permuted by the packer;
     * the reference concatenates [xyz, features], pointnet_util.py:57,84) */
    ancsh_layer_t sa1[3], sa2[3], sa3[3];
    ancsh_layer_t fp1_global; /* rows of fa_layer1/conv_0 that multiply the (broadcast) global feature */
    ancsh_layer_t fp1[2];     /* fp1[0]: remaining rows (skip features) of conv_0; fp1[1]: conv_1 */
    ancsh_layer_t fp2[2], fp3[3];
    ancsh_layer_t fc1;
    ancsh_layer_t nocs_heads;  /* 128 -> [W(K) | nocs(3K) | scale(K) | trans(3K) | confi(1)], fc11_1 folded */
    ancsh_layer_t fc3[2];      /* joint_net/fc3_0, fc3_1 */
    ancsh_layer_t joint_heads; /* 128 -> [joint_axis(3) | unitvec(3) | heatmap(1) | index(3)] */
    int tc_bias_step; /* 1: every W_tc image carries one extra 16-deep k-step after cin_pad whose rows are (fp16(b),
                         fp16(b - fp16(b)), 0, ...): the bias is added by the GEMM itself (weights.tc_image /
                         ancsh_tc_image); required by the layer-specialised kernels (net_lean.cu) */
    const float *sa1_conv0_host; /* HOST pointer or NULL: 256 floats = rows 0..2 of the BN-folded sa1[0].W (64 columns)
                                    followed by its bias; the xyz-only first conv of layer1 is evaluated on the CUDA cores
                                    with these values in the kernel parameter (constant) bank */
} ancsh_net_t;

/* ---- weight import in C: TF1 checkpoint variables -> ancsh_net_t ---------------------------------------------
 * The import contract is the checkpoint's variable names: `<prefix>/est_net/layer{1,2,3}/conv{0,1,2}`,
 * `<prefix>/est_net/fa_layer{1,2,3}/conv_{i}`, `<prefix>/est_net/fc1`, `<prefix>/nocs_net/{fc11_1,fc2_i}`,
 * `<prefix>/joint_net/{fc3_0,fc3_1,fc4_i}` (pointnet_util.py:128,234, architectures.py:65-90, lib/architecture.py:105-120,
 * 195-208), each with `/weights`, `/biases` (tf_util.py:164,174) and, where the layer has batch norm,
 * `/bn/{beta,gamma,moving_mean,moving_variance}` (tf_util.py:527-531).  articulated_pose_b200/weights.py is the Python
 * mirror of the same steps (BN fold in f64, [xyz, features] -> [features, xyz] row permutation, fa_layer1 split, head
 * packing incl. the fc11_1 fold, padding, tensor-core images). */
typedef struct ancsh_packed ancsh_packed_t;         /* host-side packed network (owns its buffers) */
typedef struct ancsh_net_handle ancsh_net_handle_t; /* device-resident network (owns the device buffers of its ancsh_net_t) */

/* names[i] / data_host[i] / counts[i]: variable name, HOST f32 values, element count.  prefix NULL = "SPFN".
 * ANCSH_ERR_INVALID_ARG: a variable is missing or has the wrong size; ANCSH_ERR_UNSUPPORTED: non-finite values. */
int ancsh_weights_pack(int n_vars, const char *const *names, const float *const *data_host, const size_t *counts, int n_parts,
                       int mixed_pred, int early_split_nocs, const char *prefix, ancsh_packed_t **out);
void ancsh_packed_destroy(ancsh_packed_t *p);
/* Introspection (host): the f32 buffer of all padded W / b, the fp16 buffer of all tensor-core images, and per slot
 * ("sa1[0]" ... "joint_heads", the field names of ancsh_net_t) the element offsets into them, dims5 = {cin, cout, cin_pad,
 * cout_pad, relu} and the image's power-of-two scale exponent.  tc_off = (size_t)-1: the slot has no image. */
const float *ancsh_packed_flat(const ancsh_packed_t *p, size_t *count);
const unsigned short *ancsh_packed_tc(const ancsh_packed_t *p, size_t *count);
int ancsh_packed_layer(const ancsh_packed_t *p, const char *slot, size_t *w_off, size_t *b_off, size_t *tc_off, int *dims5,
                       int *tc_exp);
/* Uploads the packed buffers to the current device (cudaMalloc + synchronous copies) and fills an ancsh_net_t with the
 * reference's level settings (npoint 512 / 128, radius 0.2 / 0.4, architectures.py:62-70) and nsample. */
int ancsh_net_create(const ancsh_packed_t *p, int nsample, int use_tensor_cores, ancsh_net_handle_t **out);
const ancsh_net_t *ancsh_net_get(const ancsh_net_handle_t *h);
void ancsh_net_destroy(ancsh_net_handle_t *h);

/* Output tensors of pred_dict (architecture.py:141-159), each (B,N,width) f32 dense.  gocs_per_point,
 * global_scale, global_translation are written only when mixed_pred (may be NULL otherwise). */
typedef struct {
    float *W;                    /* (B,N,K)  softmax */
    float *nocs_per_point;       /* (B,N,3K) sigmoid */
    float *confi_per_point;      /* (B,N,1)  sigmoid */
    float *heatmap_per_point;    /* (B,N,1)  sigmoid */
    float *unitvec_per_point;    /* (B,N,3)  tanh */
    float *joint_axis_per_point; /* (B,N,3)  tanh */
    float *index_per_point;      /* (B,N,3)  softmax */
    float *gocs_per_point;       /* (B,N,3K) nocs*scale+trans */
    float *global_scale;         /* (B,N,K)  sigmoid */
    float *global_translation;   /* (B,N,3K) tanh */
    float *net;                  /* optional (may be NULL): (B,N,128) trunk feature after fc1 (architectures.py:89-93),
                                    not part of pred_dict; used to fit synthetic head weights */
} ancsh_pred_t;

/* Byte offsets of the intermediates inside the caller-provided workspace (all 256-B aligned). */
typedef struct {
    size_t fps_idx1;  /* int32 (B,npoint1) */
    size_t l1_xyz;    /* f32   (B,npoint1,3) */
    size_t fps_idx2;  /* int32 (B,npoint2) */
    size_t l2_xyz;    /* f32   (B,npoint2,3) */
    size_t ball_idx1; /* int32 (B,npoint1,nsample1) */
    size_t ball_cnt1; /* int32 (B,npoint1) */
    size_t ball_idx2; /* int32 (B,npoint2,nsample2) */
    size_t ball_cnt2; /* int32 (B,npoint2) */
    size_t l1_points; /* f32 (B,npoint1,128)  layer1 output */
    size_t l2_points; /* f32 (B,npoint2,256)  layer2 output */
    size_t l3_points; /* f32 (B,1024)         layer3 output */
    size_t fp1_bias;  /* f32 (B,256)          per-cloud bias of fa_layer1/conv_0 */
    size_t l2_points_fp; /* f32 (B,npoint2,256) fa_layer1 output */
    size_t l1_points_fp; /* f32 (B,npoint1,128) fa_layer2 output */
    size_t interp3;      /* (B*npoint2*512*4 bytes) scratch: layer3 conv_1 output as the fp16 hi/lo operand image of conv_2, per
                            128-row tile [512/8][hi|lo][128][8] (tensor-core path only; the name is historic -- the interpolated
                            rows of the fa layers are built on chip and never stored) */
    size_t raw_heads;    /* unused (0 bytes): the head activations run in the epilogue of the head layers */
    size_t nn_idx2;      /* int32 (B,npoint1,3) three_nn of l1_xyz in l2_xyz (fa_layer2); geometry, shared like the FPS indices */
    size_t nn_w2;        /* f32   (B,npoint1,3) inverse-distance weights (pointnet_util.py:219-222) */
    size_t nn_idx3;      /* int32 (B,N,3) three_nn of P in l1_xyz (fa_layer3) */
    size_t nn_w3;        /* f32   (B,N,3) */
    size_t total_bytes;
} ancsh_ws_layout_t;

/* Fills *layout for a batch of B clouds of N points.  N % 128 == 0 is required. */
int ancsh_net_plan(const ancsh_net_t *net, int B, int N, ancsh_ws_layout_t *layout);

/* Stages of ancsh_net_forward, in launch order (for per-stage CUDA-event timing). */
enum {
    ANCSH_STAGE_FPS1 = 0, ANCSH_STAGE_FPS2, ANCSH_STAGE_BALL1, ANCSH_STAGE_SA1, ANCSH_STAGE_BALL2, ANCSH_STAGE_SA2,
    ANCSH_STAGE_SA3, ANCSH_STAGE_FP1, ANCSH_STAGE_FP2, ANCSH_STAGE_FP3_HEADS, ANCSH_NET_NSTAGES
};

/* Forward pass for B clouds: P (B,N,3) -> pred.  workspace must hold layout.total_bytes.
 * stage_events: NULL, or ANCSH_NET_NSTAGES+1 cudaEvent_t handles (ancsh_event_create); event i is recorded on
 * `stream` before stage i and the last one after the final stage, so stage i took elapsed(ev[i], ev[i+1]). */
int ancsh_net_forward(const ancsh_net_t *net, int B, int N, const float *P, void *workspace, size_t workspace_bytes,
                      const ancsh_pred_t *pred, void *const *stage_events, void *stream);

/* Timing-enabled CUDA events for callers that have no CUDA runtime binding of their own. */
/* Second network over the SAME cloud batch (the reference's pose stage reads NOCS + segmentation from the NPCS baseline
 * network and the joint axis from the ANCSH network, parallel_ancsh_pose.py:197,232-236; both run
 * farthest_point_sample / query_ball_point on the same xyz, pointnet_util.py:47-49): the sampling indices, level
 * coordinates and ball-query indices are read from `geometry_workspace`, the workspace of a completed or enqueued
 * ancsh_net_forward of `geometry_net` on the same stream, instead of being recomputed.  The two networks must agree
 * in npoint / radius / nsample (ANCSH_ERR_INVALID_ARG otherwise). */
int ancsh_net_forward_shared(const ancsh_net_t *net, int B, int N, const float *P, void *workspace, size_t workspace_bytes,
                             const ancsh_net_t *geometry_net, const void *geometry_workspace, const ancsh_pred_t *pred,
                             void *const *stage_events, void *stream);

int ancsh_event_create(void **event_out);
int ancsh_event_record(void *event, void *stream);
int ancsh_event_elapsed_ms(void *start, void *stop, float *ms_out); /* both events must have completed */
int ancsh_event_destroy(void *event);


/* ------------------------------------------------------------------------------------------------
 * Pose level -- replaces the per-cloud body of solver_ransac_nonlinear
 * (evaluation/parallel_ancsh_pose.py:214-352): K single-part RANSACs (ransac() :20-33 with
 * single_transformation_estimator/_verifier :35-54) and K-1 joint RANSACs part 0 <-> part j
 * (joint_transformation_estimator/_verifier :106-194, LM of :154-155), each followed by the refit on the
 * best hypothesis' inliers.  All arithmetic in f64 on f32 inputs.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int n_parts;             /* K */
    int niter_single;        /* reference: 10000 (parallel_ancsh_pose.py:262) */
    int niter_joint;         /* reference: 200   (parallel_ancsh_pose.py:288) */
    double inlier_th;        /* reference: 0.1   (pose_multi_process.py:32) */
    unsigned long long seed; /* Philox4x32-10 key, used when the idx_* tensors are NULL */
} ancsh_pose_cfg_t;

typedef struct {
    const float *P;          /* (B,N,3)  input cloud                       (h5 'P') */
    const float *nocs;       /* (B,N,3K) nocs_per_point                    (h5 'nocs_per_point') */
    const float *mask;       /* (B,N,K)  instance_per_point (W); argmax = part label, :238 */
    const float *joint_axis; /* (B,N,3)  joint_axis_per_point; median over the joint's points, :295 */
    const int *joint_cls;    /* (B,N)    joint_cls_gt: j in 1..K-1 marks points of joint j, :244-247 */
    const int *idx_single;   /* NULL or (B,K,niter_single,3): sample POSITIONS inside the part's point list */
    const int *idx_joint0;   /* NULL or (B,K-1,niter_joint,3): samples of part 0 */
    const int *idx_joint1;   /* NULL or (B,K-1,niter_joint,3): samples of part j */
} ancsh_pose_in_t;

#define ANCSH_POSE_EMPTY_PART 1   /* status bit: the part has no points (the reference raises ValueError) */
#define ANCSH_POSE_NO_INLIERS 2   /* status bit: best hypothesis has no inliers -> NaN model */

typedef struct {
    /* single-part RANSAC ('baseline' entries of the reference's result dict, :281-285) */
    double *single_R;               /* (B,K,9) row-major rotation */
    double *single_s;               /* (B,K) */
    double *single_t;               /* (B,K,3) */
    int *single_score;              /* (B,K) inlier count of the best hypothesis */
    unsigned char *single_inliers;  /* (B,K,N) best_inliers mask over the part's point list (first part_count entries) */
    /* joint RANSAC ('nonlinear' entries, :325-343); entry j-1 couples part 0 with part j */
    double *joint_R0, *joint_s0, *joint_t0; /* (B,K-1,9) (B,K-1) (B,K-1,3) */
    double *joint_R1, *joint_s1, *joint_t1;
    double *joint_score;            /* (B,K-1) score of the best hypothesis, (n0/3+n1/3)/2 as in :193 */
    unsigned char *joint_inliers0;  /* (B,K-1,N) */
    unsigned char *joint_inliers1;  /* (B,K-1,N) */
    int *part_count;                /* (B,K) */
    int *status;                    /* (B,K) ANCSH_POSE_* bits (joint j reports in entry j) */
} ancsh_pose_out_t;

typedef struct {
    size_t part_idx;       /* int32 (B,K,N) point indices of each part, ascending (np.where, :241) */
    size_t part_src;       /* f32 (B,K,N,3) nocs of the part's points, channels 3j..3j+2 */
    size_t part_tgt;       /* f32 (B,K,N,3) P of the part's points */
    size_t axis_med;       /* f64 (B,K-1,3) per-joint median axis */
    size_t single_scores;  /* int32 (B,K,niter_single) per-hypothesis inlier counts */
    size_t joint_scores;   /* f64 (B,K-1,niter_joint) per-hypothesis scores */
    size_t single_best;    /* int32 (B,K) index of the winning hypothesis */
    size_t joint_best;     /* int32 (B,K-1) */
    size_t joint_nfev;     /* int32 (B,K-1,niter_joint) residual evaluations each hypothesis' LM solve took */
    size_t joint_models;   /* f64 (B,K-1,niter_joint,26) per-hypothesis {R0[9],s0,t0[3],R1[9],s1,t1[3]} */
    size_t joint_tail;     /* internal: 64 int32 counters, then the LM solve records and the work lists of suspended solves.
                              Counters [60] / [61] = Jacobian evaluations / lmpar iterations summed over all LM solves of the
                              call (statistics for the FLOP model of the joint stage) */
    size_t total_bytes;
} ancsh_pose_ws_t;

int ancsh_pose_plan(const ancsh_pose_cfg_t *cfg, int B, int N, ancsh_pose_ws_t *layout);

enum {
    ANCSH_POSE_STAGE_PARTITION = 0, ANCSH_POSE_STAGE_SINGLE_SCORE, ANCSH_POSE_STAGE_SINGLE_REFIT,
    ANCSH_POSE_STAGE_JOINT_SCORE, ANCSH_POSE_STAGE_JOINT_REFIT, ANCSH_POSE_NSTAGES
};

/* Solves B clouds.  N <= 4096, K <= 8.  stage_events: NULL or ANCSH_POSE_NSTAGES+1 events (see ancsh_net_forward). */
int ancsh_pose_solve(const ancsh_pose_cfg_t *cfg, const ancsh_pose_in_t *in, int B, int N, void *workspace,
                     size_t workspace_bytes, const ancsh_pose_out_t *out, void *const *stage_events, void *stream);

/* ransac(dataset, estimator, verifier, inlier_th, niter) of evaluation/parallel_ancsh_pose.py:20 on a caller dataset, for
 * hosts without the Python mirror (pose.ransac_single / ransac_joint do the same through ancsh_pose_solve).
 * source / target: (n,3) f32 DEVICE arrays (dataset['source'], dataset['target']); sample_idx: NULL (Philox, `seed`) or
 * DEVICE (niter,3) int32 sample positions (the reference draws np.random.randint(nsource, size=3), :38).
 * Results (DEVICE): best_model = {R (9, row-major), scale, t (3)} refitted on the winner's inliers, score = its inlier count
 * (may be NULL), inliers (n) bytes, status (1 int: ANCSH_POSE_* bits).  workspace: ancsh_ransac_workspace_bytes(n, 1, niter). */
int ancsh_ransac_workspace_bytes(int n_total, int n_parts, int niter, size_t *bytes);
int ancsh_ransac_single(int n, const float *source, const float *target, double inlier_th, int niter, const int *sample_idx,
                        unsigned long long seed, void *workspace, size_t workspace_bytes, double *R, double *scale, double *t,
                        int *score, unsigned char *inliers, int *status, void *stream);
/* Joint variant (joint_transformation_estimator / _verifier, :106-194): dataset = {source0, target0, source1, target1,
 * joint_direction (3 HOST doubles)} -> rotation0/scale0/translation0/rotation1/... (:177-184), score (may be NULL),
 * inliers0 (n0) / inliers1 (n1) bytes, status (2 ints; entry 1 reports the joint).
 * workspace: ancsh_ransac_workspace_bytes(n0 + n1, 2, niter). */
int ancsh_ransac_joint(int n0, const float *source0, const float *target0, int n1, const float *source1, const float *target1,
                       const double *joint_direction_host, double inlier_th, int niter, const int *sample_idx0,
                       const int *sample_idx1, unsigned long long seed, void *workspace, size_t workspace_bytes, double *R0,
                       double *s0, double *t0, double *R1, double *s1, double *t1, double *score, unsigned char *inliers0,
                       unsigned char *inliers1, int *status, void *stream);

/* Writes the sample positions the Philox generator produces for (seed, problem, hypothesis) so that a CPU
 * checker can replay them: idx (nprob,niter,3); n_per_problem (nprob) is the part size each problem samples
 * from; stream_id 0 = single RANSAC, 1 / 2 = part 0 / part j of the joint RANSAC. */
int ancsh_pose_sample_indices(unsigned long long seed, int stream_id, int nprob, int niter, const int *n_per_problem,
                              int *idx, void *stream);

/* Batched Umeyama similarity (lib/aligning.py:580-622 estimateSimilarityUmeyama, as called by
 * evaluation/compute_gt_pose.py:87 for the GT part poses): src/tgt (nprob,nmax,3) f32, cnt (nprob) ->
 * scale (nprob), R (nprob,9) row-major in the REFERENCE's convention (Rotation = (U Vh)^T), t (nprob,3). */
int ancsh_umeyama(int nprob, int nmax, const float *src, const float *tgt, const int *cnt, double *scale, double *R,
                  double *t, void *stream);

/* estimateSimilarityTransform (lib/aligning.py:17-33): per problem the thresholds of set_config
 * (:88-103: PassT = max(|T|/|S|, |S|/|T|) of the mean point norms, StopT = PassT / 100), the 5-point Umeyama RANSAC of
 * getRANSACInliers (:485-507) over niter <= 256 hypotheses whose samples are idx (nprob,niter,5) (the reference draws them
 * from np.random.randint), evaluateModel's bookkeeping (:540-547, incl. np.count_nonzero over the inlier INDICES, i.e.
 * point 0 never counts, strict `>` on the ratio, stop once BestResidual < StopT), and the Umeyama refit on the best
 * hypothesis' inliers.  src/tgt (nprob,nmax,3) f32, cnt (nprob) ->
 * scale (nprob), R (nprob,9) [reference convention (U Vh)^T], t (nprob,3), inlier_ratio (nprob) = BestInlierRatio,
 * inliers (nprob,nmax) bytes = BestInlierIdx as a mask, iters_run (nprob), status (nprob): 0, ANCSH_POSE_EMPTY_PART, or
 * ANCSH_POSE_NO_INLIERS when BestInlierRatio < 0.1 (the reference returns four Nones, :25-27) -> NaN model. */
int ancsh_similarity_ransac(int nprob, int nmax, int niter, const float *src, const float *tgt, const int *cnt,
                            const int *idx, double *scale, double *R, double *t, double *inlier_ratio,
                            unsigned char *inliers, int *iters_run, int *status, void *stream);

/* ---- metric kernels behind the pose stage -------------------------------------------------------- */

/* Amodal box extents of the predicted parts (evaluation/compute_miou.py:187,196-199): for cloud b and part j,
 * extent[b,j,:] = 2 * max_i |nocs[b,i,3j:3j+3] - 0.5| over the points with argmax_k mask[b,i,k] == j (first maximum),
 * evaluated in f32 like NumPy does on the f32 h5 arrays.  nocs (B,N,3K) f32, mask (B,N,K) f32 ->
 * extent (B,K,3) f32, count (B,K) int32.  Empty part: extent = NaN, count = 0 (the reference raises inside np.max
 * and its bare `except` drops the cloud, :250). */
int ancsh_amodal_extent(int B, int N, int K, const float *nocs, const float *mask, float *extent, int *count,
                        void *stream);

/* lib/d3_utils.py:55-69 iou_3d(bbox1, bbox2, nres=50) for npairs box pairs: bbox1/bbox2 (npairs,8,3) f64 corners in
 * get_3d_bbox's order (:7-38; pts_inside_box reads corners 4,5,7,0, :43-45) -> iou (npairs) f64 = intersect/union of the
 * nres^3 np.linspace grid samples strictly inside both / either box, 1.0 when union == 0 (:66-67).  inter / uni
 * (npairs) int32 receive the two counts; either may be NULL. */
int ancsh_box_iou_3d(int npairs, int nres, const double *bbox1, const double *bbox2, double *iou, int *inter, int *uni,
                     void *stream);

/* Joint parameter voting (evaluation/eval_joint_params.py:178-190): for cloud b and joint j = 1..K-1, over
 * the points with argmax(index_per_point) == j (first maximum; index_per_point is n_index wide -- 3 even for the 4-part
 * drawer, lib/architecture.py:129), axis_out[b,j-1,:] = median(joint_axis_per_point), pt_out[b,j-1,:] =
 * median(gn + unitvec * (1 - heatmap) * thres_r) with gn[i] = gocs[i, 3c:3c+3], c = argmax(mask[i]) (gn_width == 3K) or
 * gocs[i, :3] (gn_width == 3), all in f32 like NumPy on the f32 h5 arrays.  count (B,K-1) = voters; none -> NaN.
 * gocs (B,N,gn_width), mask (B,N,K), unitvec / joint_axis (B,N,3), heatmap (B,N,1), index_per_point (B,N,n_index) f32. */
int ancsh_joint_vote(int B, int N, int K, int gn_width, int n_index, const float *gocs, const float *mask,
                     const float *unitvec, const float *heatmap, const float *joint_axis, const float *index_per_point,
                     float thres_r, float *axis_out, float *pt_out, int *count, void *stream);

/* Sampling / normalisation of create_unit_data_from_hdf5 (lib/dataset.py:290-317, 346-372), the step in
 * front of the network: per cloud b, output point i takes source point perm[b,i] % n_total[b] (the reference tiles clouds
 * with fewer than num_points points, :290-317, and draws perm = np.random.permutation(n_total)[:num_points], :346 -- here
 * the caller passes perm, like the RANSAC samples), P = pts * norm_factor[b] (:351), cls_gt, the one-hot mask_array (:360),
 * joint_cls_mask = joint_cls > 0 (:356-358) and the gathered GT arrays.  rot: NULL, or (B,9) row-major f64 rotation of the
 * sapien branch (:369-377): NOCS arrays are rotated about 0.5, unitvec / orient about 0.
 * Inputs (B,n_max,width) f32, outputs (B,num_points,width) f32; optional arrays may be NULL on either side. */
typedef struct {
    const float *pts;       /* (B,n_max,3) required */
    const float *cls;       /* (B,n_max)   required: part label per point */
    const float *heatmap;   /* (B,n_max)   */
    const float *unitvec;   /* (B,n_max,3) */
    const float *orient;    /* (B,n_max,3) joint orientation per point */
    const float *joint_cls; /* (B,n_max)   */
    const float *nocs_p;    /* (B,n_max,3) part NOCS */
    const float *nocs_g;    /* (B,n_max,3) global NOCS */
} ancsh_unit_in_t;
typedef struct {
    float *P, *cls_gt;      /* (B,num_points,3), (B,num_points) required */
    float *mask_array;      /* (B,num_points,n_parts) */
    float *nocs_gt, *nocs_gt_g, *heatmap_gt, *unitvec_gt, *orient_gt, *joint_cls_gt, *joint_cls_mask;
} ancsh_unit_out_t;
int ancsh_unit_data(int B, int n_max, int num_points, int n_parts, const int *n_total, const int *perm,
                    const float *norm_factor, const double *rot, const ancsh_unit_in_t *in, const ancsh_unit_out_t *out,
                    void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ANCSH_B200_H */
