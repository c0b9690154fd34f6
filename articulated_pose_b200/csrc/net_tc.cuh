// net_tc.cuh -- argument structs of the tensor-core (tcgen05) network kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../include/ancsh_b200.h"

struct TcLayer {
    const __half *Wimg;   // [K/8][2 (hi,lo)][N][8] fp16: W^T pre-split and pre-tiled (weights.tc_image)
    const float *bias;           // [N] f32 (BN folded)
    int K, N, relu;              // K = cin_pad (multiple of 16), N = cout_pad (64 / 128 / 256)
    int has_bias_step;           // the image carries the extra bias k-step after K (ancsh_net_t::tc_bias_step)
    float descale;               // the image holds W * 2^s: accumulators are multiplied by descale = 2^-s (ancsh_layer_t::tc_descale)
};

struct SaTcArgs {
    const float *xyz, *points, *new_xyz;
    const int *idx;
    float *out;
    TcLayer L[3];
    int n, m, S, C;
    const float *W0;             // optional f32 [K0][N0] weights of L[0] (ancsh_layer_t::W): with C == 0 (xyz-only input) the
                                 // warp-specialised kernel evaluates L[0] on the CUDA cores inside the gather
    int kmax8, nch;              // filled by the launcher: operand width / 8, output chunk width
    uint32_t tmem_cols;
};

int sa_tc2_launch(const SaTcArgs &a, int B, cudaStream_t st);    // generic warp-specialised chain (net_tc2.cu)

// ---- layer-specialised set abstraction with fused ball query (net_lean.cu) --------------------------------------
struct SaLeanArgs2 {
    const float *xyz, *points, *new_xyz;
    const int *idx_in;           // (B,m,S) ball-query indices, or NULL: the kernel runs the ball query itself
    int *idx_out, *cnt_out;      // idx_in == NULL: optional copies of idx (B,m,S) / pts_cnt (B,m)
    float *out;                  // (B,m,L[2].N)
    TcLayer L[3];
    int n, m, S, C;
    float radius;
    const float *w0_host;        // HOST: ancsh_net_t::sa1_conv0_host (C == 0 only)
};
// ANCSH_ERR_UNSUPPORTED when the stage is not one of the specialised shapes (caller falls back to sa_tc2_launch)
int sa_lean_launch(const SaLeanArgs2 &a, int B, cudaStream_t st);

// ---- generic row-tile chain (feature propagation / heads) ------------------------------------------------------
enum { TC_DST_INPLACE = 0, TC_DST_GLOBAL = 1, TC_DST_IMAGE = 2 };

enum { TC_ACT_NONE = 0, TC_ACT_NOCS_HEADS = 1, TC_ACT_JOINT_HEADS = 2 };

struct ChainStep {
    TcLayer L;
    int dst;        // TC_DST_INPLACE: output becomes the next step's operand; TC_DST_GLOBAL: rows go to `out`, operand kept;
                    // TC_DST_IMAGE: as GLOBAL, but `out` receives the rows as the fp16 hi/lo operand image of a following
                    // gemm_img_launch (GemmImgArgs::Aimg; rows_total * N * 4 bytes, ldo unused)
    float *out;     // [rows_total][ldo] f32; with TC_DST_INPLACE and out != NULL the rows are written as well
    int ldo;
    int act;        // TC_DST_GLOBAL only: TC_ACT_*_HEADS = the step is a packed head layer; its activations
                    // (lib/architecture.py:122-139, 150-157) run in the epilogue and the results go straight to
                    // ChainTcArgs::pred (out may be NULL)
};

struct ChainTcArgs {
    const float *X1;          // [rows_total][C1]
    const float *X2;          // [rows_total][C2] appended after X1 (or NULL)
    int C1, C2;
    long rows_per_cloud;      // for the per-cloud bias of step 0
    const float *bias0;       // NULL, or [clouds][bias0_stride] bias of step 0
    long bias0_stride;
    ChainStep S[8];
    int nsteps;
    int pool_S;               // > 0 (warp-specialised kernel only): the last step is max-pooled over groups of pool_S consecutive
                              // rows (multiple of 32, ReLU output) into S[last].out = [rows_total / pool_S][N]
    // Fused feature propagation (pointnet_util.py:206-236): with fp_points2 != NULL the X1 part of a row is
    // three_interpolate(fp_points2, fp_idx, fp_w) evaluated in the gather (X1 itself is ignored); (fp_idx, fp_w) are the
    // stage's three_nn indices and inverse-distance weights (ancsh_three_nn_tables_impl).
    const float *fp_points2;  // [clouds][fp_m2][C1]
    int fp_m2;
    const int *fp_idx;        // [rows_total][3]
    const float *fp_w;        // [rows_total][3]
    ancsh_pred_t pred;        // destination of the TC_ACT_* steps
    int n_parts, mixed;
    int kmax8;                // filled by the launcher
    uint32_t tmem_cols;
};

// number of hi*hi accumulators a layer's k range is split over (net_tc.cu, "Numerics")
__host__ __device__ inline int tc_num_acc(int K, int max_acc)
{
    const int g = (K + 143) / 144;
    return g < max_acc ? (g < 1 ? 1 : g) : max_acc;
}

int chain_tc2_launch(const ChainTcArgs &a, long rows_total, cudaStream_t st);   // net_tc2.cu

// ---- streaming GEMM (layers whose K or N do not fit the operand-resident chain: layer3 / group_all) -------------
// Image of a [rows][K] operand as the chain's TC_DST_IMAGE steps write it and gemm_img_launch reads it: per 128-row tile
// [K/8][hi | lo][128][8] fp16, tile stride K * 512 bytes -- 4 bytes per element, the size of the f32 rows.
struct GemmImgArgs {
    const void *Aimg;         // [rows_total / 128] tiles
    TcLayer L;                // K = image width (multiple of 16), N multiple of 128, ReLU
    float *out;               // [rows_total / pool_S][N] max over each group of pool_S consecutive rows
    int pool_S;               // multiple of 32
};
int gemm_img_launch(const GemmImgArgs &a, long rows_total, cudaStream_t st);
