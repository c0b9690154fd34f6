// net_tc.cuh -- argument structs of the tensor-core (tcgen05) network kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

struct TcLayer {
    const __half *Wimg;   // [K/8][2 (hi,lo)][N][8] fp16: W^T pre-split and pre-tiled (weights.tc_image)
    const float *bias;           // [N] f32 (BN folded)
    int K, N, relu;              // K = cin_pad (multiple of 16), N = cout_pad (64 / 128 / 256)
};

struct SaTcArgs {
    const float *xyz, *points, *new_xyz;
    const int *idx;
    float *out;
    TcLayer L[3];
    int n, m, S, C;
    int kmax8, nmax;             // filled by the launcher
    uint32_t tmem_cols;
    int terms;                   // products per k-step: 3 = hi*hi + hi*lo + lo*hi (default), 4 adds lo*lo
};

int sa_tc_launch(const SaTcArgs &a, int B, cudaStream_t st);

// ---- generic row-tile chain (feature propagation / heads) ------------------------------------------------------
enum { TC_DST_INPLACE = 0, TC_DST_GLOBAL = 1 };

struct ChainStep {
    TcLayer L;
    int dst;        // TC_DST_INPLACE: output becomes the next step's operand; TC_DST_GLOBAL: rows go to `out`, operand kept
    float *out;     // [rows_total][ldo] f32; with TC_DST_INPLACE and out != NULL the rows are written as well
    int ldo;
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
    int kmax8, nmax;          // filled by the launcher
    uint32_t tmem_cols;
};

int chain_tc_launch(const ChainTcArgs &a, long rows_total, cudaStream_t st);
