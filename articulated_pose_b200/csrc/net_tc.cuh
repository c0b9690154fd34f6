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
