// mlp_simt.cuh -- shared-memory-resident row-tile MLP building block on the FP32 CUDA cores.
//
// A 256-thread CTA owns a tile of TM rows whose activations stay in shared memory (row-major,
// leading dimension C_pad+4 so that the two rows a warp touches per A-fragment load fall into
// different banks).  dense_layer() multiplies the tile by one BN-folded, zero-padded weight matrix
// streamed from L2 through a cp.async double buffer (KC x NC k-slices), each thread accumulating a
// (TM/16) x (NC/16) register tile, and hands every finished NC-wide column pass to an epilogue functor
// (store to shared memory for the next layer, store to global, or max-pool over the group axis).
#pragma once
#include "common.cuh"

namespace mlp {

constexpr int NT = 256;  // threads per CTA
constexpr int KC = 16;   // k-slice depth of one weight tile
constexpr int WS_FLOATS = 2 * KC * 128;

template <int NC>
__device__ __forceinline__ void load_w_tile(float *Ws, const float *__restrict__ Wg, int ldw, int k0, int n0)
{
    constexpr int V = KC * NC / 4;
#pragma unroll
    for (int v = threadIdx.x; v < V; v += NT) {
        int r = v / (NC / 4), c4 = v - r * (NC / 4);
        cp_async16(Ws + r * NC + c4 * 4, Wg + (size_t)(k0 + r) * ldw + n0 + c4 * 4);
    }
}

// row / column owned by this thread: 4-wide strips, second strip 64 further
template <int TM>
__device__ __forceinline__ int thr_row(int i)
{
    const int ty = threadIdx.x >> 4;
    return (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4);
}
__device__ __forceinline__ int thr_col4(int h) { return h * 64 + (threadIdx.x & 15) * 4; }

template <int TM, int NC, class Epi>
__device__ __forceinline__ void dense_layer(const float *Xs, int ldx, const ancsh_layer_t &L, float *Wsm, Epi &&epi)
{
    constexpr int TR = TM / 16, TCN = NC / 16, CH = TCN / 4;
    const int nk = L.cin_pad / KC;
    const float *__restrict__ Wg = L.W;
    const int ldw = L.cout_pad;
    for (int n0 = 0; n0 < L.cout_pad; n0 += NC) {
        float acc[TR][TCN];
#pragma unroll
        for (int i = 0; i < TR; ++i)
#pragma unroll
            for (int j = 0; j < TCN; ++j) acc[i][j] = 0.f;
        __syncthreads();  // producers of Xs and previous readers of Wsm are done
        load_w_tile<NC>(Wsm, Wg, ldw, 0, n0);
        cp_async_commit();
        for (int kt = 0; kt < nk; ++kt) {
            if (kt + 1 < nk) {
                load_w_tile<NC>(Wsm + ((kt + 1) & 1) * KC * NC, Wg, ldw, (kt + 1) * KC, n0);
                cp_async_commit();
                cp_async_wait<1>();
            } else {
                cp_async_wait<0>();
            }
            __syncthreads();
            const float *Ws = Wsm + (kt & 1) * KC * NC;
#pragma unroll
            for (int k4 = 0; k4 < KC / 4; ++k4) {
                float4 a[TR];
#pragma unroll
                for (int i = 0; i < TR; ++i)
                    a[i] = *reinterpret_cast<const float4 *>(Xs + thr_row<TM>(i) * ldx + kt * KC + k4 * 4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float bv[TCN];
#pragma unroll
                    for (int h = 0; h < CH; ++h) {
                        float4 t = *reinterpret_cast<const float4 *>(Ws + (k4 * 4 + kk) * NC + thr_col4(h));
                        bv[h * 4 + 0] = t.x; bv[h * 4 + 1] = t.y; bv[h * 4 + 2] = t.z; bv[h * 4 + 3] = t.w;
                    }
#pragma unroll
                    for (int i = 0; i < TR; ++i) {
                        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
                        for (int j = 0; j < TCN; ++j) acc[i][j] = fmaf(av, bv[j], acc[i][j]);
                    }
                }
            }
            __syncthreads();
        }
        epi(n0, acc);
    }
}

// ---- epilogues -----------------------------------------------------------------------------------
// store bias+ReLU'd tile into a row-major shared-memory buffer (input of the next layer)
template <int TM, int NC>
struct EpiSmem {
    float *Ys;
    int ldy;
    const float *bias;
    int relu;
    __device__ __forceinline__ void operator()(int n0, float (&acc)[TM / 16][NC / 16]) const
    {
        constexpr int TR = TM / 16, CH = NC / 64;
#pragma unroll
        for (int h = 0; h < CH; ++h) {
            const int c = n0 + thr_col4(h);
            const float4 bb = ldg4(bias + c);
#pragma unroll
            for (int i = 0; i < TR; ++i) {
                float4 v;
                v.x = acc[i][h * 4 + 0] + bb.x; v.y = acc[i][h * 4 + 1] + bb.y;
                v.z = acc[i][h * 4 + 2] + bb.z; v.w = acc[i][h * 4 + 3] + bb.w;
                if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4 *>(Ys + thr_row<TM>(i) * ldy + c) = v;
            }
        }
    }
};

// store bias+ReLU'd tile to global rows (out + row*ldo), requires cout == cout_pad
template <int TM, int NC>
struct EpiGlobal {
    float *out;  // already offset to the tile's first row
    int ldo;
    const float *bias;
    int relu;
    __device__ __forceinline__ void operator()(int n0, float (&acc)[TM / 16][NC / 16]) const
    {
        constexpr int TR = TM / 16, CH = NC / 64;
#pragma unroll
        for (int h = 0; h < CH; ++h) {
            const int c = n0 + thr_col4(h);
            const float4 bb = ldg4(bias + c);
#pragma unroll
            for (int i = 0; i < TR; ++i) {
                float4 v;
                v.x = acc[i][h * 4 + 0] + bb.x; v.y = acc[i][h * 4 + 1] + bb.y;
                v.z = acc[i][h * 4 + 2] + bb.z; v.w = acc[i][h * 4 + 3] + bb.w;
                if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4 *>(out + (size_t)thr_row<TM>(i) * ldo + c) = v;
            }
        }
    }
};

// max-pool over groups of S consecutive rows (S % 8 == 0) into a ZERO-INITIALISED global buffer.
// The pooled layer always ends in ReLU (pointnet_util.py:124-134), so values are >= 0 and the integer
// ordering of their bit patterns equals the float ordering: atomicMax on int is exact.
template <int TM, int NC>
struct EpiPool {
    float *out;     // (groups, cout) of this cloud
    int cout;       // leading dimension of out
    const float *bias;
    long row0;      // first row of the tile inside the cloud
    int S;
    __device__ __forceinline__ void operator()(int n0, float (&acc)[TM / 16][NC / 16]) const
    {
        constexpr int RH = TM / 64, CH = NC / 64;
        const int ty = threadIdx.x >> 4;
#pragma unroll
        for (int rh = 0; rh < RH; ++rh) {
            const long g = (row0 + rh * 64 + ty * 4) / S;
#pragma unroll
            for (int h = 0; h < CH; ++h) {
                const int c = n0 + thr_col4(h);
                const float4 bb = ldg4(bias + c);
                const float bbv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = fmaxf(fmaxf(acc[rh * 4 + 0][h * 4 + j], acc[rh * 4 + 1][h * 4 + j]),
                                    fmaxf(acc[rh * 4 + 2][h * 4 + j], acc[rh * 4 + 3][h * 4 + j]));
                    v = fmaxf(v + bbv[j], 0.f);
                    v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, 16));
                    if ((ty & 1) == 0) atomicMax(reinterpret_cast<int *>(out + (size_t)g * cout + c + j), __float_as_int(v));
                }
            }
        }
    }
};

}  // namespace mlp
