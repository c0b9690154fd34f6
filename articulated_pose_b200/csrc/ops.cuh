// ops.cuh -- device helpers shared by the op-level kernels (ops.cu) and the fused network kernels (net.cu).
#pragma once
#include "common.cuh"

// Squared distance exactly as the reference CPU op evaluates it (tf_interpolate.cpp:73):
// f32, ((dx*dx + dy*dy) + dz*dz), NO fma contraction, (known - query) operand order.
__device__ __forceinline__ float nn_dist_unfused(float x2, float y2, float z2, float x1, float y1, float z1)
{
    float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// tf_interpolate.cpp:121: p1*w1 + p2*w2 + p3*w3, f32, left to right, no fma.
__device__ __forceinline__ float interp3_unfused(float p1, float p2, float p3, float w1, float w2, float w3)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(p1, w1), __fmul_rn(p2, w2)), __fmul_rn(p3, w3));
}

// Running three smallest (distance, index) pairs with the reference's strict '<' chain
// (tf_interpolate.cpp:74-89): among equal distances the earlier-inserted (lower) index ranks first.
// The reference keeps doubles initialised to 1e40; with f32 candidates an f32 +inf sentinel is
// equivalent (a candidate that overflowed to +inf is never inserted in either formulation, and
// (float)1e40 == +inf is what the reference writes out for missing neighbours).
struct Best3 {
    float d1, d2, d3;
    int i1, i2, i3;
    __device__ __forceinline__ void init()
    {
        d1 = d2 = d3 = __int_as_float(0x7f800000);
        i1 = i2 = i3 = 0;
    }
    __device__ __forceinline__ void insert(float d, int k)
    {
        if (d < d1) {
            d3 = d2; i3 = i2; d2 = d1; i2 = i1; d1 = d; i1 = k;
        } else if (d < d2) {
            d3 = d2; i3 = i2; d2 = d; i2 = k;
        } else if (d < d3) {
            d3 = d; i3 = k;
        }
    }
    // merge a triple that covers strictly HIGHER indices than this one
    __device__ __forceinline__ void merge_higher(const Best3 &o)
    {
        insert(o.d1, o.i1);
        insert(o.d2, o.i2);
        insert(o.d3, o.i3);
    }
};

// pointnet_util.py:219-222: dist=max(dist,1e-10); norm=sum(1/dist); weight=(1/dist)/norm (IEEE f32 division)
__device__ __forceinline__ void three_weights(float d1, float d2, float d3, float &w1, float &w2, float &w3)
{
    d1 = fmaxf(d1, 1e-10f); d2 = fmaxf(d2, 1e-10f); d3 = fmaxf(d3, 1e-10f);
    float r1 = __fdiv_rn(1.0f, d1), r2 = __fdiv_rn(1.0f, d2), r3 = __fdiv_rn(1.0f, d3);
    float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
    w1 = __fdiv_rn(r1, norm); w2 = __fdiv_rn(r2, norm); w3 = __fdiv_rn(r3, norm);
}

int ancsh_fps_impl(int b, int n, int m, const float *xyz, int *idx, float *new_xyz, cudaStream_t st);
int ancsh_fps2_impl(int b, int n, int m, const float *xyz, int *idx, float *new_xyz, int m2, int *idx2, float *new_xyz2,
                    cudaStream_t st);
int ancsh_ball_query_impl(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                          int *pts_cnt, cudaStream_t st);
// (idx, weight) tables of a feature-propagation stage: three_nn(xyz1 queries (b,n,3), xyz2 known (b,m,3)) + the
// inverse-distance weights, both (b,n,3)
int ancsh_three_nn_tables_impl(int b, int n, int m, const float *xyz1, const float *xyz2, int *idx, float *weight, cudaStream_t st);
