// net.cu -- ANCSH network forward (PointNet++ trunk + heads) as fused row-tile kernels.
//
//   sa_kernel   : [ball-query indices] -> gather(xyz - centroid, features) -> 3 x (1x1 conv + BN + ReLU)
//                 -> max over nsample        (pointnet_util.py:29-63, 94-161; the grouped tensor that the
//                 reference materialises in HBM -- tf_grouping_g.cu:40-57 -- never leaves shared memory)
//   fp_kernel   : three_nn -> inverse-distance weights -> three_interpolate -> concat skip -> MLP
//                 (pointnet_util.py:206-236); with HEADS also fc1 + all ANCSH heads + activations + gocs
//                 (architectures.py:89-93, lib/architecture.py:98-159, 195-208)
//
// Arithmetic: f32 FMA accumulation in natural k order on the CUDA cores (this file is the exact-f32 path).
#include <stdlib.h>
#include "common.cuh"
#include "ops.cuh"
#include "mlp_simt.cuh"
#include "net_tc.cuh"

using namespace mlp;

// ------------------------------------------------------------------------------------------------
// dispatch helpers (NC = 64 for 64-wide layers, 128 otherwise)
// ------------------------------------------------------------------------------------------------
template <int TM>
__device__ __forceinline__ void layer_to_smem(const float *Xs, int ldx, const ancsh_layer_t &L, const float *bias,
                                              float *Ys, int ldy, float *Wsm)
{
    if (L.cout_pad == 64)
        dense_layer<TM, 64>(Xs, ldx, L, Wsm, EpiSmem<TM, 64>{Ys, ldy, bias, L.relu});
    else
        dense_layer<TM, 128>(Xs, ldx, L, Wsm, EpiSmem<TM, 128>{Ys, ldy, bias, L.relu});
}
template <int TM>
__device__ __forceinline__ void layer_to_global(const float *Xs, int ldx, const ancsh_layer_t &L, const float *bias,
                                                float *out, int ldo, float *Wsm)
{
    if (L.cout_pad == 64)
        dense_layer<TM, 64>(Xs, ldx, L, Wsm, EpiGlobal<TM, 64>{out, ldo, bias, L.relu});
    else
        dense_layer<TM, 128>(Xs, ldx, L, Wsm, EpiGlobal<TM, 128>{out, ldo, bias, L.relu});
}
template <int TM>
__device__ __forceinline__ void layer_to_pool(const float *Xs, int ldx, const ancsh_layer_t &L, const float *bias,
                                              float *out, long row0, int S, float *Wsm)
{
    if (L.cout_pad == 64)
        dense_layer<TM, 64>(Xs, ldx, L, Wsm, EpiPool<TM, 64>{out, L.cout, bias, row0, S});
    else
        dense_layer<TM, 128>(Xs, ldx, L, Wsm, EpiPool<TM, 128>{out, L.cout, bias, row0, S});
}

// ================================================================================================
// Set abstraction
// ================================================================================================
struct SaArgs {
    const float *xyz;      // (B,n,3) dataset coordinates
    const float *points;   // (B,n,C) dataset features (C may be 0)
    const float *new_xyz;  // (B,m,3) centroids; NULL = no centring (group_all, pointnet_util.py:80-88)
    const int *idx;        // (B,m,S) ball-query result; NULL = identity (group_all)
    float *out;            // (B,m,cout) zero-initialised
    ancsh_layer_t L[3];
    int n, m, S, C;
    int ldA, ldB;
};

template <int TM>
__global__ void __launch_bounds__(NT) sa_kernel(const SaArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float *bufA = smem;
    float *bufB = bufA + TM * a.ldA;
    float *Wsm = bufB + TM * a.ldB;

    const int tile = blockIdx.x, b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = a.C, cin_pad = a.L[0].cin_pad;
    const long row0 = (long)tile * TM;

    // ---- gather the tile: channel order [features(C), xyz(3), zero pad] --------------------------
    for (int r = warp; r < TM; r += NT / 32) {
        const long R = row0 + r;
        const int g = (int)(R / a.S);
        const int id = a.idx ? __ldg(a.idx + (size_t)b * a.m * a.S + R) : (int)R;
        float *xr = bufA + r * a.ldA;
        if (C > 0) {
            const float *prow = a.points + ((size_t)b * a.n + id) * C;
            for (int c4 = lane; c4 < C / 4; c4 += 32) *reinterpret_cast<float4 *>(xr + c4 * 4) = ldg4(prow + c4 * 4);
        }
        for (int c = C + lane; c < cin_pad; c += 32) {
            float v = 0.f;
            if (c < C + 3) {
                v = __ldg(a.xyz + ((size_t)b * a.n + id) * 3 + (c - C));
                if (a.new_xyz) v = __fsub_rn(v, __ldg(a.new_xyz + ((size_t)b * a.m + g) * 3 + (c - C)));   // :53
            }
            xr[c] = v;
        }
    }
    // (dense_layer starts with __syncthreads)
    layer_to_smem<TM>(bufA, a.ldA, a.L[0], a.L[0].b, bufB, a.ldB, Wsm);
    layer_to_smem<TM>(bufB, a.ldB, a.L[1], a.L[1].b, bufA, a.ldA, Wsm);
    layer_to_pool<TM>(bufA, a.ldA, a.L[2], a.L[2].b, a.out + (size_t)b * a.m * a.L[2].cout, row0, a.S, Wsm);
}

template <int TM>
static int sa_launch(const SaArgs &a0, int B, cudaStream_t st)
{
    SaArgs a = a0;
    const long rows = (long)a.m * a.S;
    if (rows % TM != 0 || a.S % 8 != 0 || a.C % 4 != 0) return ANCSH_ERR_UNSUPPORTED;
    if (a.L[2].cout != a.L[2].cout_pad || !a.L[2].relu) return ANCSH_ERR_INVALID_ARG;
    if (a.L[0].cin != a.C + 3) return ANCSH_ERR_INVALID_ARG;
    if (a.L[1].cin_pad != a.L[0].cout_pad || a.L[2].cin_pad != a.L[1].cout_pad) return ANCSH_ERR_INVALID_ARG;
    a.ldA = (a.L[0].cin_pad > a.L[1].cout_pad ? a.L[0].cin_pad : a.L[1].cout_pad) + 4;
    a.ldB = a.L[0].cout_pad + 4;
    size_t smem = ((size_t)TM * (a.ldA + a.ldB) + WS_FLOATS) * sizeof(float);
    if (smem > 227 * 1024) return ANCSH_ERR_UNSUPPORTED;
    ANCSH_CUDA(cudaFuncSetAttribute(sa_kernel<TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ANCSH_CUDA(cudaFuncSetAttribute(sa_kernel<TM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    ANCSH_CUDA(cudaMemsetAsync(a.out, 0, (size_t)B * a.m * a.L[2].cout * sizeof(float), st));
    dim3 grid((unsigned)(rows / TM), B);
    sa_kernel<TM><<<grid, NT, smem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

// A/B switches for profiling (read once): ANCSH_SA_LEAN_OFF = generic chain kernel for layer1 / layer2,
// ANCSH_BALL_FUSED_OFF = separate ball-query kernel in front of the lean kernel
static bool sa_lean_off()
{
    static const bool v = getenv("ANCSH_SA_LEAN_OFF") != nullptr;
    return v;
}
static bool fps_prefix_off()
{
    static const bool v = getenv("ANCSH_FPS_PREFIX_OFF") != nullptr;
    return v;
}
static bool ball_fused_off()
{
    static const bool v = getenv("ANCSH_BALL_FUSED_OFF") != nullptr;
    return v;
}

// tensor-core variant of a set-abstraction stage.  `idx_ready`: s.idx already holds the ball-query result (shared
// geometry); otherwise the layer-specialised kernel (net_lean.cu) runs the ball query itself and leaves a copy of the
// indices in s.idx / cnt for the second network of the pipeline, or -- shapes it does not cover -- the separate
// ball-query kernel runs in front of the generic chain kernel (net_tc2.cu).  Errors propagate; no other fallback.
static int sa_tc_from(const ancsh_net_t *net, const SaArgs &s, int B, float radius, int *idx_rw, int *cnt_rw, bool idx_ready,
                      cudaStream_t st)
{
    if (s.L[2].cout != s.L[2].cout_pad) return ANCSH_ERR_INVALID_ARG;
    TcLayer L[3];
    for (int l = 0; l < 3; ++l) {
        L[l].Wimg = reinterpret_cast<const __half *>(s.L[l].W_tc);
        L[l].bias = s.L[l].b; L[l].K = s.L[l].cin_pad; L[l].N = s.L[l].cout_pad; L[l].relu = s.L[l].relu;
        L[l].has_bias_step = net->tc_bias_step;
        L[l].descale = s.L[l].tc_descale;
    }
    int rc;
    if (!sa_lean_off() && idx_rw) {
        SaLeanArgs2 t{};
        t.xyz = s.xyz; t.points = s.points; t.new_xyz = s.new_xyz; t.out = s.out;
        t.n = s.n; t.m = s.m; t.S = s.S; t.C = s.C; t.radius = radius;
        t.w0_host = s.C == 0 ? net->sa1_conv0_host : nullptr;
        for (int l = 0; l < 3; ++l) t.L[l] = L[l];
        const bool fused = !idx_ready && !ball_fused_off();
        if (fused) { t.idx_in = nullptr; t.idx_out = idx_rw; t.cnt_out = cnt_rw; }
        else {
            if (!idx_ready && (rc = ancsh_ball_query_impl(B, s.n, s.m, radius, s.S, s.xyz, s.new_xyz, idx_rw, cnt_rw, st))) return rc;
            idx_ready = true;
            t.idx_in = idx_rw;
        }
        rc = sa_lean_launch(t, B, st);
        if (rc != ANCSH_ERR_UNSUPPORTED) return rc;
    }
    if (idx_rw && !idx_ready && (rc = ancsh_ball_query_impl(B, s.n, s.m, radius, s.S, s.xyz, s.new_xyz, idx_rw, cnt_rw, st))) return rc;
    SaTcArgs t{};
    t.xyz = s.xyz; t.points = s.points; t.new_xyz = s.new_xyz; t.idx = s.idx; t.out = s.out;
    t.n = s.n; t.m = s.m; t.S = s.S; t.C = s.C;
    t.W0 = s.L[0].W;
    for (int l = 0; l < 3; ++l) t.L[l] = L[l];
    return sa_tc2_launch(t, B, st);
}

// ================================================================================================
// fa_layer1's global-feature term: the single "known" point of FP1 makes three_interpolate a broadcast
// (weights 1,0,0 -- tf_interpolate.cpp:60-103 with m=1), so conv_0 splits into a per-cloud bias
//   cb[b][o] = b[o] + sum_k l3[b][k] * W[k][o]          (rows of W that multiply the interpolated part)
// plus the skip-feature rows handled by fp_kernel.
// ================================================================================================
__global__ void __launch_bounds__(256) cloud_bias_kernel(const float *__restrict__ feat, int cin,
                                                         const float *__restrict__ W, const float *__restrict__ bias,
                                                         int cout_pad, float *__restrict__ out)
{
    extern __shared__ float s_f[];
    const int b = blockIdx.x;
    for (int k = threadIdx.x; k < cin; k += blockDim.x) s_f[k] = feat[(size_t)b * cin + k];
    __syncthreads();
    for (int o = threadIdx.x; o < cout_pad; o += blockDim.x) {
        // the FMA chain is serial in k (fixed summation order); the loads are not: 32 of them in flight per thread
        // (4 in flight left the kernel waiting on L2 latency: 0.11 ms for a 0.13 GFLOP matvec batch, now 0.06 ms and
        // L2-bandwidth bound -- every block streams the whole 1 MB matrix; grouping 4 clouds per block to cut that
        // traffic was measured slower, 64 blocks are too few to hide the load latency)
        float acc = 0.f;
        const float *w = W + o;
        int k = 0;
        for (; k + 32 <= cin; k += 32) {
            float wv[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) wv[i] = __ldg(w + (size_t)(k + i) * cout_pad);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc = fmaf(s_f[k + i], wv[i], acc);
        }
        for (; k < cin; ++k) acc = fmaf(s_f[k], __ldg(w + (size_t)k * cout_pad), acc);
        out[(size_t)b * cout_pad + o] = acc + bias[o];
    }
}

// Tiled variant for the usual shapes (cin % 128 == 0, cout_pad % 64 == 0): a block owns CB_G clouds x 64 columns, its 256
// threads = 64 columns x 4 quarters of the k range.  Every weight is loaded once per 8 clouds (the one-cloud-per-block
// kernel above streams the whole matrix per cloud: 256 MB of L2 reads per 256 clouds, 0.060 ms), the features come from
// shared memory as 16-byte broadcasts.  The four partial sums of a column are added in a fixed order.
constexpr int CB_G = 8;
__global__ void __launch_bounds__(256) cloud_bias_tiled_kernel(const float *__restrict__ feat, int cin, int nclouds,
                                                               const float *__restrict__ W, const float *__restrict__ bias,
                                                               int cout_pad, float *__restrict__ out)
{
    extern __shared__ __align__(16) float s_f[];                 // [CB_G][cin], then [4][CB_G][64] partial sums
    float *s_part = s_f + (size_t)CB_G * cin;
    const int b0 = blockIdx.x * CB_G, c0 = blockIdx.y * 64;
    const int col = threadIdx.x & 63, kq = threadIdx.x >> 6;
    // the CB_G feature rows are consecutive in memory: one linear float4 copy, 8 loads in flight per thread
    {
        const int rows = min(CB_G, nclouds - b0);
        const float4 *src = reinterpret_cast<const float4 *>(feat + (size_t)b0 * cin);
        float4 *dst = reinterpret_cast<float4 *>(s_f);
        const int n4 = rows * cin / 4, all4 = CB_G * cin / 4;
        for (int i0 = threadIdx.x; i0 < all4; i0 += 256 * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * 256;
                v[u] = i < n4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * 256;
                if (i < all4) dst[i] = v[u];
            }
        }
    }
    __syncthreads();
    const int kn = cin / 4, k0 = kq * kn;
    float acc[CB_G];
#pragma unroll
    for (int g = 0; g < CB_G; ++g) acc[g] = 0.f;
    const float *w = W + (size_t)k0 * cout_pad + c0 + col;
    // cin % 128 == 0: two groups of 16 weights per trip, the second group's loads fly under the first group's FMAs
    for (int k = 0; k < kn; k += 32) {
        float wa[16], wb[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) wa[i] = __ldg(w + (size_t)(k + i) * cout_pad);
#pragma unroll
        for (int i = 0; i < 16; ++i) wb[i] = __ldg(w + (size_t)(k + 16 + i) * cout_pad);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int g = 0; g < CB_G; ++g) {
                    const float4 f = *reinterpret_cast<const float4 *>(s_f + (size_t)g * cin + k0 + k + 16 * half + 4 * q);
                    const float *wv = half ? wb : wa;
                    acc[g] = fmaf(f.x, wv[4 * q], acc[g]);
                    acc[g] = fmaf(f.y, wv[4 * q + 1], acc[g]);
                    acc[g] = fmaf(f.z, wv[4 * q + 2], acc[g]);
                    acc[g] = fmaf(f.w, wv[4 * q + 3], acc[g]);
                }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < CB_G; ++g) s_part[(kq * CB_G + g) * 64 + col] = acc[g];
    __syncthreads();
    for (int i = threadIdx.x; i < CB_G * 64; i += 256) {
        const int g = i >> 6, c = i & 63;
        if (b0 + g < nclouds)
            out[(size_t)(b0 + g) * cout_pad + c0 + c] =
                ((s_part[(0 * CB_G + g) * 64 + c] + s_part[(1 * CB_G + g) * 64 + c]) + s_part[(2 * CB_G + g) * 64 + c]) +
                s_part[(3 * CB_G + g) * 64 + c] + bias[c0 + c];
    }
}

static int cloud_bias_launch(int B, const float *feat, const ancsh_layer_t &G, float *out, cudaStream_t st)
{
    const size_t smem = ((size_t)CB_G * G.cin + 4 * CB_G * 64) * sizeof(float);
    if (G.cin % 128 == 0 && G.cout_pad % 64 == 0 && smem <= 200 * 1024) {
        if (smem > 48 * 1024)
            ANCSH_CUDA(cudaFuncSetAttribute(cloud_bias_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cloud_bias_tiled_kernel<<<dim3((B + CB_G - 1) / CB_G, G.cout_pad / 64), 256, smem, st>>>(feat, G.cin, B, G.W, G.b,
                                                                                                 G.cout_pad, out);
    } else {
        cloud_bias_kernel<<<B, 256, G.cin * sizeof(float), st>>>(feat, G.cin, G.W, G.b, G.cout_pad, out);
    }
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

// ================================================================================================
// Feature propagation (+ optional fc1 / heads)
// ================================================================================================
struct FpArgs {
    const float *xyz1;     // (B,n1,3) query points = rows
    const float *xyz2;     // (B,m2,3) known points; NULL = no interpolated part
    const float *points2;  // (B,m2,C2)
    const float *skip;     // (B,n1,C1)
    int n1, m2, C2, C1;
    ancsh_layer_t L[3];
    int nl;
    const float *bias0;    // bias of L[0]; per cloud when bias0_stride != 0
    long bias0_stride;
    float *out;            // (B,n1,cout_last)  (unused with HEADS)
    // HEADS only
    ancsh_layer_t fc1, nocs_heads, fc3[2], joint_heads;
    ancsh_pred_t pred;
    int K, mixed;
    int ldA, ldB, ldC;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ void copy_tile_out(float *dst, int w, const float *src, int ld, int off, int TM)
{
    if (!dst) return;
    for (int e = threadIdx.x; e < TM * w; e += NT) {
        int r = e / w, c = e - r * w;
        dst[e] = src[r * ld + off + c];
    }
}

template <int TM, bool HEADS>
__global__ void __launch_bounds__(NT) fp_kernel(const FpArgs a)
{
    extern __shared__ __align__(16) float smem[];
    float *bufA = smem;
    float *bufB = bufA + TM * a.ldA;
    float *bufC = bufB + TM * a.ldB;                 // HEADS only (ldC == 0 otherwise)
    float *Wsm = bufC + TM * a.ldC;
    float *s_known = Wsm + WS_FLOATS;                // m2*3
    float *s_w = s_known + a.m2 * 3;                 // TM*3
    int *s_i = reinterpret_cast<int *>(s_w + TM * 3);  // TM*3

    const int tile = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long row0 = (long)tile * TM;
    const int cin_pad = a.L[0].cin_pad;

    // ---- three_nn + weights (pointnet_util.py:217-222) ---------------------------------------------
    if (a.xyz2) {
        constexpr int PARTS = NT / TM;
        for (int i = tid; i < a.m2 * 3; i += NT) s_known[i] = __ldg(a.xyz2 + (size_t)b * a.m2 * 3 + i);
        __syncthreads();
        const int row = tid / PARTS, part = tid % PARTS;
        const float *q = a.xyz1 + ((size_t)b * a.n1 + row0 + row) * 3;
        const float x1 = __ldg(q), y1 = __ldg(q + 1), z1 = __ldg(q + 2);
        const int chunk = (a.m2 + PARTS - 1) / PARTS;
        const int k0 = part * chunk, k1 = min(a.m2, k0 + chunk);
        Best3 best;
        best.init();
        for (int k = k0; k < k1; ++k)
            best.insert(nn_dist_unfused(s_known[k * 3], s_known[k * 3 + 1], s_known[k * 3 + 2], x1, y1, z1), k);
#pragma unroll
        for (int off = 1; off < PARTS; off <<= 1) {
            Best3 o;
            o.d1 = __shfl_xor_sync(0xFFFFFFFFu, best.d1, off); o.i1 = __shfl_xor_sync(0xFFFFFFFFu, best.i1, off);
            o.d2 = __shfl_xor_sync(0xFFFFFFFFu, best.d2, off); o.i2 = __shfl_xor_sync(0xFFFFFFFFu, best.i2, off);
            o.d3 = __shfl_xor_sync(0xFFFFFFFFu, best.d3, off); o.i3 = __shfl_xor_sync(0xFFFFFFFFu, best.i3, off);
            if ((part & off) == 0) {
                best.merge_higher(o);
            } else {
                o.merge_higher(best);
                best = o;
            }
        }
        if (part == 0) {
            float w1, w2, w3;
            three_weights(best.d1, best.d2, best.d3, w1, w2, w3);
            s_w[row * 3 + 0] = w1; s_w[row * 3 + 1] = w2; s_w[row * 3 + 2] = w3;
            s_i[row * 3 + 0] = best.i1; s_i[row * 3 + 1] = best.i2; s_i[row * 3 + 2] = best.i3;
        }
        __syncthreads();
    }
    // ---- build the tile: [interpolated(C2), skip(C1), zero pad] (pointnet_util.py:223-229) ---------
    for (int r = warp; r < TM; r += NT / 32) {
        float *xr = bufA + r * a.ldA;
        if (a.xyz2) {
            const float w1 = s_w[r * 3], w2 = s_w[r * 3 + 1], w3 = s_w[r * 3 + 2];
            const float *p1 = a.points2 + ((size_t)b * a.m2 + s_i[r * 3 + 0]) * a.C2;
            const float *p2 = a.points2 + ((size_t)b * a.m2 + s_i[r * 3 + 1]) * a.C2;
            const float *p3 = a.points2 + ((size_t)b * a.m2 + s_i[r * 3 + 2]) * a.C2;
            for (int c4 = lane; c4 < a.C2 / 4; c4 += 32) {
                const float4 u = ldg4(p1 + c4 * 4), v = ldg4(p2 + c4 * 4), w = ldg4(p3 + c4 * 4);
                float4 o;
                o.x = interp3_unfused(u.x, v.x, w.x, w1, w2, w3);
                o.y = interp3_unfused(u.y, v.y, w.y, w1, w2, w3);
                o.z = interp3_unfused(u.z, v.z, w.z, w1, w2, w3);
                o.w = interp3_unfused(u.w, v.w, w.w, w1, w2, w3);
                *reinterpret_cast<float4 *>(xr + c4 * 4) = o;
            }
        }
        const float *srow = a.skip + ((size_t)b * a.n1 + row0 + r) * a.C1;
        for (int c = a.C2 + lane; c < cin_pad; c += 32) xr[c] = (c - a.C2 < a.C1) ? __ldg(srow + (c - a.C2)) : 0.f;
    }

    const float *bias0 = a.bias0 + (size_t)b * a.bias0_stride;
    if (!HEADS) {
        // two-layer chain: A -> B -> global
        layer_to_smem<TM>(bufA, a.ldA, a.L[0], bias0, bufB, a.ldB, Wsm);
        layer_to_global<TM>(bufB, a.ldB, a.L[1], a.L[1].b, a.out + ((size_t)b * a.n1 + row0) * a.L[1].cout, a.L[1].cout,
                            Wsm);
    } else {
        // fa_layer3 (3 layers) -> fc1 -> heads
        layer_to_smem<TM>(bufA, a.ldA, a.L[0], bias0, bufB, a.ldB, Wsm);
        layer_to_smem<TM>(bufB, a.ldB, a.L[1], a.L[1].b, bufA, a.ldA, Wsm);
        layer_to_smem<TM>(bufA, a.ldA, a.L[2], a.L[2].b, bufB, a.ldB, Wsm);
        layer_to_smem<TM>(bufB, a.ldB, a.fc1, a.fc1.b, bufA, a.ldA, Wsm);                    // net
        if (a.pred.net) {
            __syncthreads();
            copy_tile_out(a.pred.net + ((size_t)b * a.n1 + row0) * a.fc1.cout, a.fc1.cout, bufA, a.ldA, 0, TM);
        }
        layer_to_smem<TM>(bufA, a.ldA, a.nocs_heads, a.nocs_heads.b, bufC, a.ldC, Wsm);      // raw nocs_net outputs
        layer_to_smem<TM>(bufA, a.ldA, a.fc3[0], a.fc3[0].b, bufB, a.ldB, Wsm);
        layer_to_smem<TM>(bufB, a.ldB, a.fc3[1], a.fc3[1].b, bufA, a.ldA, Wsm);
        layer_to_smem<TM>(bufA, a.ldA, a.joint_heads, a.joint_heads.b, bufB, a.ldC, Wsm);    // raw joint_net outputs
        __syncthreads();
        // ---- activations (lib/architecture.py:122-139, 150-157), one thread per row and head group --
        const int K = a.K;
        if (tid < TM) {
            float *x = bufC + tid * a.ldC;
            float mx = x[0];
            for (int k = 1; k < K; ++k) mx = fmaxf(mx, x[k]);
            float s = 0.f;
            for (int k = 0; k < K; ++k) { float e = expf(x[k] - mx); x[k] = e; s += e; }
            for (int k = 0; k < K; ++k) x[k] = x[k] / s;                                   // W softmax
            for (int k = K; k < 4 * K; ++k) x[k] = sigmoidf_(x[k]);                        // nocs
            if (a.mixed) {
                for (int k = 4 * K; k < 5 * K; ++k) x[k] = sigmoidf_(x[k]);                // scale
                for (int k = 5 * K; k < 8 * K; ++k) x[k] = tanhf(x[k]);                    // translation
                x[8 * K] = sigmoidf_(x[8 * K]);                                            // confidence
                for (int k = 0; k < 3 * K; ++k)                                            // gocs :154-157
                    x[8 * K + 1 + k] = __fadd_rn(__fmul_rn(x[K + k], x[4 * K + k / 3]), x[5 * K + k]);
            } else {
                x[4 * K] = sigmoidf_(x[4 * K]);
            }
        } else if (tid < 2 * TM) {
            float *x = bufB + (tid - TM) * a.ldC;
            for (int k = 0; k < 6; ++k) x[k] = tanhf(x[k]);                                // joint_axis, unitvec
            x[6] = sigmoidf_(x[6]);                                                        // heatmap
            float mx = fmaxf(x[7], fmaxf(x[8], x[9]));
            float e0 = expf(x[7] - mx), e1 = expf(x[8] - mx), e2 = expf(x[9] - mx);
            float s = e0 + e1 + e2;
            x[7] = e0 / s; x[8] = e1 / s; x[9] = e2 / s;                                   // index softmax
        }
        __syncthreads();
        const size_t p0 = (size_t)b * a.n1 + row0;
        const ancsh_pred_t &o = a.pred;
        copy_tile_out(o.W ? o.W + p0 * K : nullptr, K, bufC, a.ldC, 0, TM);
        copy_tile_out(o.nocs_per_point ? o.nocs_per_point + p0 * 3 * K : nullptr, 3 * K, bufC, a.ldC, K, TM);
        if (a.mixed) {
            copy_tile_out(o.global_scale ? o.global_scale + p0 * K : nullptr, K, bufC, a.ldC, 4 * K, TM);
            copy_tile_out(o.global_translation ? o.global_translation + p0 * 3 * K : nullptr, 3 * K, bufC, a.ldC, 5 * K, TM);
            copy_tile_out(o.confi_per_point ? o.confi_per_point + p0 : nullptr, 1, bufC, a.ldC, 8 * K, TM);
            copy_tile_out(o.gocs_per_point ? o.gocs_per_point + p0 * 3 * K : nullptr, 3 * K, bufC, a.ldC, 8 * K + 1, TM);
        } else {
            copy_tile_out(o.confi_per_point ? o.confi_per_point + p0 : nullptr, 1, bufC, a.ldC, 4 * K, TM);
        }
        copy_tile_out(o.joint_axis_per_point ? o.joint_axis_per_point + p0 * 3 : nullptr, 3, bufB, a.ldC, 0, TM);
        copy_tile_out(o.unitvec_per_point ? o.unitvec_per_point + p0 * 3 : nullptr, 3, bufB, a.ldC, 3, TM);
        copy_tile_out(o.heatmap_per_point ? o.heatmap_per_point + p0 : nullptr, 1, bufB, a.ldC, 6, TM);
        copy_tile_out(o.index_per_point ? o.index_per_point + p0 * 3 : nullptr, 3, bufB, a.ldC, 7, TM);
    }
}

template <int TM, bool HEADS>
static int fp_launch(const FpArgs &a0, int B, cudaStream_t st)
{
    FpArgs a = a0;
    if (a.n1 % TM != 0 || a.C2 % 4 != 0) return ANCSH_ERR_UNSUPPORTED;
    if (a.L[0].cin != a.C2 + a.C1) return ANCSH_ERR_INVALID_ARG;
    int maxA = a.L[0].cin_pad, maxB = a.L[0].cout_pad;
    if (HEADS) {
        if (a.nl != 3) return ANCSH_ERR_INVALID_ARG;
        const int wa[] = {a.L[1].cout_pad, a.fc1.cout_pad, a.fc3[1].cout_pad};
        const int wb[] = {a.L[2].cout_pad, a.fc3[0].cout_pad};
        for (int v : wa) maxA = v > maxA ? v : maxA;
        for (int v : wb) maxB = v > maxB ? v : maxB;
        if (a.nocs_heads.cout_pad != 64 || a.joint_heads.cout_pad != 64) return ANCSH_ERR_INVALID_ARG;
        if (11 * a.K + 1 > 64 || a.K < 1) return ANCSH_ERR_UNSUPPORTED;
        a.ldC = 64 + 4;
    } else {
        if (a.nl != 2 || a.L[1].cout != a.L[1].cout_pad) return ANCSH_ERR_INVALID_ARG;
        a.ldC = 0;
    }
    a.ldA = maxA + 4;
    a.ldB = maxB + 4;
    size_t smem = ((size_t)TM * (a.ldA + a.ldB + a.ldC) + WS_FLOATS + (size_t)a.m2 * 3 + (size_t)TM * 6) * sizeof(float);
    if (smem > 227 * 1024) return ANCSH_ERR_UNSUPPORTED;
    ANCSH_CUDA(cudaFuncSetAttribute(fp_kernel<TM, HEADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ANCSH_CUDA(cudaFuncSetAttribute(fp_kernel<TM, HEADS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    dim3 grid(a.n1 / TM, B);
    fp_kernel<TM, HEADS><<<grid, NT, smem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

static TcLayer tc_layer(const ancsh_layer_t &l, int has_bias_step = 0)
{
    TcLayer t;
    t.Wimg = reinterpret_cast<const __half *>(l.W_tc);
    t.bias = l.b; t.K = l.cin_pad; t.N = l.cout_pad; t.relu = l.relu;
    t.has_bias_step = has_bias_step;
    t.descale = l.tc_descale;
    return t;
}

// ================================================================================================
// plan + forward
// ================================================================================================
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int ancsh_net_plan(const ancsh_net_t *net, int B, int N, ancsh_ws_layout_t *L)
{
    if (!net || !L || B <= 0 || N <= 0) return ANCSH_ERR_INVALID_ARG;
    if (N % 128 != 0) return ANCSH_ERR_UNSUPPORTED;
    const size_t m1 = net->npoint1, m2 = net->npoint2, s1 = net->nsample1, s2 = net->nsample2, b = B;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align256(off + bytes); return o; };
    L->fps_idx1 = take(b * m1 * 4);
    L->l1_xyz = take(b * m1 * 3 * 4);
    L->fps_idx2 = take(b * m2 * 4);
    L->l2_xyz = take(b * m2 * 3 * 4);
    L->ball_idx1 = take(b * m1 * s1 * 4);
    L->ball_cnt1 = take(b * m1 * 4);
    L->ball_idx2 = take(b * m2 * s2 * 4);
    L->ball_cnt2 = take(b * m2 * 4);
    L->l1_points = take(b * m1 * net->sa1[2].cout * 4);
    L->l2_points = take(b * m2 * net->sa2[2].cout * 4);
    L->l3_points = take(b * net->sa3[2].cout * 4);
    L->fp1_bias = take(b * net->fp1_global.cout_pad * 4);
    L->l2_points_fp = take(b * m2 * net->fp1[1].cout * 4);
    L->l1_points_fp = take(b * m1 * net->fp2[1].cout * 4);
    L->interp3 = take(net->use_tensor_cores ? b * m2 * net->sa3[1].cout_pad * 4 : 0);
    L->raw_heads = take(0);
    L->nn_idx2 = take(net->use_tensor_cores ? b * m1 * 3 * 4 : 0);
    L->nn_w2 = take(net->use_tensor_cores ? b * m1 * 3 * 4 : 0);
    L->nn_idx3 = take(net->use_tensor_cores ? b * N * 3 * 4 : 0);
    L->nn_w3 = take(net->use_tensor_cores ? b * N * 3 * 4 : 0);
    L->total_bytes = off;
    return ANCSH_OK;
}

// geo_net / geo_ws != NULL: sampling and grouping geometry (FPS indices, level xyz, ball-query indices) is read from the
// workspace of a forward of `geo_net` over the same P instead of being recomputed (ancsh_net_forward_shared).
static int net_forward_impl(const ancsh_net_t *net, int B, int N, const float *P, void *workspace, size_t workspace_bytes,
                            const ancsh_net_t *geo_net, const void *geo_ws, const ancsh_pred_t *pred,
                            void *const *stage_events, void *stream)
{
    if (!net || !P || !workspace || !pred) return ANCSH_ERR_INVALID_ARG;
    if (B > 65535) return ANCSH_ERR_UNSUPPORTED;
    ancsh_ws_layout_t L;
    int rc = ancsh_net_plan(net, B, N, &L);
    if (rc != ANCSH_OK) return rc;
    if (workspace_bytes < L.total_bytes) return ANCSH_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    char *gws = ws;                       // where the geometry lives
    ancsh_ws_layout_t GL = L;
    const bool shared = geo_net != nullptr && geo_ws != nullptr;
    if (shared) {
        if (geo_net->npoint1 != net->npoint1 || geo_net->npoint2 != net->npoint2 || geo_net->nsample1 != net->nsample1 ||
            geo_net->nsample2 != net->nsample2 || geo_net->radius1 != net->radius1 || geo_net->radius2 != net->radius2 ||
            geo_net->use_tensor_cores != net->use_tensor_cores)       // the three_nn tables exist on the tensor-core path only
            return ANCSH_ERR_INVALID_ARG;
        if ((rc = ancsh_net_plan(geo_net, B, N, &GL)) != ANCSH_OK) return rc;
        gws = (char *)const_cast<void *>(geo_ws);
    }
    int *fps1 = (int *)(gws + GL.fps_idx1), *fps2 = (int *)(gws + GL.fps_idx2);
    float *l1_xyz = (float *)(gws + GL.l1_xyz), *l2_xyz = (float *)(gws + GL.l2_xyz);
    int *bidx1 = (int *)(gws + GL.ball_idx1), *bcnt1 = (int *)(gws + GL.ball_cnt1);
    int *bidx2 = (int *)(gws + GL.ball_idx2), *bcnt2 = (int *)(gws + GL.ball_cnt2);
    float *l1_points = (float *)(ws + L.l1_points), *l2_points = (float *)(ws + L.l2_points);
    float *l3_points = (float *)(ws + L.l3_points), *fp1_bias = (float *)(ws + L.fp1_bias);
    float *l2_fp = (float *)(ws + L.l2_points_fp), *l1_fp = (float *)(ws + L.l1_points_fp);
    const int m1 = net->npoint1, m2 = net->npoint2;
    int stage = 0;
#define STAGE_MARK()                                                                        \
    do {                                                                                    \
        if (stage_events) ANCSH_CUDA(cudaEventRecord((cudaEvent_t)stage_events[stage], st)); \
        ++stage;                                                                            \
    } while (0)

    // sampling (pointnet_util.py:47) -- level 2 samples the level-1 centroids
    STAGE_MARK();
    // (level 2 is the prefix of level 1 whenever the reference's tie rule orders ties by index: one launch for both)
    const bool fps_prefix = m1 <= 512 && m2 <= m1 && !fps_prefix_off();
    if (!shared && (rc = ancsh_fps2_impl(B, N, m1, P, fps1, l1_xyz, fps_prefix ? m2 : 0, fps2, l2_xyz, st))) return rc;
    STAGE_MARK();
    if (!shared && !fps_prefix && (rc = ancsh_fps_impl(B, m1, m2, l1_xyz, fps2, l2_xyz, st))) return rc;
    STAGE_MARK();

    // layer1 (tensor-core path: the ball query runs inside the set-abstraction kernel, stage BALL1 is empty)
    if (!shared && !net->use_tensor_cores &&
        (rc = ancsh_ball_query_impl(B, N, m1, net->radius1, net->nsample1, P, l1_xyz, bidx1, bcnt1, st))) return rc;
    STAGE_MARK();
    {
        SaArgs a{};
        a.xyz = P; a.points = nullptr; a.new_xyz = l1_xyz; a.idx = bidx1; a.out = l1_points;
        a.L[0] = net->sa1[0]; a.L[1] = net->sa1[1]; a.L[2] = net->sa1[2];
        a.n = N; a.m = m1; a.S = net->nsample1; a.C = 0;
        if ((rc = net->use_tensor_cores ? sa_tc_from(net, a, B, net->radius1, bidx1, bcnt1, shared, st) : sa_launch<128>(a, B, st))) return rc;
    }
    // layer2
    STAGE_MARK();
    if (!shared && !net->use_tensor_cores &&
        (rc = ancsh_ball_query_impl(B, m1, m2, net->radius2, net->nsample2, l1_xyz, l2_xyz, bidx2, bcnt2, st))) return rc;
    STAGE_MARK();
    {
        SaArgs a{};
        a.xyz = l1_xyz; a.points = l1_points; a.new_xyz = l2_xyz; a.idx = bidx2; a.out = l2_points;
        a.L[0] = net->sa2[0]; a.L[1] = net->sa2[1]; a.L[2] = net->sa2[2];
        a.n = m1; a.m = m2; a.S = net->nsample2; a.C = net->sa1[2].cout;
        if ((rc = net->use_tensor_cores ? sa_tc_from(net, a, B, net->radius2, bidx2, bcnt2, shared, st) : sa_launch<128>(a, B, st))) return rc;
    }
    // layer3 (group_all)
    STAGE_MARK();
    {
        SaArgs a{};
        a.xyz = l2_xyz; a.points = l2_points; a.new_xyz = nullptr; a.idx = nullptr; a.out = l3_points;
        a.L[0] = net->sa3[0]; a.L[1] = net->sa3[1]; a.L[2] = net->sa3[2];
        a.n = m2; a.m = 1; a.S = m2; a.C = net->sa2[2].cout;
        if (net->use_tensor_cores && net->sa3[0].W_tc && net->sa3[2].W_tc) {
            // group_all over the B*npoint2 rows.  conv0 (in place) + conv1 as one warp-specialised chain whose last epilogue
            // writes conv1's rows directly as the fp16 hi/lo operand image of conv2 (workspace slot `interp3`, 4 bytes per
            // element like the f32 rows); conv2 (K = 512: a resident operand would need 256 KB of shared memory) streams
            // image and weights through a TMA ring, with the max over the cloud's npoint2 rows in its epilogue.
            void *img = ws + L.interp3;
            if (m2 % 32 != 0 || ((long)B * m2) % 128 != 0 || net->sa3[2].cout != net->sa3[2].cout_pad) return ANCSH_ERR_UNSUPPORTED;
            if (net->sa3[2].cin_pad != net->sa3[1].cout_pad) return ANCSH_ERR_INVALID_ARG;
            ChainTcArgs c{};
            c.X1 = l2_points; c.C1 = net->sa2[2].cout; c.X2 = l2_xyz; c.C2 = 3; c.rows_per_cloud = m2;
            c.S[0].L = tc_layer(net->sa3[0]); c.S[0].dst = TC_DST_INPLACE;
            c.S[1].L = tc_layer(net->sa3[1]); c.S[1].dst = TC_DST_IMAGE; c.S[1].out = (float *)img;
            c.nsteps = 2;
            if ((rc = chain_tc2_launch(c, (long)B * m2, st))) return rc;
            GemmImgArgs g{};
            g.Aimg = img; g.L = tc_layer(net->sa3[2]); g.out = l3_points; g.pool_S = m2;
            if ((rc = gemm_img_launch(g, (long)B * m2, st))) return rc;
        } else if ((rc = sa_launch<64>(a, B, st))) return rc;
    }
    // fa_layer1
    STAGE_MARK();
    {
        const ancsh_layer_t &G = net->fp1_global;
        if (G.cin != net->sa3[2].cout || G.cout_pad != net->fp1[0].cout_pad) return ANCSH_ERR_INVALID_ARG;
        if ((rc = cloud_bias_launch(B, l3_points, G, fp1_bias, st))) return rc;
        FpArgs a{};
        a.xyz1 = l2_xyz; a.xyz2 = nullptr; a.points2 = nullptr; a.skip = l2_points;
        a.n1 = m2; a.m2 = 0; a.C2 = 0; a.C1 = net->sa2[2].cout;
        a.L[0] = net->fp1[0]; a.L[1] = net->fp1[1]; a.nl = 2;
        a.bias0 = fp1_bias; a.bias0_stride = G.cout_pad;
        a.out = l2_fp;
        if (net->use_tensor_cores && net->fp1[0].W_tc && net->fp1[1].W_tc) {
            ChainTcArgs c{};
            c.X1 = l2_points; c.C1 = net->sa2[2].cout; c.X2 = nullptr; c.C2 = 0;
            c.rows_per_cloud = m2; c.bias0 = fp1_bias; c.bias0_stride = G.cout_pad;
            c.S[0].L = tc_layer(net->fp1[0]); c.S[0].dst = TC_DST_INPLACE; c.S[0].out = nullptr; c.S[0].ldo = 0;
            c.S[1].L = tc_layer(net->fp1[1]); c.S[1].dst = TC_DST_GLOBAL; c.S[1].out = l2_fp; c.S[1].ldo = net->fp1[1].cout;
            c.nsteps = 2;
            if (net->fp1[1].cout != net->fp1[1].cout_pad) return ANCSH_ERR_INVALID_ARG;
            if ((rc = chain_tc2_launch(c, (long)B * m2, st))) return rc;
        } else if ((rc = fp_launch<64, false>(a, B, st))) return rc;
    }
    // fa_layer2
    STAGE_MARK();
    {
        FpArgs a{};
        a.xyz1 = l1_xyz; a.xyz2 = l2_xyz; a.points2 = l2_fp; a.skip = l1_points;
        a.n1 = m1; a.m2 = m2; a.C2 = net->fp1[1].cout; a.C1 = net->sa1[2].cout;
        a.L[0] = net->fp2[0]; a.L[1] = net->fp2[1]; a.nl = 2;
        a.bias0 = net->fp2[0].b; a.bias0_stride = 0;
        a.out = l1_fp;
        if (net->use_tensor_cores && net->fp2[0].W_tc && net->fp2[1].W_tc) {
            // rows [three_interpolate(l2_fp)(256) | l1_points(128)] are built inside the chain kernel's gather from the
            // stage's (idx, weight) tables (24 bytes per row): no interpolated tensor in HBM
            const int C2 = net->fp1[1].cout;
            if (m1 % 128 != 0 || C2 % 8 != 0) return ANCSH_ERR_UNSUPPORTED;
            int *nn_i = (int *)(gws + GL.nn_idx2);
            float *nn_w = (float *)(gws + GL.nn_w2);
            ChainTcArgs c{};
            c.X1 = nullptr; c.C1 = C2; c.X2 = l1_points; c.C2 = net->sa1[2].cout; c.rows_per_cloud = m1;
            c.bias0 = nullptr; c.bias0_stride = 0;
            c.fp_points2 = l2_fp; c.fp_m2 = m2; c.fp_idx = nn_i; c.fp_w = nn_w;
            if (!shared && (rc = ancsh_three_nn_tables_impl(B, m1, m2, l1_xyz, l2_xyz, nn_i, nn_w, st))) return rc;
            c.S[0].L = tc_layer(net->fp2[0]); c.S[0].dst = TC_DST_INPLACE; c.S[0].out = nullptr; c.S[0].ldo = 0;
            c.S[1].L = tc_layer(net->fp2[1]); c.S[1].dst = TC_DST_GLOBAL; c.S[1].out = l1_fp; c.S[1].ldo = net->fp2[1].cout;
            c.nsteps = 2;
            if (net->fp2[1].cout != net->fp2[1].cout_pad) return ANCSH_ERR_INVALID_ARG;
            if ((rc = chain_tc2_launch(c, (long)B * m1, st))) return rc;
        } else if ((rc = fp_launch<64, false>(a, B, st))) return rc;
    }
    // fa_layer3 + fc1 + heads
    STAGE_MARK();
    {
        FpArgs a{};
        a.xyz1 = P; a.xyz2 = l1_xyz; a.points2 = l1_fp; a.skip = P;
        a.n1 = N; a.m2 = m1; a.C2 = net->fp2[1].cout; a.C1 = 3;
        a.L[0] = net->fp3[0]; a.L[1] = net->fp3[1]; a.L[2] = net->fp3[2]; a.nl = 3;
        a.bias0 = net->fp3[0].b; a.bias0_stride = 0;
        a.fc1 = net->fc1; a.nocs_heads = net->nocs_heads; a.fc3[0] = net->fc3[0]; a.fc3[1] = net->fc3[1];
        a.joint_heads = net->joint_heads;
        a.pred = *pred; a.K = net->n_parts; a.mixed = net->mixed_pred;
        if (net->use_tensor_cores && net->fp3[0].W_tc && net->joint_heads.W_tc) {
            // fa_layer3 + fc1 + both head stacks as ONE launch: the rows [three_interpolate(l1_fp)(128) | xyz(3)] are built in
            // the gather, the head activations and gocs run in the epilogues of the two head layers, the predictions are
            // written once (no interpolated / raw-head tensors in HBM)
            const int C2 = net->fp2[1].cout;
            if (N % 128 != 0 || C2 % 8 != 0 || net->nocs_heads.cout_pad != 64 || net->joint_heads.cout_pad != 64)
                return ANCSH_ERR_UNSUPPORTED;
            int *nn_i = (int *)(gws + GL.nn_idx3);
            float *nn_w = (float *)(gws + GL.nn_w3);
            ChainTcArgs c{};
            c.X1 = nullptr; c.C1 = C2; c.X2 = P; c.C2 = 3; c.rows_per_cloud = N; c.bias0 = nullptr; c.bias0_stride = 0;
            c.fp_points2 = l1_fp; c.fp_m2 = m1; c.fp_idx = nn_i; c.fp_w = nn_w;
            if (!shared && (rc = ancsh_three_nn_tables_impl(B, N, m1, P, l1_xyz, nn_i, nn_w, st))) return rc;
            c.pred = *pred; c.n_parts = net->n_parts; c.mixed = net->mixed_pred;
            const ancsh_layer_t *seq[8] = {&net->fp3[0], &net->fp3[1], &net->fp3[2], &net->fc1, &net->nocs_heads,
                                           &net->fc3[0], &net->fc3[1], &net->joint_heads};
            for (int i = 0; i < 8; ++i) {
                c.S[i].L = tc_layer(*seq[i]);
                c.S[i].dst = TC_DST_INPLACE; c.S[i].out = nullptr; c.S[i].ldo = 0;
            }
            c.S[3].out = pred->net; c.S[3].ldo = net->fc1.cout;                       // optional copy of the trunk feature
            c.S[4].dst = TC_DST_GLOBAL; c.S[4].act = TC_ACT_NOCS_HEADS;               // nocs_net heads (operand `net` kept)
            c.S[7].dst = TC_DST_GLOBAL; c.S[7].act = TC_ACT_JOINT_HEADS;              // joint_net heads
            // the joint branch (fc3_0, fc3_1, joint heads: 3 of the 8 layers) is skipped when none of its four outputs is
            // requested -- e.g. the NPCS-baseline network of the pipeline, whose pose stage reads W and NOCS only
            const bool want_joint = pred->joint_axis_per_point || pred->unitvec_per_point || pred->heatmap_per_point ||
                                    pred->index_per_point;
            c.nsteps = want_joint ? 8 : 5;
            if ((rc = chain_tc2_launch(c, (long)B * N, st))) return rc;
        } else if ((rc = fp_launch<128, true>(a, B, st))) return rc;
    }
    STAGE_MARK();
#undef STAGE_MARK
    return ANCSH_OK;
}

extern "C" int ancsh_net_forward(const ancsh_net_t *net, int B, int N, const float *P, void *workspace,
                                 size_t workspace_bytes, const ancsh_pred_t *pred, void *const *stage_events,
                                 void *stream)
{
    return net_forward_impl(net, B, N, P, workspace, workspace_bytes, nullptr, nullptr, pred, stage_events, stream);
}

extern "C" int ancsh_net_forward_shared(const ancsh_net_t *net, int B, int N, const float *P, void *workspace,
                                        size_t workspace_bytes, const ancsh_net_t *geometry_net,
                                        const void *geometry_workspace, const ancsh_pred_t *pred,
                                        void *const *stage_events, void *stream)
{
    if (!geometry_net || !geometry_workspace) return ANCSH_ERR_INVALID_ARG;
    return net_forward_impl(net, B, N, P, workspace, workspace_bytes, geometry_net, geometry_workspace, pred, stage_events,
                            stream);
}

extern "C" int ancsh_event_create(void **event_out)
{
    if (!event_out) return ANCSH_ERR_INVALID_ARG;
    cudaEvent_t e;
    ANCSH_CUDA(cudaEventCreate(&e));
    *event_out = (void *)e;
    return ANCSH_OK;
}
extern "C" int ancsh_event_record(void *event, void *stream)
{
    ANCSH_CUDA(cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream));
    return ANCSH_OK;
}
extern "C" int ancsh_event_elapsed_ms(void *start, void *stop, float *ms_out)
{
    if (!ms_out) return ANCSH_ERR_INVALID_ARG;
    ANCSH_CUDA(cudaEventElapsedTime(ms_out, (cudaEvent_t)start, (cudaEvent_t)stop));
    return ANCSH_OK;
}
extern "C" int ancsh_event_destroy(void *event)
{
    ANCSH_CUDA(cudaEventDestroy((cudaEvent_t)event));
    return ANCSH_OK;
}
