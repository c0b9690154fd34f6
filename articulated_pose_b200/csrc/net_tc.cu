// net_tc.cu -- set-abstraction stage on the Blackwell tensor cores (tcgen05 + TMEM).
//
// Same fusion as sa_kernel (net.cu): ball-query indices -> gather(features, xyz - centroid) -> 3 x (1x1 conv + folded
// BN + ReLU) -> max over nsample, the grouped tensor never leaving the SM -- but every layer is a
// 128 x N x K tcgen05.mma contraction with the f32 accumulator in tensor memory:
//   * one CTA = 128 threads = 128 rows (thread t owns row t for the gather, for TMEM lane t in the epilogue and for
//     the write-back of the next layer's operand);
//   * activations live in shared memory as fp16 hi/lo operand images (tc_common.cuh), overwritten in place by the
//     epilogue of each layer; weights are pre-split / pre-tiled on the host into the same image layout and streamed
//     from L2 in 32-wide k slices with cp.async;
//   * 3 MMAs per k-step (hi*hi + hi*lo + lo*hi) give f32-class accuracy at fp16 tensor-core throughput;
//   * the last layer's epilogue max-pools each group with one redux.sync per column.
// Two CTAs per SM (<= 113 KB shared memory, <= 256 TMEM columns each) overlap one CTA's epilogue / weight fetch
// with the other's MMAs.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "net_tc.cuh"

namespace {

constexpr int TM = 128;      // rows per CTA == threads per CTA
constexpr int KSLICE = 32;   // k elements per weight slice (2 MMA k-steps)

__global__ void __launch_bounds__(TM) sa_tc_kernel(const SaTcArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *A_hi = smem;
    uint8_t *A_lo = A_hi + (size_t)a.kmax8 * 2048;
    uint8_t *Wst = A_lo + (size_t)a.kmax8 * 2048;
    uint64_t *bar = reinterpret_cast<uint64_t *>(Wst + (size_t)(KSLICE / 8) * 2 * a.nmax * 16);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, b = blockIdx.y;
    const long row0 = (long)tile * TM;

    if (warp == 0) tc::tmem_alloc(s_tmem, a.tmem_cols);
    if (tid == 0) tc::mbar_init(bar, 1);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    uint32_t phase = 0;

    // ---- gather row `tid` of the tile: channel order [features(C), xyz(3), zero pad] (pointnet_util.py:52-57) ----
    {
        const long R = row0 + tid;
        const int g = (int)(R / a.S);
        const int id = a.idx ? __ldg(a.idx + (size_t)b * a.m * a.S + R) : (int)R;
        const float *prow = a.points ? a.points + ((size_t)b * a.n + id) * a.C : nullptr;
        float rel[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __ldg(a.xyz + ((size_t)b * a.n + id) * 3 + c);
            if (a.new_xyz) v = __fsub_rn(v, __ldg(a.new_xyz + ((size_t)b * a.m + g) * 3 + c));
            rel[c] = v;
        }
        const int K0 = a.L[0].K;
        for (int kc = 0; kc < K0 / 8; ++kc) {
            float v[8];
            if (kc * 8 + 8 <= a.C) {
                const float4 p0 = ldg4(prow + kc * 8), p1 = ldg4(prow + kc * 8 + 4);
                v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = kc * 8 + i;
                    v[i] = c < a.C ? __ldg(prow + c) : (c < a.C + 3 ? rel[c - a.C] : 0.f);
                }
            }
            tc::store_split8(v, reinterpret_cast<uint4 *>(A_hi + (size_t)kc * 2048 + tid * 16),
                             reinterpret_cast<uint4 *>(A_lo + (size_t)kc * 2048 + tid * 16));
        }
    }
    tc::fence_proxy_async();

    const uint32_t a_hi0 = tc::smem_u32(A_hi), a_lo0 = tc::smem_u32(A_lo), w0 = tc::smem_u32(Wst);

    for (int l = 0; l < 3; ++l) {
        const TcLayer &L = a.L[l];
        const int N = L.N, nk16 = L.K / 16;
        const uint32_t idesc = tc::instr_desc_f16(TM, N);
        const uint32_t slab = 2u * N * 16u;                  // one group of 8 k: hi rows then lo rows
        for (int k16 = 0; k16 < nk16; k16 += KSLICE / 16) {
            const int steps = min(KSLICE / 16, nk16 - k16);
            const uint32_t bytes = (uint32_t)steps * 2u * slab;
            const uint8_t *src = reinterpret_cast<const uint8_t *>(L.Wimg) + (size_t)k16 * 2 * slab;
            for (uint32_t off = tid * 16; off < bytes; off += TM * 16) cp_async16(Wst + off, src + off);
            cp_async_commit();
            cp_async_wait<0>();
            tc::fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                for (int s = 0; s < steps; ++s) {
                    const int kk = k16 + s;
                    const uint64_t ah = tc::smem_desc(a_hi0 + (uint32_t)(2 * kk) * 2048u, 2048u, 128u);
                    const uint64_t al = tc::smem_desc(a_lo0 + (uint32_t)(2 * kk) * 2048u, 2048u, 128u);
                    const uint64_t bh = tc::smem_desc(w0 + (uint32_t)(2 * s) * slab, slab, 128u);
                    const uint64_t bl = tc::smem_desc(w0 + (uint32_t)(2 * s) * slab + (uint32_t)N * 16u, slab, 128u);
                    tc::mma_f16(tmem, ah, bh, idesc, kk > 0 ? 1u : 0u);
                    tc::mma_f16(tmem, ah, bl, idesc, 1u);
                    tc::mma_f16(tmem, al, bh, idesc, 1u);
                    if (a.terms >= 4) tc::mma_f16(tmem, al, bl, idesc, 1u);
                }
                tc::mma_commit(bar);
            }
            tc::mbar_wait(bar, phase);       // the slice buffer is free again / the accumulator is complete
            phase ^= 1;
        }
        tc::fence_after_sync();
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        if (l < 2) {
            // bias + ReLU, split to fp16 hi/lo, becomes the next layer's operand (in place)
            for (int c0 = 0; c0 < N; c0 += 32) {
                float v[32];
                tc::tmem_ld32(trow + c0, v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float x = v[q * 8 + i] + __ldg(L.bias + c0 + q * 8 + i);
                        w[i] = L.relu ? fmaxf(x, 0.f) : x;
                    }
                    const int kc = (c0 >> 3) + q;
                    tc::store_split8(w, reinterpret_cast<uint4 *>(A_hi + (size_t)kc * 2048 + tid * 16),
                                     reinterpret_cast<uint4 *>(A_lo + (size_t)kc * 2048 + tid * 16));
                }
            }
            tc::fence_proxy_async();
            tc::fence_before_sync();
            __syncthreads();
        } else {
            // bias + ReLU + max over the group's rows (pointnet_util.py:134).  Values are >= 0, so the unsigned
            // ordering of their bit patterns is the float ordering: one redux.sync per column.
            const long g = (row0 + warp * 32) / a.S;
            float *orow = a.out + ((size_t)b * a.m + g) * N;
            for (int c0 = 0; c0 < N; c0 += 32) {
                float v[32];
                tc::tmem_ld32(trow + c0, v);
                uint32_t keep = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float x = fmaxf(v[i] + __ldg(L.bias + c0 + i), 0.f);
                    const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(x));
                    if (lane == i) keep = mx;
                }
                if (a.S == 32) orow[c0 + lane] = __uint_as_float(keep);
                else atomicMax(reinterpret_cast<int *>(orow + c0 + lane), (int)keep);
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------
// Row-tile chain for the point-wise stages (fa_layer1, fa_layer3 + fc1 + all heads): same operand-in-place scheme as
// sa_tc_kernel, rows come straight from global memory, and every step either feeds the next one or writes rows out.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TM) chain_tc_kernel(const ChainTcArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *A_hi = smem;
    uint8_t *A_lo = A_hi + (size_t)a.kmax8 * 2048;
    uint8_t *Wst = A_lo + (size_t)a.kmax8 * 2048;
    uint64_t *bar = reinterpret_cast<uint64_t *>(Wst + (size_t)(KSLICE / 8) * 2 * a.nmax * 16);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const long R = (long)blockIdx.x * TM + tid;

    if (warp == 0) tc::tmem_alloc(s_tmem, a.tmem_cols);
    if (tid == 0) tc::mbar_init(bar, 1);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    uint32_t phase = 0;

    {
        const float *r1 = a.X1 + (size_t)R * a.C1;
        const float *r2 = a.X2 ? a.X2 + (size_t)R * a.C2 : nullptr;
        const int K0 = a.S[0].L.K;
        for (int kc = 0; kc < K0 / 8; ++kc) {
            float v[8];
            if (kc * 8 + 8 <= a.C1) {
                const float4 p0 = ldg4(r1 + kc * 8), p1 = ldg4(r1 + kc * 8 + 4);
                v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
            } else if (r2 && (a.C2 & 7) == 0 && kc * 8 >= a.C1 && kc * 8 + 8 <= a.C1 + a.C2) {
                const float4 p0 = ldg4(r2 + (kc * 8 - a.C1)), p1 = ldg4(r2 + (kc * 8 - a.C1) + 4);
                v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = kc * 8 + i;
                    v[i] = c < a.C1 ? __ldg(r1 + c) : (r2 && c < a.C1 + a.C2 ? __ldg(r2 + (c - a.C1)) : 0.f);
                }
            }
            tc::store_split8(v, reinterpret_cast<uint4 *>(A_hi + (size_t)kc * 2048 + tid * 16),
                             reinterpret_cast<uint4 *>(A_lo + (size_t)kc * 2048 + tid * 16));
        }
    }
    tc::fence_proxy_async();
    const uint32_t a_hi0 = tc::smem_u32(A_hi), a_lo0 = tc::smem_u32(A_lo), w0 = tc::smem_u32(Wst);

    for (int st = 0; st < a.nsteps; ++st) {
        const ChainStep &S = a.S[st];
        const TcLayer &L = S.L;
        const int N = L.N, nk16 = L.K / 16;
        const uint32_t idesc = tc::instr_desc_f16(TM, N);
        const uint32_t slab = 2u * N * 16u;
        const int G = tc_num_acc(L.K, a.max_acc);
        uint32_t started = 0;
        for (int k16 = 0; k16 < nk16; k16 += KSLICE / 16) {
            const int steps = min(KSLICE / 16, nk16 - k16);
            const uint32_t bytes = (uint32_t)steps * 2u * slab;
            const uint8_t *src = reinterpret_cast<const uint8_t *>(L.Wimg) + (size_t)k16 * 2 * slab;
            for (uint32_t off = tid * 16; off < bytes; off += TM * 16) cp_async16(Wst + off, src + off);
            cp_async_commit();
            cp_async_wait<0>();
            tc::fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                for (int s = 0; s < steps; ++s) {
                    const int kk = k16 + s;
                    const uint64_t ah = tc::smem_desc(a_hi0 + (uint32_t)(2 * kk) * 2048u, 2048u, 128u);
                    const uint64_t al = tc::smem_desc(a_lo0 + (uint32_t)(2 * kk) * 2048u, 2048u, 128u);
                    const uint64_t bh = tc::smem_desc(w0 + (uint32_t)(2 * s) * slab, slab, 128u);
                    const uint64_t bl = tc::smem_desc(w0 + (uint32_t)(2 * s) * slab + (uint32_t)N * 16u, slab, 128u);
                    const int g = kk * G / nk16;                       // accumulator of this k-step
                    const uint32_t d = tmem + (uint32_t)(g * a.acc_stride);
                    tc::mma_f16(d, ah, bh, idesc, (started >> g) & 1u);
                    started |= 1u << g;
                    tc::mma_f16(d, ah, bl, idesc, 1u);
                    tc::mma_f16(d, al, bh, idesc, 1u);
                }
                tc::mma_commit(bar);
            }
            tc::mbar_wait(bar, phase);
            phase ^= 1;
        }
        tc::fence_after_sync();
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const float *bias = (st == 0 && a.bias0) ? a.bias0 + (size_t)(R / a.rows_per_cloud) * a.bias0_stride : L.bias;
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            tc::tmem_ld32(trow + c0, v);
            for (int g = 1; g < G; ++g) {
                float u[32];
                tc::tmem_ld32(trow + g * a.acc_stride + c0, u);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += u[i];
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float x = v[i] + __ldg(bias + c0 + i);
                v[i] = L.relu ? fmaxf(x, 0.f) : x;
            }
            if (S.out) {
                float4 *o = reinterpret_cast<float4 *>(S.out + (size_t)R * S.ldo + c0);
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            if (S.dst == TC_DST_INPLACE) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = v[q * 8 + i];
                    const int kc = (c0 >> 3) + q;
                    tc::store_split8(w, reinterpret_cast<uint4 *>(A_hi + (size_t)kc * 2048 + tid * 16),
                                     reinterpret_cast<uint4 *>(A_lo + (size_t)kc * 2048 + tid * 16));
                }
            }
        }
        tc::fence_proxy_async();
        tc::fence_before_sync();
        __syncthreads();
    }
    if (warp == 0) tc::tmem_dealloc(tmem, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------
// Streaming GEMM: neither operand is resident.  Per 32-wide k slice every thread converts its row's 32 f32 inputs to
// the fp16 hi/lo images, the weight slice arrives by cp.async, 2 k-steps x 3 MMAs are issued.  blockIdx.y selects a
// chunk of <= 256 output columns.  Used for layer3 (259 -> 256 -> 512 -> 1024 on the 128 points of each cloud).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TM) gemm_tc_kernel(const GemmTcArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *A_hi = smem;                                   // (KSLICE/8) * 2048
    uint8_t *A_lo = A_hi + (KSLICE / 8) * 2048;
    uint8_t *Wst = A_lo + (KSLICE / 8) * 2048;
    uint64_t *bar = reinterpret_cast<uint64_t *>(Wst + (size_t)(KSLICE / 8) * 2 * a.nchunk * 16);
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long R = (long)blockIdx.x * TM + tid;
    const int n0 = blockIdx.y * a.nchunk;
    const int NC = min(a.nchunk, a.L.N - n0);

    if (warp == 0) tc::tmem_alloc(s_tmem, a.tmem_cols);
    if (tid == 0) tc::mbar_init(bar, 1);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    uint32_t phase = 0;

    const float *r1 = a.X1 + (size_t)R * a.C1;
    const float *r2 = a.X2 ? a.X2 + (size_t)R * a.C2 : nullptr;
    const uint32_t a_hi0 = tc::smem_u32(A_hi), a_lo0 = tc::smem_u32(A_lo), w0 = tc::smem_u32(Wst);
    const uint32_t idesc = tc::instr_desc_f16(TM, NC);
    const uint32_t gslab = 2u * a.L.N * 16u;                 // global image: [K/8][2][N][8]
    const uint32_t sslab = 2u * NC * 16u;                    // staged slice: [kc][2][NC][8]
    const int nk16 = a.L.K / 16;
    const int G = tc_num_acc(a.L.K, a.max_acc);
    uint32_t started = 0;
    for (int k16 = 0; k16 < nk16; k16 += KSLICE / 16) {
        const int steps = min(KSLICE / 16, nk16 - k16);
        // weights: for each of the steps*2 k-groups copy the hi and lo rows [n0, n0+NC)
        const uint8_t *src = reinterpret_cast<const uint8_t *>(a.L.Wimg);
        const int pieces = steps * 2 * 2 * NC;                // 16-byte pieces
        for (int p = tid; p < pieces; p += TM) {
            const int row = p % NC, part = (p / NC) & 1, kc = p / (2 * NC);
            cp_async16(Wst + (size_t)kc * sslab + (size_t)part * NC * 16 + row * 16,
                       src + (size_t)(k16 * 2 + kc) * gslab + (size_t)part * a.L.N * 16 + (size_t)(n0 + row) * 16);
        }
        cp_async_commit();
        // activations: this thread's row, k in [k16*16, k16*16 + steps*16)
        for (int kc = 0; kc < steps * 2; ++kc) {
            const int c0 = (k16 * 2 + kc) * 8;
            float v[8];
            if (c0 + 8 <= a.C1) {
                const float4 p0 = ldg4(r1 + c0), p1 = ldg4(r1 + c0 + 4);
                v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = c0 + i;
                    v[i] = c < a.C1 ? __ldg(r1 + c) : (r2 && c < a.C1 + a.C2 ? __ldg(r2 + (c - a.C1)) : 0.f);
                }
            }
            tc::store_split8(v, reinterpret_cast<uint4 *>(A_hi + (size_t)kc * 2048 + tid * 16),
                             reinterpret_cast<uint4 *>(A_lo + (size_t)kc * 2048 + tid * 16));
        }
        cp_async_wait<0>();
        tc::fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after_sync();
            for (int s = 0; s < steps; ++s) {
                const uint64_t ah = tc::smem_desc(a_hi0 + (uint32_t)(2 * s) * 2048u, 2048u, 128u);
                const uint64_t al = tc::smem_desc(a_lo0 + (uint32_t)(2 * s) * 2048u, 2048u, 128u);
                const uint64_t bh = tc::smem_desc(w0 + (uint32_t)(2 * s) * sslab, sslab, 128u);
                const uint64_t bl = tc::smem_desc(w0 + (uint32_t)(2 * s) * sslab + (uint32_t)NC * 16u, sslab, 128u);
                const int g = (k16 + s) * G / nk16;
                const uint32_t d = tmem + (uint32_t)(g * a.acc_stride);
                tc::mma_f16(d, ah, bh, idesc, (started >> g) & 1u);
                started |= 1u << g;
                tc::mma_f16(d, ah, bl, idesc, 1u);
                tc::mma_f16(d, al, bh, idesc, 1u);
            }
            tc::mma_commit(bar);
        }
        tc::mbar_wait(bar, phase);
        phase ^= 1;
    }
    tc::fence_after_sync();
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < NC; c0 += 32) {
        float v[32];
        tc::tmem_ld32(trow + c0, v);
        for (int g = 1; g < G; ++g) {
            float u[32];
            tc::tmem_ld32(trow + g * a.acc_stride + c0, u);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += u[i];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float x = v[i] + __ldg(a.L.bias + n0 + c0 + i);
            v[i] = a.L.relu ? fmaxf(x, 0.f) : x;
        }
        if (a.pool_S == 0) {
            float4 *o = reinterpret_cast<float4 *>(a.out + (size_t)R * a.ldo + n0 + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
            const long g = ((long)blockIdx.x * TM + warp * 32) / a.pool_S;
            float *orow = a.out + (size_t)g * a.L.N + n0 + c0;
            uint32_t keep = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(v[i]));
                if (lane == i) keep = mx;
            }
            if (a.pool_S == 32) orow[lane] = __uint_as_float(keep);
            else atomicMax(reinterpret_cast<int *>(orow + lane), (int)keep);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, a.tmem_cols);
}

}  // namespace

int sa_tc_launch(const SaTcArgs &a0, int B, cudaStream_t st)
{
    SaTcArgs a = a0;
    const long rows = (long)a.m * a.S;
    if (rows % TM != 0 || a.S % 32 != 0 || a.C % 8 != 0) return ANCSH_ERR_UNSUPPORTED;
    int kmax = 0, nmax = 0;
    for (int l = 0; l < 3; ++l) {
        if (!a.L[l].Wimg || a.L[l].K % 16 != 0 || a.L[l].N % 32 != 0 || a.L[l].N > 256) return ANCSH_ERR_INVALID_ARG;
        if (l > 0 && a.L[l].K != a.L[l - 1].N) return ANCSH_ERR_INVALID_ARG;
        kmax = a.L[l].K > kmax ? a.L[l].K : kmax;
        nmax = a.L[l].N > nmax ? a.L[l].N : nmax;
    }
    if (a.L[0].K < a.C + 3 || !a.L[2].relu) return ANCSH_ERR_INVALID_ARG;
    {
        const char *e = getenv("ANCSH_TC_TERMS");
        a.terms = e ? atoi(e) : 3;
    }
    a.kmax8 = kmax / 8;
    a.nmax = nmax;
    a.tmem_cols = nmax <= 32 ? 32 : nmax <= 64 ? 64 : nmax <= 128 ? 128 : 256;
    const size_t smem = (size_t)2 * a.kmax8 * 2048 + (size_t)(KSLICE / 8) * 2 * nmax * 16 + 16;
    if (smem > 113 * 1024) return ANCSH_ERR_UNSUPPORTED;
    ANCSH_CUDA(cudaFuncSetAttribute(sa_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ANCSH_CUDA(cudaFuncSetAttribute(sa_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (a.S != 32) ANCSH_CUDA(cudaMemsetAsync(a.out, 0, (size_t)B * a.m * a.L[2].N * sizeof(float), st));
    dim3 grid((unsigned)(rows / TM), B);
    sa_tc_kernel<<<grid, TM, smem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

int chain_tc_launch(const ChainTcArgs &a0, long rows_total, cudaStream_t st)
{
    ChainTcArgs a = a0;
    if (rows_total % TM != 0 || a.C1 % 8 != 0 || a.nsteps < 1 || a.nsteps > 8) return ANCSH_ERR_UNSUPPORTED;
    int kmax = 0, nmax = 0;
    for (int i = 0; i < a.nsteps; ++i) {
        const TcLayer &L = a.S[i].L;
        if (!L.Wimg || L.K % 16 != 0 || L.N % 32 != 0 || L.N > 256) return ANCSH_ERR_INVALID_ARG;
        if (a.S[i].dst == TC_DST_GLOBAL && !a.S[i].out) return ANCSH_ERR_INVALID_ARG;
        if (a.S[i].out && (a.S[i].ldo < L.N || a.S[i].ldo % 4 != 0)) return ANCSH_ERR_INVALID_ARG;
        kmax = L.K > kmax ? L.K : kmax;
        if (a.S[i].dst == TC_DST_INPLACE && L.N > kmax) kmax = L.N;
        nmax = L.N > nmax ? L.N : nmax;
    }
    if (a.S[0].L.K < a.C1 + (a.X2 ? a.C2 : 0)) return ANCSH_ERR_INVALID_ARG;
    a.kmax8 = kmax / 8;
    a.nmax = nmax;
    a.acc_stride = nmax <= 32 ? 32 : nmax <= 64 ? 64 : nmax <= 128 ? 128 : 256;
    const size_t smem = (size_t)2 * a.kmax8 * 2048 + (size_t)(KSLICE / 8) * 2 * nmax * 16 + 16;
    if (smem > 227 * 1024) return ANCSH_ERR_UNSUPPORTED;
    {
        // two CTAs per SM (<= 256 columns each) while shared memory allows it, else the whole 512-column TMEM
        const int budget = smem > 113 * 1024 ? 512 : 256;
        int need = 1;
        for (int i = 0; i < a.nsteps; ++i) {
            const int g = tc_num_acc(a.S[i].L.K, budget / a.acc_stride);
            need = g > need ? g : need;
        }
        a.max_acc = need;
        int cols = a.acc_stride * need;
        a.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    }
    ANCSH_CUDA(cudaFuncSetAttribute(chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ANCSH_CUDA(cudaFuncSetAttribute(chain_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    chain_tc_kernel<<<(unsigned)(rows_total / TM), TM, smem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

int gemm_tc_launch(const GemmTcArgs &a0, long rows_total, cudaStream_t st)
{
    GemmTcArgs a = a0;
    if (rows_total % TM != 0 || a.C1 % 8 != 0) return ANCSH_ERR_UNSUPPORTED;
    if (!a.L.Wimg || a.L.K % 16 != 0 || a.L.N % 32 != 0 || a.L.K < a.C1 + (a.X2 ? a.C2 : 0)) return ANCSH_ERR_INVALID_ARG;
    if (a.pool_S && (a.pool_S % 32 != 0 || !a.L.relu)) return ANCSH_ERR_INVALID_ARG;
    if (!a.pool_S && (a.ldo < a.L.N || a.ldo % 4 != 0)) return ANCSH_ERR_INVALID_ARG;
    a.nchunk = a.L.N <= 128 ? a.L.N : 128;               // 128-column chunks leave room for up to 4 accumulators
    if (a.L.N % a.nchunk != 0) return ANCSH_ERR_UNSUPPORTED;
    a.acc_stride = a.nchunk <= 32 ? 32 : a.nchunk <= 64 ? 64 : 128;
    a.max_acc = tc_num_acc(a.L.K, 256 / a.acc_stride);   // <= 256 columns per CTA: two CTAs per SM
    {
        const int cols = a.acc_stride * a.max_acc;
        a.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : 256;
    }
    const size_t smem = (size_t)2 * (KSLICE / 8) * 2048 + (size_t)(KSLICE / 8) * 2 * a.nchunk * 16 + 16;
    ANCSH_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ANCSH_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (a.pool_S && a.pool_S != 32)
        ANCSH_CUDA(cudaMemsetAsync(a.out, 0, (size_t)(rows_total / a.pool_S) * a.L.N * sizeof(float), st));
    dim3 grid((unsigned)(rows_total / TM), (unsigned)(a.L.N / a.nchunk));
    gemm_tc_kernel<<<grid, TM, smem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}
