// net_lean.cu -- layer-specialised tcgen05 kernel for the two ball-query set-abstraction stages (layer1 / layer2):
//   [ball query] -> gather(features, xyz - centroid) -> 3 x (1x1 conv + folded BN + ReLU) -> max over nsample
// (pointnet_util.py:29-63, 94-161; tf_grouping_g.cu:3-57).  One launch per stage, nothing but the pooled (B,m,C_out)
// tensor leaves the SM (optionally the ball-query indices, for the second network of the pipeline / intermediates()).
//
// Same CTA shape as net_tc2.cu (128 rows = 8 worker warps + MMA warp + weight-producer warp, operand images resident in
// shared memory, accumulators in TMEM, weights streamed through a bulk-copy ring), but every worker-side loop is
// compile-time specialised and the epilogues are cut to the instruction counts the tensor pipe can hide:
//   * bias: folded into the GEMM.  The weight images carry one extra 16-deep k-step whose rows are (fp16(b), fp16(b -
//     fp16(b)), 0 ...) (weights.tc_image / ancsh_tc_image); its A operand is a constant "ones" slab.  No bias loads or
//     adds on the CUDA cores.
//   * ReLU + fp16 hi/lo split of an intermediate layer: hi = v & 0xFFFFE000 (f32 truncated to 11 significant bits,
//     exactly representable in fp16), lo = v - hi (exact), both packed with cvt.rn.relu.f16x2.f32 -- lo has the sign of
//     v, so the saturating conversion is the ReLU.  3 instructions per element instead of 7.
//   * last layer TRANSPOSED: D^T[c_out][row] = W^T (A operand, from the ring) x Act^T (B operand = the resident
//     activation image, same K-major layout).  A TMEM lane then holds one output channel and the nsample rows of a
//     centroid are consecutive COLUMNS: the max-pool is an in-thread max over tcgen05.ld registers (3-input FMNMX), no
//     warp reductions, and the pooled row is written with lanes on consecutive channels (coalesced).
//   * xyz-only first conv of layer1 (3 input channels) on the CUDA cores with its weights in the kernel parameter
//     (constant) bank: FFMA with constant operands, no weight loads.
//   * ball query fused into the gather: one warp per centroid, ballot + popc ordered compaction (bit-exact with
//     tf_grouping_g.cu:3-36: first nsample in index order, padded with the first hit), indices handed to the gathering
//     threads through shared memory.
// Numerics: unchanged from net_tc2.cu (hi*hi and cross terms in separate f32 TMEM accumulators, summed in the epilogue).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "tc_common.cuh"
#include "net_tc.cuh"

namespace {

constexpr int TM = 128;                 // rows per tile
constexpr int NWORK = 256;              // worker threads
constexpr int NTHR = NWORK + 64;        // + MMA warp + producer warp
constexpr int STAGE_BYTES = 8192;       // one ring stage = one 16-deep k-step of <= 128 weight rows (hi + lo)
constexpr int MAX_STAGES = 8;
constexpr int MAXU = 4;
constexpr uint32_t TMEM_COLS = 256;

struct LUnit {
    const uint8_t *img;       // weight image [K/8][hi|lo][Nfull][8] fp16, offset to the unit's first row
    uint32_t kstep_stride;    // bytes between consecutive k-steps in the image (4 * Nfull * 16)
    uint32_t piece_stride;    // bytes between the four (kc, hi|lo) blocks of a k-step (Nfull * 16)
    uint32_t piece_bytes;     // rows of this unit * 16
    int nk;                   // k-steps, including the bias step
    int bias;                 // the last k-step is the bias step (A = ones slab, hi*hi product only)
    int transposed;           // D^T = W^T x Act^T (weights are the A operand)
    uint32_t idesc;
};

struct SaLeanArgs {
    const float *xyz;         // (B,n,3)
    const float *points;      // (B,n,C) or NULL
    const float *new_xyz;     // (B,m,3)
    const int *idx_in;        // (B,m,S) precomputed ball-query indices (BALL == false)
    int *idx_out;             // BALL: optional copy of the indices, (B,m,S)
    int *cnt_out;             // BALL: optional pts_cnt (B,m)
    float *out;               // (B,m,N2)
    const float *bias_last;   // [N2] bias of the pooled layer
    int n, m;
    float radius;
    int tiles_per_cta;
    int nst;
    int nunits;
    LUnit U[MAXU];
    long long *trace;         // profiling aid (ANCSH_LEAN_TRACE): per-CTA clock64() stamps of the phase boundaries, or NULL
    float w0[4 * 64];         // xyz-only first conv: rows 0..2 = W[k][0..63], row 3 = bias
};

// slot layout of a CTA's trace record: [0..31] worker warp 0, [32..63] worker warp 4, [64..127] MMA warp, tiles 0 and 1 only
constexpr int TRACE_SLOTS = 128;
#define LEAN_TRACE(base, k)                                                                                              \
    do {                                                                                                                 \
        if (a.trace && lane == 0 && t < 2)                                                                               \
            a.trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * TRACE_SLOTS + (base) + t * ((base) == 64 ? 32 : 16) + (k)] = clock64(); \
    } while (0)

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void work_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK + 32) : "memory"); }
__device__ __forceinline__ void gather_sync() { asm volatile("bar.sync 2, %0;" ::"n"(NWORK) : "memory"); }

// {upper half, lower half} = {fp16(hi_elem), fp16(lo_elem)}, round to nearest, optionally clamped at 0 (ReLU)
template <bool RELU>
__device__ __forceinline__ uint32_t pack_f16x2(float upper, float lower)
{
    uint32_t d;
    if (RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
    else asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
    return d;
}
// 8 consecutive k of one row -> one 16-byte piece each of the hi and lo operand images
template <bool RELU>
__device__ __forceinline__ void split8_lean(const float (&v)[8], uint8_t *dst_hi, uint8_t *dst_lo)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        const float ha = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), hb = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
        h[i] = pack_f16x2<RELU>(hb, ha);
        l[i] = pack_f16x2<RELU>(b - hb, a - ha);
    }
    *reinterpret_cast<uint4 *>(dst_hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(dst_lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// epilogue of an in-place layer: this thread's row, columns [h*N/2, (h+1)*N/2): (hi*hi + cross) -> ReLU -> split -> operand
template <int N>
__device__ __forceinline__ void epi_inplace(uint32_t trow, int h, int r, uint8_t *A_hi, uint8_t *A_lo)
{
#pragma unroll
    for (int cc = 0; cc < N / 2; cc += 32) {
        const int c0 = h * (N / 2) + cc;
        uint32_t ra[32], rb[32];
        tc::tmem_ld32_issue(trow + (uint32_t)c0, ra);
        tc::tmem_ld32_issue(trow + (uint32_t)(N + c0), rb);
        tc::tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(ra[q * 8 + i]) + __uint_as_float(rb[q * 8 + i]);
            const int kc = (c0 >> 3) + q;
            split8_lean<true>(v, A_hi + (size_t)kc * 2048 + r * 16, A_lo + (size_t)kc * 2048 + r * 16);
        }
    }
}

// epilogue of a transposed pooled block: this thread's TMEM lane = one output channel, columns [h*64, h*64+64) = rows of
// 64 / S centroids.  out_c points at out[centroid 0 of the tile][channel of this lane]; ld = channels per centroid.
template <int S>
__device__ __forceinline__ void epi_pool_T(uint32_t trow, int h, float bias, float *out_c, int ld)
{
    static_assert(S == 32 || S == 64, "nsample 32 or 64");
#pragma unroll
    for (int g = 0; g < 64 / S; ++g) {
        float mx = -3.0e38f;
#pragma unroll
        for (int cc = 0; cc < S; cc += 32) {
            const int c0 = h * 64 + g * S + cc;
            uint32_t ra[32], rb[32];
            tc::tmem_ld32_issue(trow + (uint32_t)c0, ra);
            tc::tmem_ld32_issue(trow + (uint32_t)(TM + c0), rb);
            tc::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 2)
                mx = max3(mx, __uint_as_float(ra[i]) + __uint_as_float(rb[i]), __uint_as_float(ra[i + 1]) + __uint_as_float(rb[i + 1]));
        }
        const int centroid = h * (64 / S) + g;
        out_c[(size_t)centroid * ld] = fmaxf(mx + bias, 0.f);   // relu(max(x) + b) == max(relu(x + b))
    }
}

// xyz-only first conv (3 -> 64) of one row, channels [CH0, CH0 + 32), weights in the constant bank
template <int CH0>
__device__ __forceinline__ void conv0_xyz(const SaLeanArgs &a, const float (&rel)[3], int r, uint8_t *A_hi, uint8_t *A_lo)
{
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = CH0 + q * 8 + i;
            v[i] = fmaf(rel[2], a.w0[128 + c], fmaf(rel[1], a.w0[64 + c], fmaf(rel[0], a.w0[c], a.w0[192 + c])));
        }
        const int kc = CH0 / 8 + q;
        split8_lean<true>(v, A_hi + (size_t)kc * 2048 + r * 16, A_lo + (size_t)kc * 2048 + r * 16);
    }
}

// C: feature channels of the dataset (0: xyz only, first conv on the CUDA cores).  N0/N1/N2: layer widths.
// S: nsample (32 / 64).  BALL: ball query in the kernel.
template <int C, int N0, int N1, int N2, int S, bool BALL>
__global__ void __launch_bounds__(NTHR, 2) sa_lean_kernel(const __grid_constant__ SaLeanArgs a)
{
    constexpr int K0 = C == 0 ? N0 : (C + 3 + 15) / 16 * 16;                  // width of the gathered operand
    constexpr int KMAX = K0 > N0 ? (K0 > N1 ? K0 : N1) : (N0 > N1 ? N0 : N1);
    constexpr int K8 = KMAX / 8;
    constexpr int NSTD = C == 0 ? 1 : 2;                                      // in-place (standard orientation) units
    constexpr int NTB = N2 / 128;                                             // transposed pooled blocks
    static_assert(N2 % 128 == 0 && NSTD + NTB <= MAXU, "layer widths");
    static_assert(C % 8 == 0 && N0 % 64 == 0 && N1 % 64 == 0, "layer widths");

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *A_hi = smem;
    uint8_t *A_lo = A_hi + (size_t)K8 * 2048;
    uint8_t *ones = A_lo + (size_t)K8 * 2048;                                 // [2 kc][128 rows][8] fp16: (1, 1, 0 ...)
    uint8_t *ring = ones + 4096;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(ring + (size_t)a.nst * STAGE_BYTES);
    uint64_t *bar_empty = bar_full + MAX_STAGES;
    uint64_t *bar_acc = bar_empty + MAX_STAGES;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar_acc + 1);
    int *s_idx = reinterpret_cast<int *>(s_tmem + 4);                         // [TM]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc(s_tmem, TMEM_COLS);
    if (tid == NWORK) {
        for (int s = 0; s < MAX_STAGES; ++s) { tc::mbar_init(bar_full + s, 1); tc::mbar_init(bar_empty + s, 1); }
        tc::mbar_init(bar_acc, 1);
    }
    if (tid < TM) {
        *reinterpret_cast<uint4 *>(ones + tid * 16) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(ones + 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    const uint32_t a_hi0 = tc::smem_u32(A_hi), a_lo0 = tc::smem_u32(A_lo), ring0 = tc::smem_u32(ring), ones0 = tc::smem_u32(ones);
    const int b = blockIdx.y;
    const int ntiles = a.tiles_per_cta;

    if (warp == NWORK / 32 + 1) {
        // =========================== producer warp: weight k-steps -> ring ===========================
        int slot = 0, round = 0;
        for (int t = 0; t < ntiles; ++t)
            for (int u = 0; u < a.nunits; ++u) {
                const LUnit &U = a.U[u];
                for (int kk = 0; kk < U.nk; ++kk) {
                    if (round > 0) tc::mbar_wait(bar_empty + slot, (uint32_t)((round - 1) & 1));
                    if (elect_one()) {
                        mbar_expect_tx(bar_full + slot, 4u * U.piece_bytes);
                        const uint32_t dst = ring0 + (uint32_t)slot * STAGE_BYTES;
                        const uint8_t *src = U.img + (size_t)kk * U.kstep_stride;
                        if (U.piece_stride == U.piece_bytes) {
                            bulk_g2s(dst, src, 4u * U.piece_bytes, bar_full + slot);
                        } else {
#pragma unroll
                            for (int p = 0; p < 4; ++p)
                                bulk_g2s(dst + (uint32_t)p * U.piece_bytes, src + (size_t)p * U.piece_stride, U.piece_bytes, bar_full + slot);
                        }
                    }
                    __syncwarp();
                    if (++slot == a.nst) { slot = 0; ++round; }
                }
            }
    } else if (warp == NWORK / 32) {
        // =========================== MMA warp ===========================
        int slot = 0, round = 0;
        const uint64_t ah0 = tc::smem_desc(a_hi0, 2048u, 128u), al0 = tc::smem_desc(a_lo0, 2048u, 128u);
        const uint64_t on0 = tc::smem_desc(ones0, 2048u, 128u);
        for (int t = 0; t < ntiles; ++t) {
            work_sync();                                          // operand of the tile gathered, TMEM drained
            LEAN_TRACE(64, 0);
            for (int u = 0; u < a.nunits; ++u) {
                const LUnit &U = a.U[u];
                tc::fence_after_sync();
                const uint32_t hh = tmem, cr = tmem + (U.transposed ? (uint32_t)TM : (U.piece_bytes >> 4));
                for (int kk = 0; kk < U.nk; ++kk) {
                    tc::mbar_wait(bar_full + slot, (uint32_t)(round & 1));
                    if (kk == 0) LEAN_TRACE(64, 1 + 4 * u);
                    if (kk == U.nk - 1) LEAN_TRACE(64, 2 + 4 * u);
                    if (elect_one()) {
                        const uint32_t st = ring0 + (uint32_t)slot * STAGE_BYTES;
                        const uint64_t wh = tc::smem_desc(st, 2u * U.piece_bytes, 128u);
                        const uint64_t wl = wh + (uint64_t)(U.piece_bytes >> 4);
                        const uint64_t ah = ah0 + (uint64_t)(kk * 256), al = al0 + (uint64_t)(kk * 256);   // 2 * 2048 / 16
                        if (U.bias && kk == U.nk - 1) {
                            tc::mma_f16(hh, on0, wh, U.idesc, 1u);
                        } else if (!U.transposed) {
                            tc::mma_f16(hh, ah, wh, U.idesc, kk > 0);
                            tc::mma_f16(cr, ah, wl, U.idesc, kk > 0);
                            tc::mma_f16(cr, al, wh, U.idesc, 1u);
                        } else {
                            tc::mma_f16(hh, wh, ah, U.idesc, kk > 0);
                            tc::mma_f16(cr, wh, al, U.idesc, kk > 0);
                            tc::mma_f16(cr, wl, ah, U.idesc, 1u);
                        }
                        tc::mma_commit(bar_empty + slot);
                    }
                    __syncwarp();
                    if (++slot == a.nst) { slot = 0; ++round; }
                }
                if (elect_one()) tc::mma_commit(bar_acc);
                __syncwarp();
                LEAN_TRACE(64, 3 + 4 * u);
                tc::fence_before_sync();
                work_sync();                                      // epilogue done: operand rewritten / TMEM drained
                LEAN_TRACE(64, 4 + 4 * u);
            }
        }
    } else {
        // =========================== workers ===========================
        const int r = tid & (TM - 1), h = tid >> 7, wq = warp & 3;
        const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);
        uint32_t uc = 0;                                          // units completed (parity of bar_acc)
        for (int t = 0; t < ntiles; ++t) {
            const int tile = blockIdx.x * ntiles + t;
            const int tb = (warp == 0 ? 0 : 32);                  // trace base (warps 0 and 4 record)
            if (warp == 0 || warp == 4) LEAN_TRACE(tb, 0);
            const long R = (long)tile * TM + r;                   // row inside the cloud = centroid * S + sample
            const int g = (int)(R / S);
            int id;
            if (BALL) {
                if (warp < TM / S) {
                    // query_ball_point for centroid `cen` (tf_grouping_g.cu:3-36): first S points in index order with
                    // max(sqrt(d2), 1e-20) < radius, row padded with the first hit (all zeros when the ball is empty)
                    const int cen = tile * (TM / S) + warp;
                    const float *c3 = a.new_xyz + ((size_t)b * a.m + cen) * 3;
                    const float x2 = __ldg(c3), y2 = __ldg(c3 + 1), z2 = __ldg(c3 + 2);
                    const float *p1 = a.xyz + (size_t)b * a.n * 3;
                    int *row = s_idx + warp * S;
                    int cnt = 0, first = 0;
                    for (int base = 0; base < a.n; base += 32) {
                        const int k = base + lane;
                        bool in = false;
                        if (k < a.n) {
                            const float dx = x2 - __ldg(p1 + k * 3 + 0);
                            const float dy = y2 - __ldg(p1 + k * 3 + 1);
                            const float dz = z2 - __ldg(p1 + k * 3 + 2);
                            float d = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
                            d = fmaxf(d, 1e-20f);
                            in = d < a.radius;
                        }
                        const unsigned mask = __ballot_sync(0xFFFFFFFFu, in);
                        if (mask) {
                            if (cnt == 0) first = base + __ffs(mask) - 1;
                            const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
                            if (in && pos < S) row[pos] = k;
                            cnt += __popc(mask);
                            if (cnt >= S) break;
                        }
                    }
                    cnt = min(cnt, S);
                    for (int l = cnt + lane; l < S; l += 32) row[l] = first;
                    if (lane == 0 && a.cnt_out) a.cnt_out[(size_t)b * a.m + cen] = cnt;
                }
                gather_sync();
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 1);
                id = s_idx[r];
                if (a.idx_out && h == 0) a.idx_out[(size_t)b * a.m * S + R] = id;
            } else {
                id = __ldg(a.idx_in + (size_t)b * a.m * S + R);
            }
            float rel[3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                rel[c] = __fsub_rn(__ldg(a.xyz + ((size_t)b * a.n + id) * 3 + c), __ldg(a.new_xyz + ((size_t)b * a.m + g) * 3 + c));
            if (C == 0) {
                if (h == 0) conv0_xyz<0>(a, rel, r, A_hi, A_lo);
                else conv0_xyz<32>(a, rel, r, A_hi, A_lo);
            } else {
                // channel order [features(C), xyz(3), zero pad] (pointnet_util.py:52-57 with the rows of W permuted)
                const float *prow = a.points + ((size_t)b * a.n + id) * C;
                constexpr int HALF = K0 / 16;                     // kc slices per thread
#pragma unroll
                for (int q0 = 0; q0 < HALF; q0 += 3) {
                    float4 p[6];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int kc = h * HALF + q0 + q;
                        if (q0 + q < HALF && kc * 8 < C) { p[2 * q] = ldg4(prow + kc * 8); p[2 * q + 1] = ldg4(prow + kc * 8 + 4); }
                    }
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int kc = h * HALF + q0 + q;
                        if (q0 + q >= HALF) continue;
                        float v[8];
                        if (kc * 8 < C) {
                            v[0] = p[2 * q].x; v[1] = p[2 * q].y; v[2] = p[2 * q].z; v[3] = p[2 * q].w;
                            v[4] = p[2 * q + 1].x; v[5] = p[2 * q + 1].y; v[6] = p[2 * q + 1].z; v[7] = p[2 * q + 1].w;
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = 0.f;
                            if (kc * 8 == C) { v[0] = rel[0]; v[1] = rel[1]; v[2] = rel[2]; }
                        }
                        split8_lean<false>(v, A_hi + (size_t)kc * 2048 + r * 16, A_lo + (size_t)kc * 2048 + r * 16);
                    }
                }
            }
            if (warp == 0 || warp == 4) LEAN_TRACE(tb, 2);
            tc::fence_proxy_async();
            work_sync();                                          // operand gathered
            if (warp == 0 || warp == 4) LEAN_TRACE(tb, 3);
            // ---- in-place layers ----
            if (NSTD == 2) {
                tc::mbar_wait(bar_acc, uc & 1u); ++uc;
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 4);
                tc::fence_after_sync();
                epi_inplace<N0>(trow, h, r, A_hi, A_lo);
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 5);
                tc::fence_proxy_async();
                tc::fence_before_sync();
                work_sync();
            }
            {
                tc::mbar_wait(bar_acc, uc & 1u); ++uc;
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 6);
                tc::fence_after_sync();
                epi_inplace<N1>(trow, h, r, A_hi, A_lo);
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 7);
                tc::fence_proxy_async();
                tc::fence_before_sync();
                work_sync();
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 8);
            }
            // ---- pooled layer, transposed: lane = output channel ----
#pragma unroll
            for (int blk = 0; blk < NTB; ++blk) {
                tc::mbar_wait(bar_acc, uc & 1u); ++uc;
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 9 + 3 * blk);
                tc::fence_after_sync();
                const int ch = blk * 128 + wq * 32 + lane;
                float *out_c = a.out + ((size_t)b * a.m + (size_t)tile * (TM / S)) * N2 + ch;
                epi_pool_T<S>(trow, h, __ldg(a.bias_last + ch), out_c, N2);
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 10 + 3 * blk);
                tc::fence_before_sync();
                work_sync();
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 11 + 3 * blk);
            }
        }
        if (warp == 0) tc::tmem_dealloc(tmem, TMEM_COLS);
    }
}

// ---- host -------------------------------------------------------------------------------------------------------
LUnit make_unit(const TcLayer &L, int n0, int nc, bool bias, bool transposed)
{
    LUnit U{};
    U.img = reinterpret_cast<const uint8_t *>(L.Wimg) + (size_t)n0 * 16;
    U.kstep_stride = 4u * (uint32_t)L.N * 16u;
    U.piece_stride = (uint32_t)L.N * 16u;
    U.piece_bytes = (uint32_t)nc * 16u;
    U.nk = L.K / 16 + (bias ? 1 : 0);
    U.bias = bias ? 1 : 0;
    U.transposed = transposed ? 1 : 0;
    U.idesc = transposed ? tc::instr_desc_f16(128, TM) : tc::instr_desc_f16(TM, nc);
    return U;
}

template <int C, int N0, int N1, int N2, int S, bool BALL>
int launch(const SaLeanArgs &a, dim3 grid, size_t smem, cudaStream_t st)
{
    auto k = sa_lean_kernel<C, N0, N1, N2, S, BALL>;
    ANCSH_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ANCSH_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    k<<<grid, NTHR, smem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

template <int C, int N0, int N1, int N2>
int dispatch(const SaLeanArgs &a, int S, bool ball, dim3 grid, size_t smem, cudaStream_t st)
{
    if (S == 32) return ball ? launch<C, N0, N1, N2, 32, true>(a, grid, smem, st) : launch<C, N0, N1, N2, 32, false>(a, grid, smem, st);
    if (S == 64) return ball ? launch<C, N0, N1, N2, 64, true>(a, grid, smem, st) : launch<C, N0, N1, N2, 64, false>(a, grid, smem, st);
    return ANCSH_ERR_UNSUPPORTED;
}

}  // namespace

// Returns ANCSH_ERR_UNSUPPORTED when the stage does not have one of the two specialised shapes (the caller then uses the
// generic chain kernel of net_tc2.cu).
int sa_lean_launch(const SaLeanArgs2 &s, int B, cudaStream_t st)
{
    const long rows = (long)s.m * s.S;
    if (rows % TM != 0 || (s.S != 32 && s.S != 64) || B > 65535) return ANCSH_ERR_UNSUPPORTED;
    for (int l = 0; l < 3; ++l)
        if (!s.L[l].Wimg || !s.L[l].relu || !s.L[l].has_bias_step) return ANCSH_ERR_UNSUPPORTED;
    const bool sa1 = s.C == 0 && s.L[0].N == 64 && s.L[1].K == 64 && s.L[1].N == 64 && s.L[2].K == 64 && s.L[2].N == 128 && s.w0_host;
    const bool sa2 = s.C == 128 && s.L[0].K == 144 && s.L[0].N == 128 && s.L[1].K == 128 && s.L[1].N == 128 && s.L[2].K == 128 &&
                     s.L[2].N == 256;
    if (!sa1 && !sa2) return ANCSH_ERR_UNSUPPORTED;
    const bool ball = s.idx_in == nullptr;
    SaLeanArgs a{};
    a.xyz = s.xyz; a.points = s.points; a.new_xyz = s.new_xyz; a.idx_in = s.idx_in; a.idx_out = s.idx_out; a.cnt_out = s.cnt_out;
    a.out = s.out; a.bias_last = s.L[2].bias; a.n = s.n; a.m = s.m; a.radius = s.radius;
    const int tiles = (int)(rows / TM);
    a.tiles_per_cta = tiles % 4 == 0 ? 4 : (tiles % 2 == 0 ? 2 : 1);
    int k8;
    if (sa1) {
        for (int i = 0; i < 256; ++i) a.w0[i] = s.w0_host[i];
        a.U[0] = make_unit(s.L[1], 0, 64, true, false);
        a.U[1] = make_unit(s.L[2], 0, 128, false, true);
        a.nunits = 2; k8 = 8; a.nst = 6;
    } else {
        a.U[0] = make_unit(s.L[0], 0, 128, true, false);
        a.U[1] = make_unit(s.L[1], 0, 128, true, false);
        a.U[2] = make_unit(s.L[2], 0, 128, false, true);
        a.U[3] = make_unit(s.L[2], 128, 128, false, true);
        a.nunits = 4; k8 = 18; a.nst = 4;
    }
    const size_t smem = (size_t)2 * k8 * 2048 + 4096 + (size_t)a.nst * STAGE_BYTES + (2 * MAX_STAGES + 1) * sizeof(uint64_t) + 16 +
                        TM * sizeof(int);
    const dim3 grid((unsigned)(tiles / a.tiles_per_cta), (unsigned)B);
    // profiling aid: ANCSH_LEAN_TRACE=<file prefix> dumps the phase time stamps of the 4th launch of each stage
    static const char *trace_path = getenv("ANCSH_LEAN_TRACE");
    static int trace_calls[2] = {0, 0};
    long long *trace_dev = nullptr;
    const size_t trace_n = (size_t)grid.x * grid.y * TRACE_SLOTS;
    if (trace_path && ++trace_calls[sa1 ? 0 : 1] == 4) {
        if (cudaMalloc(&trace_dev, trace_n * sizeof(long long)) != cudaSuccess) trace_dev = nullptr;
        else cudaMemsetAsync(trace_dev, 0, trace_n * sizeof(long long), st);
    }
    a.trace = trace_dev;
    const int rc = sa1 ? dispatch<0, 64, 64, 128>(a, s.S, ball, grid, smem, st) : dispatch<128, 128, 128, 256>(a, s.S, ball, grid, smem, st);
    if (trace_dev) {
        std::vector<long long> h(trace_n);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), trace_dev, trace_n * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(trace_dev);
        char name[512];
        snprintf(name, sizeof name, "%s_%s.txt", trace_path, sa1 ? "sa1" : "sa2");
        if (FILE *f = fopen(name, "w")) {
            fprintf(f, "# grid %u x %u, %d tiles per CTA; one line per CTA: %d clock64 stamps\n", grid.x, grid.y, a.tiles_per_cta, TRACE_SLOTS);
            for (size_t c = 0; c < (size_t)grid.x * grid.y; c += 37) {       // a sample of the CTAs
                for (int k = 0; k < TRACE_SLOTS; ++k) fprintf(f, "%lld ", h[c * TRACE_SLOTS + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    }
    return rc;
}
