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
// Numerics: as net_tc2.cu (hi*hi and the 2^11-scaled cross terms in separate f32 TMEM accumulators, combined in the epilogue).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "tc_common.cuh"
#include "net_tc.cuh"

namespace {

constexpr int TM = 128;                 // rows per tile
constexpr int NWORK = 256;              // worker threads
constexpr int NTHR = NWORK + 64;        // + MMA warp + producer warp (one lane of each runs the role)
constexpr int MAX_STAGES = 8;
constexpr int MAXU = 4;
constexpr uint32_t TMEM_COLS = 256;

// Compile-time plan of a stage: units = the GEMMs of a tile in issue order.
//   in-place units (standard orientation, activations = A operand): conv0 (only with input features) and conv1, each with
//   the bias k-step; then N2 / 128 transposed pooled blocks of the last conv (weights = A operand).
template <int C, int N0, int N1, int N2>
struct Plan {
    static constexpr int K0 = C == 0 ? N0 : (C + 3 + 15) / 16 * 16;          // width of the gathered operand
    static constexpr int KMAX = K0 > N0 ? (K0 > N1 ? K0 : N1) : (N0 > N1 ? N0 : N1);
    static constexpr int K8 = KMAX / 8;
    static constexpr int NSTD = C == 0 ? 1 : 2;
    static constexpr int NTB = N2 / 128;
    static constexpr int NU = NSTD + NTB;
    __host__ __device__ static constexpr bool transposed(int u) { return u >= NSTD; }
    __host__ __device__ static constexpr int kin(int u) { return C == 0 ? (u == 0 ? N0 : N1) : (u == 0 ? K0 : (u == 1 ? N0 : N1)); }
    __host__ __device__ static constexpr int nc(int u) { return transposed(u) ? 128 : (C == 0 ? N1 : (u == 0 ? N0 : N1)); }    // weight rows of the unit
    __host__ __device__ static constexpr int nk(int u) { return kin(u) / 16 + (transposed(u) ? 0 : 1); }                      // k-steps incl. bias step
    __host__ __device__ static constexpr int kstep_bytes(int u) { return nc(u) * 64; }                                        // 2 kc x (hi, lo) x nc x 16 B
    __host__ __device__ static constexpr int unit_bytes(int u) { return nk(u) * kstep_bytes(u); }
    __host__ __device__ static constexpr int all_bytes() { int t = 0; for (int u = 0; u < NU; ++u) t += unit_bytes(u); return t; }
    // weights of all units fit shared memory next to the operand: loaded once per CTA, one ring slot per unit
    static constexpr bool RESIDENT = all_bytes() + 2 * K8 * 2048 <= 100 * 1024;
    __host__ __device__ static constexpr int kps(int u) { return RESIDENT ? nk(u) : 1; }                                      // k-steps per ring stage
    __host__ __device__ static constexpr int max_unit_bytes() { int t = 0; for (int u = 0; u < NU; ++u) t = unit_bytes(u) > t ? unit_bytes(u) : t; return t; }
    static constexpr int SLOT_BYTES = RESIDENT ? max_unit_bytes() : 8192;
    static constexpr int NST = RESIDENT ? NU : 4;
    // resident plans: every unit has a slot of exactly its size (slot_off); streamed plans: NST stages of SLOT_BYTES
    __host__ __device__ static constexpr int slot_off(int u) { int t = 0; for (int v = 0; v < u; ++v) t += unit_bytes(v); return t; }
    static constexpr int RING_BYTES = RESIDENT ? all_bytes() : NST * SLOT_BYTES;
    // xyz-only stages stage the cloud's coordinates in shared memory for the ball query (n <= XYZ_MAX points)
    static constexpr int XYZ_MAX = C == 0 ? 1024 : 0;
    static_assert(N2 % 128 == 0 && NU <= MAXU && NST <= MAX_STAGES, "layer widths");
    static_assert(C % 8 == 0 && N0 % 64 == 0 && N1 % 64 == 0, "layer widths");
};

struct LUnit {
    const uint8_t *img;       // weight image [K/8][hi|lo][Nfull][8] fp16, offset to the unit's first row
    uint32_t kstep_stride;    // bytes between consecutive k-steps in the image (4 * Nfull * 16)
    uint32_t piece_stride;    // bytes between the four (kc, hi|lo) blocks of a k-step (Nfull * 16)
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
    float d2_below;           // smallest f32 t with sqrt_rn(t) >= radius: max(sqrt(d2), 1e-20) < radius  <=>  d2 < t (sqrt_rn is monotonic)
    int tiles_per_cta;
    LUnit U[MAXU];
    float descale[3];         // per layer (conv0, conv1, conv2): accumulators * descale = layer output (TcLayer::descale)
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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void gather_sync() { asm volatile("bar.sync 2, %0;" ::"n"(NWORK) : "memory"); }

// {upper half, lower half} = {fp16(hi_elem), fp16(lo_elem)}, round to nearest, optionally clamped at 0 (ReLU)
template <bool RELU>
__device__ __forceinline__ uint32_t pack_f16x2(float upper, float lower)
{
    uint32_t d;
    if (RELU) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(upper), "f"(lower));
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
        l[i] = pack_f16x2<RELU>((b - hb) * tc::LO_SCALE, (a - ha) * tc::LO_SCALE);        // lo piece stored * 2^11 (tc_common.cuh)
    }
    *reinterpret_cast<uint4 *>(dst_hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(dst_lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// 4 consecutive k of one row -> the 8-byte halves of the hi and lo pieces (no ReLU: gathered inputs)
__device__ __forceinline__ void split4_lean(const float (&v)[4], uint8_t *dst_hi, uint8_t *dst_lo)
{
    uint32_t h[2], l[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        const float ha = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), hb = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
        h[i] = pack_f16x2<false>(hb, ha);
        l[i] = pack_f16x2<false>((b - hb) * tc::LO_SCALE, (a - ha) * tc::LO_SCALE);
    }
    *reinterpret_cast<uint2 *>(dst_hi) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2 *>(dst_lo) = make_uint2(l[0], l[1]);
}

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// epilogue of an in-place layer: this thread's row, columns [h*N/2, (h+1)*N/2): (hi*hi + 2^-11 cross) * sc -> ReLU -> split -> operand
template <int N>
__device__ __forceinline__ void epi_inplace(uint32_t trow, int h, int r, float sc, uint8_t *A_hi, uint8_t *A_lo)
{
    const float sc_lo = sc * tc::LO_UNSCALE;
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
            for (int i = 0; i < 8; ++i) v[i] = fmaf(__uint_as_float(rb[q * 8 + i]), sc_lo, __uint_as_float(ra[q * 8 + i]) * sc);
            const int kc = (c0 >> 3) + q;
            split8_lean<true>(v, A_hi + (size_t)kc * 2048 + r * 16, A_lo + (size_t)kc * 2048 + r * 16);
        }
    }
}

// epilogue of a transposed pooled block: this thread's TMEM lane = one output channel, columns [h*64, h*64+64) = rows of
// 64 / S centroids.  out_c points at out[centroid 0 of the tile][channel of this lane]; ld = channels per centroid.
template <int S>
__device__ __forceinline__ void epi_pool_T(uint32_t trow, int h, float sc, float bias, float *out_c, int ld)
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
                mx = max3(mx, fmaf(__uint_as_float(rb[i]), tc::LO_UNSCALE, __uint_as_float(ra[i])),
                          fmaf(__uint_as_float(rb[i + 1]), tc::LO_UNSCALE, __uint_as_float(ra[i + 1])));
        }
        const int centroid = h * (64 / S) + g;
        out_c[(size_t)centroid * ld] = fmaxf(fmaf(mx, sc, bias), 0.f);   // relu(max(x) * sc + b) == max(relu(x * sc + b)), sc > 0
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

// ---- MMA role: all tcgen05.mma of unit U of the plan (one thread) -------------------------------------------------
struct MmaCtx {
    uint64_t ah0, al0, on0;          // descriptors of the activation images and the ones slab
    uint32_t ring0, tmem;
    uint64_t *bar_full, *bar_empty;
    int slot, round;                 // ring position (streaming plans)
};

template <class P, int U>
__device__ __forceinline__ void issue_unit(MmaCtx &c, bool first_tile)
{
    constexpr bool T = P::transposed(U);
    constexpr int NC = P::nc(U), NK = P::nk(U), KPS = P::kps(U);
    constexpr uint32_t idesc = T ? tc::instr_desc_f16(128, TM) : tc::instr_desc_f16(TM, NC);
    constexpr bool BIAS = !T;
    const uint32_t hh = c.tmem, cr = c.tmem + (uint32_t)(T ? TM : NC);     // hi*hi accumulator, cross-term accumulator
#pragma unroll
    for (int s0 = 0; s0 < NK; s0 += KPS) {
        uint32_t st;
        if (P::RESIDENT) {
            st = c.ring0 + (uint32_t)P::slot_off(U);
            if (first_tile) tc::mbar_wait(c.bar_full + U, 0u);
        } else {
            st = c.ring0 + (uint32_t)c.slot * P::SLOT_BYTES;
            tc::mbar_wait(c.bar_full + c.slot, (uint32_t)(c.round & 1));
        }
        const uint64_t wh0 = tc::smem_desc(st, 2u * NC * 16u, 128u), wl0 = wh0 + (uint64_t)NC;
#pragma unroll
        for (int j = 0; j < KPS; ++j) {
            const int kk = s0 + j;
            const uint64_t wh = wh0 + (uint64_t)(j * NC * 4), wl = wl0 + (uint64_t)(j * NC * 4);
            const uint64_t ah = c.ah0 + (uint64_t)(kk * 256), al = c.al0 + (uint64_t)(kk * 256);          // 2 * 2048 / 16
            if (BIAS && kk == NK - 1) {
                tc::mma_f16(hh, c.on0, wh, idesc, 1u);
            } else if (!T) {
                tc::mma_f16(hh, ah, wh, idesc, kk > 0);
                tc::mma_f16(cr, ah, wl, idesc, kk > 0);
                tc::mma_f16(cr, al, wh, idesc, 1u);
            } else {
                tc::mma_f16(hh, wh, ah, idesc, kk > 0);
                tc::mma_f16(cr, wh, al, idesc, kk > 0);
                tc::mma_f16(cr, wl, ah, idesc, 1u);
            }
        }
        if (!P::RESIDENT) {
            tc::mma_commit(c.bar_empty + c.slot);
            if (++c.slot == P::NST) { c.slot = 0; ++c.round; }
        }
    }
}

// C: feature channels of the dataset (0: xyz only, first conv on the CUDA cores).  N0/N1/N2: layer widths.
// S: nsample (32 / 64).  BALL: ball query in the kernel.
template <int C, int N0, int N1, int N2, int S, bool BALL>
__global__ void __launch_bounds__(NTHR, 2) sa_lean_kernel(const __grid_constant__ SaLeanArgs a)
{
    using P = Plan<C, N0, N1, N2>;
    constexpr int K0 = P::K0, K8 = P::K8, NSTD = P::NSTD, NTB = P::NTB, NU = P::NU;
    constexpr int CPT = TM / S;                                               // centroids per tile
    constexpr int WPC = (NWORK / 32) / CPT;                                   // warps per centroid in the ball query

    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *A_hi = smem;
    uint8_t *A_lo = A_hi + (size_t)K8 * 2048;
    uint8_t *ones = A_lo + (size_t)K8 * 2048;                                 // [2 kc][128 rows][8] fp16: (1, 1, 0 ...)
    uint8_t *ring = ones + 4096;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(ring + (size_t)P::RING_BYTES);
    uint64_t *bar_empty = bar_full + MAX_STAGES;
    uint64_t *bar_acc = bar_empty + MAX_STAGES;                               // MMA thread -> workers: accumulators complete
    uint64_t *bar_ready = bar_acc + 1;                                        // workers -> MMA thread: operand written / TMEM drained
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar_ready + 1);
    int *s_cnt = reinterpret_cast<int *>(s_tmem + 4);                         // [2][8] hits found by each ball-query warp
    int *s_idx = s_cnt + 16;                                                  // [2][8 warps][S] ball-query hit lists (double buffered)
    float *s_xyz = reinterpret_cast<float *>(s_idx + 16 * S);                 // [XYZ_MAX][3] the cloud's coordinates (xyz-only stages)
    int *s_rowid = reinterpret_cast<int *>(s_xyz + P::XYZ_MAX * 3);           // [2][TM] point index of every row of a tile (feature gather)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc(s_tmem, TMEM_COLS);
    if (tid == NWORK) {
        for (int s = 0; s < MAX_STAGES; ++s) { tc::mbar_init(bar_full + s, 1); tc::mbar_init(bar_empty + s, 1); }
        tc::mbar_init(bar_acc, 1);
        tc::mbar_init(bar_ready, NWORK / 32);
    }
    if (tid < TM) {
        *reinterpret_cast<uint4 *>(ones + tid * 16) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(ones + 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    const int b = blockIdx.y;
    const int ntiles = a.tiles_per_cta;

    if (warp == NWORK / 32 + 1) {
        // =========================== producer (one thread): weight stages -> ring ===========================
        if (lane == 0) {
            const uint32_t ring0 = tc::smem_u32(ring);
            int slot = 0, round = 0;
            for (int t = 0; t < (P::RESIDENT ? 1 : ntiles); ++t) {
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const LUnit &U = a.U[u];
                    const uint32_t piece = (uint32_t)P::nc(u) * 16u;
                    for (int s0 = 0; s0 < P::nk(u); s0 += P::kps(u)) {
                        if (!P::RESIDENT && round > 0) tc::mbar_wait(bar_empty + slot, (uint32_t)((round - 1) & 1));
                        mbar_expect_tx(bar_full + slot, (uint32_t)P::kps(u) * 4u * piece);
                        for (int j = 0; j < P::kps(u); ++j) {
                            const uint32_t dst = ring0 + (uint32_t)(P::RESIDENT ? P::slot_off(u) : slot * P::SLOT_BYTES) + (uint32_t)j * 4u * piece;
                            const uint8_t *src = U.img + (size_t)(s0 + j) * U.kstep_stride;
                            if (U.piece_stride == piece) {
                                bulk_g2s(dst, src, 4u * piece, bar_full + slot);
                            } else {
#pragma unroll
                                for (int p = 0; p < 4; ++p)
                                    bulk_g2s(dst + (uint32_t)p * piece, src + (size_t)p * U.piece_stride, piece, bar_full + slot);
                            }
                        }
                        if (++slot == P::NST) { slot = 0; ++round; }
                    }
                }
            }
        }
    } else if (warp == NWORK / 32) {
        // =========================== MMA role (one thread) ===========================
        if (lane == 0) {
            MmaCtx c;
            c.ah0 = tc::smem_desc(tc::smem_u32(A_hi), 2048u, 128u);
            c.al0 = tc::smem_desc(tc::smem_u32(A_lo), 2048u, 128u);
            c.on0 = tc::smem_desc(tc::smem_u32(ones), 2048u, 128u);
            c.ring0 = tc::smem_u32(ring); c.tmem = tmem; c.bar_full = bar_full; c.bar_empty = bar_empty;
            c.slot = 0; c.round = 0;
            uint32_t rp = 0;                                      // phase of bar_ready
            for (int t = 0; t < ntiles; ++t) {
                tc::mbar_wait(bar_ready, rp & 1u); ++rp;          // operand of the tile gathered, TMEM drained
                tc::fence_after_sync();
                LEAN_TRACE(64, 0);
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    if (u == 0) issue_unit<P, 0>(c, t == 0);
                    if (u == 1) issue_unit<P, 1 < NU ? 1 : 0>(c, t == 0);
                    if (u == 2) issue_unit<P, 2 < NU ? 2 : 0>(c, t == 0);
                    if (u == 3) issue_unit<P, 3 < NU ? 3 : 0>(c, t == 0);
                    tc::mma_commit(bar_acc);
                    LEAN_TRACE(64, 3 + 4 * u);
                    tc::mbar_wait(bar_ready, rp & 1u); ++rp;      // epilogue done: operand rewritten / TMEM drained
                    tc::fence_after_sync();
                    LEAN_TRACE(64, 4 + 4 * u);
                }
            }
        }
    } else {
        // =========================== workers ===========================
        const int r = tid & (TM - 1), h = tid >> 7, wq = warp & 3;
        const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);
        uint32_t uc = 0;                                          // units completed (parity of bar_acc)
        // everything this thread wrote (operand / TMEM reads) is done: tell the MMA thread
        auto ready = [&]() {
            tc::fence_proxy_async();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready);
        };
        // clouds of up to XYZ_MAX points are scanned from shared memory, larger ones through L1 (generic loads either way)
        const bool USE_SXYZ = BALL && P::XYZ_MAX > 0 && a.n <= P::XYZ_MAX;
        // query_ball_point of one tile (tf_grouping_g.cu:3-36): first S points in index order with max(sqrt(d2), 1e-20) < radius,
        // row padded with the first hit (all zeros when the ball is empty).  WPC warps per centroid, each scans a contiguous
        // range of the points (128 per trip, 4 per lane) and keeps an ordered hit list; fetch_id() concatenates the lists in
        // range order.  Runs in the shadow of the previous tile's first MMA unit (the workers would only wait there).
        auto ball_tile = [&](int tile, int buf) {
            const int cl = warp / WPC, part = warp % WPC;              // centroid of the tile, range of this warp
            const int cen = tile * CPT + cl;
            const float *c3 = a.new_xyz + ((size_t)b * a.m + cen) * 3;
            const float x2 = __ldg(c3), y2 = __ldg(c3 + 1), z2 = __ldg(c3 + 2);
            const float *p1 = USE_SXYZ ? s_xyz : a.xyz + (size_t)b * a.n * 3;
            const int per = ((a.n + WPC - 1) / WPC + 31) & ~31;
            const int k_beg = part * per, k_end = min(a.n, k_beg + per);
            int *row = s_idx + (buf * 8 + warp) * S;
            int cnt = 0;
            for (int base = k_beg; base < k_end && cnt < S; base += 128) {
                bool in[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = base + u * 32 + lane;
                    in[u] = false;
                    if (k < k_end) {
                        const float dx = x2 - p1[k * 3 + 0];
                        const float dy = y2 - p1[k * 3 + 1];
                        const float dz = z2 - p1[k * 3 + 2];
                        // tf_grouping_g.cu:21: d = max(sqrtf(dx*dx + dy*dy + dz*dz), 1e-20f) < radius, evaluated on d2 (see d2_below)
                        in[u] = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))) < a.d2_below;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned mask = __ballot_sync(0xFFFFFFFFu, in[u]);
                    const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
                    if (in[u] && pos < S) row[pos] = base + u * 32 + lane;
                    cnt += __popc(mask);
                }
            }
            if (lane == 0) s_cnt[buf * 8 + warp] = min(cnt, S);
        };
        // index of this thread's row (after a gather_sync that follows ball_tile of the same buffer)
        auto fetch_id = [&](int tile, int buf) -> int {
            const long R = (long)tile * TM + r;                        // row inside the cloud = centroid * S + sample
            if (!BALL) return __ldg(a.idx_in + (size_t)b * a.m * S + R);
            const int cl = r / S;
            int j = r - cl * S, total = 0, first = 0, found = -1;
            bool have_first = false;
#pragma unroll
            for (int p = 0; p < WPC; ++p) {
                const int w = buf * 8 + cl * WPC + p, cp = s_cnt[w];
                if (!have_first && cp > 0) { first = s_idx[w * S]; have_first = true; }
                if (found < 0 && j < cp) found = s_idx[w * S + j];
                j -= cp; total += cp;
            }
            const int id = found >= 0 ? found : first;
            if (a.idx_out && h == 0) a.idx_out[(size_t)b * a.m * S + R] = id;
            if (a.cnt_out && h == 0 && r % S == 0) a.cnt_out[(size_t)b * a.m + R / S] = min(total, S);
            return id;
        };
        auto load_rel = [&](int tile, int id, float (&rel)[3]) {
            const int g = (int)(((long)tile * TM + r) / S);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                rel[c] = __fsub_rn(__ldg(a.xyz + ((size_t)b * a.n + id) * 3 + c), __ldg(a.new_xyz + ((size_t)b * a.m + g) * 3 + c));
        };

        const int tile0 = blockIdx.x * ntiles;
        int id;
        float rel[3];
        if (USE_SXYZ) {
            // the CTA's tiles all belong to cloud b: stage its coordinates once (AoS, stride 3: conflict-free for consecutive k)
            const float *src = a.xyz + (size_t)b * a.n * 3;
            for (int i = tid; i < a.n * 3; i += NWORK) s_xyz[i] = __ldg(src + i);
            gather_sync();
        }
        if (BALL) { ball_tile(tile0, 0); gather_sync(); }
        id = fetch_id(tile0, 0);
        load_rel(tile0, id, rel);
        if (C != 0) {
            if (h == 0) s_rowid[r] = id;
            gather_sync();
        }
        for (int t = 0; t < ntiles; ++t) {
            const int tile = tile0 + t;
            const bool has_next = t + 1 < ntiles;
            const int tb = (warp == 0 ? 0 : 32);                  // trace base (warps 0 and 4 record)
            if (warp == 0 || warp == 4) LEAN_TRACE(tb, 0);
            if constexpr (C == 0) {
                if (h == 0) conv0_xyz<0>(a, rel, r, A_hi, A_lo);
                else conv0_xyz<32>(a, rel, r, A_hi, A_lo);
            } else {
                // channel order [features(C), xyz(3), zero pad] (pointnet_util.py:52-57 with the rows of W permuted).
                // Feature rows: 8 lanes per row (one 128-byte line per row and load instruction), 4 rows per warp
                // instruction, 16 rows per warp -- one thread per row touches 32 lines per instruction and made the gather
                // 9.5k of the tile's 31k cycles.  Lane (g, j) holds channels [32 i + 4 j, +4) of row 16 warp + 4 grp + g.
                static_assert(C % 32 == 0 && K0 == C + 16, "feature gather layout");
                const int *rid = s_rowid + (t & 1) * TM;
                const int g4 = lane >> 3, j = lane & 7;
#pragma unroll
                for (int grp0 = 0; grp0 < 4; grp0 += 2) {
                    float4 p[2][C / 32];
#pragma unroll
                    for (int gi = 0; gi < 2; ++gi) {
                        const int rr = warp * 16 + (grp0 + gi) * 4 + g4;
                        const float *prow = a.points + ((size_t)b * a.n + rid[rr]) * C;
#pragma unroll
                        for (int i = 0; i < C / 32; ++i) p[gi][i] = ldg4(prow + 32 * i + 4 * j);
                    }
#pragma unroll
                    for (int gi = 0; gi < 2; ++gi) {
                        const int rr = warp * 16 + (grp0 + gi) * 4 + g4;
#pragma unroll
                        for (int i = 0; i < C / 32; ++i) {
                            const float v[4] = {p[gi][i].x, p[gi][i].y, p[gi][i].z, p[gi][i].w};
                            const size_t off = (size_t)(4 * i + (j >> 1)) * 2048 + rr * 16 + (j & 1) * 8;
                            split4_lean(v, A_hi + off, A_lo + off);
                        }
                    }
                }
                {
                    // slab C/8: (rel xyz, 0 ...) by the row's first thread, slab C/8 + 1: zeros by its second
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.f;
                    if (h == 0) { v[0] = rel[0]; v[1] = rel[1]; v[2] = rel[2]; }
                    const int kc = C / 8 + h;
                    split8_lean<false>(v, A_hi + (size_t)kc * 2048 + r * 16, A_lo + (size_t)kc * 2048 + r * 16);
                }
            }
            if (warp == 0 || warp == 4) LEAN_TRACE(tb, 2);
            ready();                                              // operand gathered
            // ball query of the NEXT tile while the tensor pipe runs this tile's first unit
            if (BALL && has_next) ball_tile(tile + 1, (t + 1) & 1);
            if (warp == 0 || warp == 4) LEAN_TRACE(tb, 3);
            // index + relative coordinate of this thread's row of the NEXT tile (the loads fly under this tile's units); the
            // row-index table of the feature gather is published by the ready() arrivals that follow
            int id_n = 0;
            float rel_n[3] = {0.f, 0.f, 0.f};
            if (has_next) {
                if (BALL) gather_sync();                          // every warp's hit list of the next tile is complete
                id_n = fetch_id(tile + 1, (t + 1) & 1);
                load_rel(tile + 1, id_n, rel_n);
                if (C != 0 && h == 0) s_rowid[((t + 1) & 1) * TM + r] = id_n;
            }
            // ---- in-place layers ----
            if (NSTD == 2) {
                tc::mbar_wait(bar_acc, uc & 1u); ++uc;
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 4);
                tc::fence_after_sync();
                epi_inplace<N0>(trow, h, r, a.descale[0], A_hi, A_lo);
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 5);
                ready();
            }
            {
                tc::mbar_wait(bar_acc, uc & 1u); ++uc;
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 6);
                tc::fence_after_sync();
                epi_inplace<N1>(trow, h, r, a.descale[1], A_hi, A_lo);
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 7);
                ready();
            }
            // ---- pooled layer, transposed: lane = output channel ----
#pragma unroll
            for (int blk = 0; blk < NTB; ++blk) {
                tc::mbar_wait(bar_acc, uc & 1u); ++uc;
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 9 + 3 * blk);
                tc::fence_after_sync();
                const int ch = blk * 128 + wq * 32 + lane;
                float *out_c = a.out + ((size_t)b * a.m + (size_t)tile * CPT) * N2 + ch;
                epi_pool_T<S>(trow, h, a.descale[2], __ldg(a.bias_last + ch), out_c, N2);
                if (warp == 0 || warp == 4) LEAN_TRACE(tb, 10 + 3 * blk);
                ready();
            }
            id = id_n;
            rel[0] = rel_n[0]; rel[1] = rel_n[1]; rel[2] = rel_n[2];
        }
        // the last ready() was consumed by the MMA thread before it left its loop; TMEM is idle
        asm volatile("bar.sync 3, %0;" ::"n"(NWORK) : "memory");
        if (warp == 0) tc::tmem_dealloc(tmem, TMEM_COLS);
    }
}

// ---- host -------------------------------------------------------------------------------------------------------
LUnit make_unit(const TcLayer &L, int n0)
{
    LUnit U{};
    U.img = reinterpret_cast<const uint8_t *>(L.Wimg) + (size_t)n0 * 16;
    U.kstep_stride = 4u * (uint32_t)L.N * 16u;
    U.piece_stride = (uint32_t)L.N * 16u;
    return U;
}

template <int C, int N0, int N1, int N2, int S, bool BALL>
int launch(const SaLeanArgs &a, dim3 grid, cudaStream_t st)
{
    using P = Plan<C, N0, N1, N2>;
    const size_t smem = (size_t)2 * P::K8 * 2048 + 4096 + (size_t)P::RING_BYTES + (2 * MAX_STAGES + 2) * sizeof(uint64_t) + 16 +
                        16 * sizeof(int) + 16 * S * sizeof(int) + (size_t)P::XYZ_MAX * 3 * sizeof(float) + (C != 0 ? 2 * TM * sizeof(int) : 0);
    auto k = sa_lean_kernel<C, N0, N1, N2, S, BALL>;
    ANCSH_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ANCSH_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    k<<<grid, NTHR, smem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

template <int C, int N0, int N1, int N2>
int dispatch(const SaLeanArgs &a, int S, bool ball, dim3 grid, cudaStream_t st)
{
    if (S == 32) return ball ? launch<C, N0, N1, N2, 32, true>(a, grid, st) : launch<C, N0, N1, N2, 32, false>(a, grid, st);
    if (S == 64) return ball ? launch<C, N0, N1, N2, 64, true>(a, grid, st) : launch<C, N0, N1, N2, 64, false>(a, grid, st);
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
    for (int l = 0; l < 3; ++l) a.descale[l] = s.L[l].descale;
    {
        // threshold on the squared distance that reproduces the reference's test on the rounded square root bit for bit
        if (!(s.radius > 1e-20f) || !std::isfinite(s.radius)) return ANCSH_ERR_UNSUPPORTED;
        float t = s.radius * s.radius;
        while (t > 0.f && sqrtf(t) >= s.radius) t = nextafterf(t, 0.f);
        while (sqrtf(t) < s.radius) t = nextafterf(t, INFINITY);
        a.d2_below = t;
    }
    const int tiles = (int)(rows / TM);
    // resident weights (layer1): long-lived CTAs amortise the one-time weight load; streamed weights: 4 tiles per CTA
    int tpc = sa1 ? 16 : 4;
    while (tpc > 1 && tiles % tpc != 0) tpc >>= 1;
    a.tiles_per_cta = tpc;
    if (sa1) {
        for (int i = 0; i < 256; ++i) a.w0[i] = s.w0_host[i];
        a.U[0] = make_unit(s.L[1], 0);
        a.U[1] = make_unit(s.L[2], 0);
    } else {
        a.U[0] = make_unit(s.L[0], 0);
        a.U[1] = make_unit(s.L[1], 0);
        a.U[2] = make_unit(s.L[2], 0);
        a.U[3] = make_unit(s.L[2], 128);
    }
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
    const int rc = sa1 ? dispatch<0, 64, 64, 128>(a, s.S, ball, grid, st) : dispatch<128, 128, 128, 256>(a, s.S, ball, grid, st);
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
