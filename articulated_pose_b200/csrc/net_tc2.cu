// net_tc2.cu -- warp-specialised tcgen05 kernel for the operand-resident layer chains of the network:
//   set abstraction (layer1 / layer2): ball-query indices -> gather(features, xyz - centroid) -> 3 x (1x1 conv + folded
//                    BN + ReLU) -> max over nsample                                  (pointnet_util.py:94-161)
//   point-wise chains (fa_layer1/2/3, fc1, all heads): rows from global memory, a chain of layers (:206-236,
//                    architectures.py:89-93, architecture.py:98-159, 195-208)
//
// One CTA = 128 rows = 10 warps:
//   warps 0..7  workers: gather the rows into the fp16 hi/lo operand images (tc_common.cuh), run every epilogue
//               (TMEM -> bias + ReLU -> next operand in place / rows to HBM / max-pool).  Warps w and w+4 own the same
//               32 TMEM lanes (rows) and split the columns in 32-wide groups.
//   warp 8      MMA: one elected lane issues the tcgen05.mma chain of each layer as soon as the weight stage it needs has
//               landed and releases the stage with tcgen05.commit.
//   warp 9      producer: one elected lane streams the pre-tiled weight images from L2 into a ring of 8 KB stages with
//               1-D bulk copies (cp.async.bulk + mbarrier complete_tx); it runs ahead across layer boundaries, so the
//               next layer's first stages arrive while the workers are still in the epilogue.
// The first version of these kernels (net_tc.cu) staged every 32-wide weight slice with cp.async from all threads and
// waited for it before the MMAs (ncu: tensor pipe 17% active, the load latency of each slice fully exposed).
//
// "Unit" = the MMAs whose accumulators are live together followed by one epilogue: a whole in-place layer (all its
// 128-column chunks must have read the operand before it is overwritten) or one chunk of an output layer.
// Numerics as in net_tc.cu: hi*hi products and the cross terms hi*lo + lo*hi go to separate f32 accumulators (the tensor
// core accumulates with truncation), long k ranges are split over several hi*hi accumulators, the epilogue adds them in
// round-to-nearest f32.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "tc_common.cuh"
#include "net_tc.cuh"
#include "ops.cuh"

namespace {

constexpr int TM = 128;                 // rows per CTA
constexpr int NWORK = 256;              // worker threads
constexpr int NTHR = NWORK + 64;        // + MMA warp + producer warp
constexpr int STAGE_BYTES = 8192;       // one ring stage of weights
constexpr int MAX_STAGES = 8;        // one-CTA-per-SM plans (operand > 80 KB): the ring is all that hides the L2 latency
constexpr int DEF_STAGES = 4;        // two / three CTAs per SM
constexpr int MAX_UNITS = 16;

struct Unit {
    const __half *Wimg;    // [K/8][hi|lo][Nfull][8] fp16
    const float *bias;     // [Nfull], or [clouds][bias_stride] when bias_stride > 0
    float *out;            // rows [R][ldo] (pool == 0) or pooled [groups][Nfull] (pool == 1), or NULL
    long bias_stride;
    int K, Nfull;
    int n0;                // first column of chunk 0
    int nc;                // chunk width == accumulator stride in TMEM columns (32 / 64 / 128)
    int nchunks;
    int G;                 // hi*hi accumulators per chunk
    int kb[8];             // kb[g] = first k-step of hi*hi accumulator g + 1 (k-step kk goes to accumulator kk * G / nk16)
    int ksl;               // k16 steps per ring stage
    int nsl;               // ring stages (slices) per chunk
    int relu, inplace, pool, ldo;
    int act;               // TC_ACT_*: head activations in the epilogue
    int image;             // out is the fp16 hi/lo operand image of a following streaming GEMM (TC_DST_IMAGE)
    float descale;         // accumulators * descale = the layer's output (TcLayer::descale)
};

struct Chain2Args {
    // rows from global memory: [X1 | X2]
    const float *X1, *X2;
    int C1, C2;
    long rows_per_cloud;
    // set-abstraction gather
    const float *xyz, *points, *new_xyz;
    const int *idx;
    int n, m, S, C;
    const float *W0, *b0;  // xyz-only set abstraction: f32 weights [>=3][N0] / bias of the first conv, evaluated in the gather
    int N0, relu0;         // (N0 == 0: disabled)
    // fused feature propagation (FP kernels): X1 part of a row = blend of three rows of fp_points2
    const float *fp_points2;
    const int *fp_idx;     // [rows][3] three_nn indices into the cloud's fp_m2 known points
    const float *fp_w;     // [rows][3] interpolation weights
    int fp_m2;
    ancsh_pred_t pred;     // outputs of the head units
    int n_parts, mixed;
    int K0;                // operand width of the first tensor-core layer (multiple of 16)
    int kmax8;             // operand image width / 8
    int nst;               // ring stages
    int nunits;
    uint32_t tmem_cols;
    long long *trace;      // profiling aid (ANCSH_CHAIN_TRACE): per-CTA clock64 stamps, or NULL
    Unit U[MAX_UNITS];
};

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

// slot layout of a CTA's trace record: [0] start, [1] operand gathered, [2 + 2u] accumulators of unit u ready (worker warp 0),
// [3 + 2u] its epilogue done; [40 + 2u] MMA warp starts unit u, [41 + 2u] has committed it
constexpr int CTRACE_SLOTS = 80;
#define CHAIN_TRACE(k)                                                                     \
    do {                                                                                   \
        if (a.trace && lane == 0) a.trace[(size_t)blockIdx.x * CTRACE_SLOTS + (k)] = clock64(); \
    } while (0)

// one lane of a converged warp (ptxas then knows that the tcgen05 / bulk-copy operands below are warp-uniform)
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// barrier of the workers and the MMA warp (the producer warp runs free of it)
__device__ __forceinline__ void work_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NWORK + 32) : "memory"); }

// 16-column variants of the epilogue primitives (three-CTAs-per-SM build: 64 registers per thread)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void load_acc2(uint32_t trow, int G, int c0, float (&v)[16], int stride)
{
    uint32_t ra[16], rb[16];
    tmem_ld16_issue(trow + c0, ra);
    tmem_ld16_issue(trow + (uint32_t)stride + c0, rb);
    tc::tmem_wait_ld();
#pragma unroll
    // accumulators 0 .. G-1: hi*hi parts; accumulator G: cross terms, scaled by 2^11 (tc_common.cuh)
    const float f1 = G == 1 ? tc::LO_UNSCALE : 1.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(rb[i]), f1, __uint_as_float(ra[i]));
    for (int g = 2; g <= G; ++g) {
        tmem_ld16_issue(trow + (uint32_t)(g * stride) + c0, ra);
        tc::tmem_wait_ld();
        const float f = g == G ? tc::LO_UNSCALE : 1.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(ra[i]), f, v[i]);
    }
}

// 32 accumulator columns of this thread's row: sum of the hi*hi accumulators and the cross-term accumulator
__device__ __forceinline__ void load_acc2(uint32_t trow, int G, int c0, float (&v)[32], int stride)
{
    if (G == 1) {
        uint32_t ra[32], rb[32];
        tc::tmem_ld32_issue(trow + c0, ra);
        tc::tmem_ld32_issue(trow + (uint32_t)stride + c0, rb);
        tc::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(rb[i]), tc::LO_UNSCALE, __uint_as_float(ra[i]));
        return;
    }
    tc::tmem_ld32(trow + c0, v);
    for (int g = 1; g <= G; ++g) {
        float u[32];
        tc::tmem_ld32(trow + (uint32_t)(g * stride) + c0, u);
        const float f = g == G ? tc::LO_UNSCALE : 1.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(u[i], f, v[i]);
    }
}

// 8 consecutive k of one row -> one 16-byte piece each for the hi and lo operand images, packed conversions
__device__ __forceinline__ void split8(const float (&v)[8], uint4 *dst_hi, uint4 *dst_lo)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hi = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 back = __half22float2(hi);
        const __half2 lo = __floats2half2_rn((v[2 * i] - back.x) * tc::LO_SCALE, (v[2 * i + 1] - back.y) * tc::LO_SCALE);
        h[i] = *reinterpret_cast<const uint32_t *>(&hi);
        l[i] = *reinterpret_cast<const uint32_t *>(&lo);
    }
    *dst_hi = make_uint4(h[0], h[1], h[2], h[3]);
    *dst_lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int CW>
__device__ __forceinline__ void store_operand2(uint8_t *A_hi, uint8_t *A_lo, int row, int col0, const float (&v)[CW])
{
#pragma unroll
    for (int q = 0; q < CW / 8; ++q) {
        float w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = v[q * 8 + i];
        const int kc = (col0 >> 3) + q;
        split8(w, reinterpret_cast<uint4 *>(A_hi + (size_t)kc * 2048 + row * 16), reinterpret_cast<uint4 *>(A_lo + (size_t)kc * 2048 + row * 16));
    }
}

// max over the 32 rows of a warp (values >= 0 after ReLU: unsigned order of the bit patterns == float order)
template <int CW>
__device__ __forceinline__ void pool_store2(float *orow, int S, int lane, const float (&v)[CW])
{
    uint32_t keep = 0;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
        const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(v[i]));
        if (lane == i) keep = mx;
    }
    if (lane < CW) {
        if (S == 32) orow[lane] = __uint_as_float(keep);
        else atomicMax(reinterpret_cast<int *>(orow + lane), (int)keep);
    }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Head activations of one row (lib/architecture.py:122-139, 150-157) on the first 32 columns of a packed head layer:
//   nocs_net  [W(K) | nocs(3K) | scale(K) | trans(3K) | confi(1)]  (MIXED)  /  [W(K) | nocs(3K) | confi(1)]
//   joint_net [joint_axis(3) | unitvec(3) | heatmap(1) | index(3)]
// The row's TWO worker threads (h = 0 / 1) share the transcendental work: parts are dealt out by parity (each part's
// nocs / scale / trans / gocs are independent of the other parts), the segmentation softmax and the confidence go to the
// thread with fewer parts.  Every output pointer may be NULL (not requested).  K and MIXED are compile-time so that every
// index into x is static (a dynamically indexed x would push the accumulator rows of ALL epilogues of the kernel into
// local memory: measured 0.61 -> 0.87 ms for fa_layer3 + heads).  confi32: the confidence column (only when it is column 32,
// K = 4 mixed) read from the second column group by thread h = 1.
template <int K, bool MIXED>
__device__ __forceinline__ void nocs_head_act(int h, const float (&x)[32], float confi32, long r, const ancsh_pred_t &o)
{
    constexpr int CONFI = MIXED ? 8 * K : 4 * K;
    static_assert((MIXED ? 8 * K : 4 * K) <= 32, "W | nocs | scale | trans of a row must share one 32-column group");
    constexpr int H_EXTRA = (K & 1) ? 1 : 0;                 // thread that also does the softmax (h = 1 has fewer parts when K is odd)
#pragma unroll
    for (int p = 0; p < K; ++p) {
        if ((p & 1) != h) continue;
        float nocs[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) nocs[c] = sigmoidf_(x[K + 3 * p + c]);
        if (o.nocs_per_point) {
#pragma unroll
            for (int c = 0; c < 3; ++c) o.nocs_per_point[(size_t)r * 3 * K + 3 * p + c] = nocs[c];
        }
        if (MIXED) {
            const float sc = sigmoidf_(x[4 * K + p]);
            float tr[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) tr[c] = tanhf(x[5 * K + 3 * p + c]);
            if (o.global_scale) o.global_scale[(size_t)r * K + p] = sc;
            if (o.global_translation) {
#pragma unroll
                for (int c = 0; c < 3; ++c) o.global_translation[(size_t)r * 3 * K + 3 * p + c] = tr[c];
            }
            if (o.gocs_per_point) {
#pragma unroll
                for (int c = 0; c < 3; ++c) o.gocs_per_point[(size_t)r * 3 * K + 3 * p + c] = __fadd_rn(__fmul_rn(nocs[c], sc), tr[c]);
            }
        }
    }
    if (h == H_EXTRA && o.W) {
        float e[K], mx = x[0], sum = 0.f;
#pragma unroll
        for (int k = 1; k < K; ++k) mx = fmaxf(mx, x[k]);
#pragma unroll
        for (int k = 0; k < K; ++k) { e[k] = expf(x[k] - mx); sum += e[k]; }
#pragma unroll
        for (int k = 0; k < K; ++k) o.W[(size_t)r * K + k] = e[k] / sum;
    }
    if (o.confi_per_point) {
        if (CONFI < 32) { if (h == H_EXTRA) o.confi_per_point[r] = sigmoidf_(x[CONFI & 31]); }
        else if (h == 1) o.confi_per_point[r] = sigmoidf_(confi32);
    }
}

__device__ __forceinline__ void joint_head_act(int h, const float (&x)[32], long r, const ancsh_pred_t &o)
{
    if (h == 0) {
        if (o.joint_axis_per_point) {
#pragma unroll
            for (int k = 0; k < 3; ++k) o.joint_axis_per_point[(size_t)r * 3 + k] = tanhf(x[k]);
        }
        if (o.heatmap_per_point) o.heatmap_per_point[r] = sigmoidf_(x[6]);
    } else {
        if (o.unitvec_per_point) {
#pragma unroll
            for (int k = 0; k < 3; ++k) o.unitvec_per_point[(size_t)r * 3 + k] = tanhf(x[3 + k]);
        }
        if (o.index_per_point) {
            const float m2 = fmaxf(x[7], fmaxf(x[8], x[9]));
            const float e0 = expf(x[7] - m2), e1 = expf(x[8] - m2), e2 = expf(x[9] - m2), es = e0 + e1 + e2;
            o.index_per_point[(size_t)r * 3 + 0] = e0 / es;
            o.index_per_point[(size_t)r * 3 + 1] = e1 / es;
            o.index_per_point[(size_t)r * 3 + 2] = e2 / es;
        }
    }
}

// Out of line on purpose: the transcendental-heavy head code must not share the register budget (96) of the epilogue loop.
__device__ __noinline__ void head_activations(int act, int h, const float (&x)[32], float confi32, long r, int K, int mixed,
                                              const ancsh_pred_t &o)
{
    if (act == TC_ACT_JOINT_HEADS) { joint_head_act(h, x, r, o); return; }
    if (mixed) {
        if (K == 2) nocs_head_act<2, true>(h, x, confi32, r, o);
        else if (K == 3) nocs_head_act<3, true>(h, x, confi32, r, o);
        else nocs_head_act<4, true>(h, x, confi32, r, o);
    } else {
        if (K == 2) nocs_head_act<2, false>(h, x, confi32, r, o);
        else if (K == 3) nocs_head_act<3, false>(h, x, confi32, r, o);
        else if (K == 4) nocs_head_act<4, false>(h, x, confi32, r, o);
        else if (K == 5) nocs_head_act<5, false>(h, x, confi32, r, o);
        else if (K == 6) nocs_head_act<6, false>(h, x, confi32, r, o);
        else nocs_head_act<7, false>(h, x, confi32, r, o);
    }
}

// SA: set-abstraction gather (ball-query indices).  FP: feature-propagation gather (three_nn + three_interpolate built into
// the rows) and head activations in the epilogue.
template <bool SA, int MINB, bool FP>
__global__ void __launch_bounds__(NTHR, MINB) chain2_kernel(const __grid_constant__ Chain2Args a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *A_hi = smem;
    uint8_t *A_lo = A_hi + (size_t)a.kmax8 * 2048;
    uint8_t *ring = A_lo + (size_t)a.kmax8 * 2048;
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(ring + (size_t)a.nst * STAGE_BYTES);
    uint64_t *bar_empty = bar_full + MAX_STAGES;
    uint64_t *bar_acc = bar_empty + MAX_STAGES;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar_acc + 1);
    int *s_nni = reinterpret_cast<int *>(s_tmem + 4);          // FP: [128][3] three_nn indices of the tile's rows
    float *s_nnw = reinterpret_cast<float *>(s_nni + TM * 3);  // FP: [128][3] interpolation weights

    constexpr int CW = MINB >= 3 ? 16 : 32;   // columns a worker handles at a time (register budget)
    static_assert(!FP || (!SA && MINB == 2), "the FP variant is a 2-CTA rows kernel");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc(s_tmem, a.tmem_cols);
    if (tid == NWORK) {
        for (int s = 0; s < MAX_STAGES; ++s) { tc::mbar_init(bar_full + s, 1); tc::mbar_init(bar_empty + s, 1); }
        tc::mbar_init(bar_acc, 1);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;
    const uint32_t a_hi0 = tc::smem_u32(A_hi), a_lo0 = tc::smem_u32(A_lo), ring0 = tc::smem_u32(ring);

    if (warp == NWORK / 32 + 1) {
        // =========================== producer warp: weight slices -> ring ===========================
        int slot = 0, round = 0;
        for (int u = 0; u < a.nunits; ++u) {
            const Unit &U = a.U[u];
            const int nk16 = U.K / 16;
            const uint32_t piece = (uint32_t)U.nc * 16u;                               // one (kc, hi|lo) block of the chunk
            const size_t slab = (size_t)U.Nfull * 16;                                  // one (kc, hi|lo) slab of the image
            const bool whole = U.nc == U.Nfull;                                        // a stage is one contiguous slice
            for (int j = 0; j < U.nchunks; ++j) {
                const uint8_t *src = reinterpret_cast<const uint8_t *>(U.Wimg) + (size_t)(U.n0 + j * U.nc) * 16;
                int left = nk16;
                for (int i = 0; i < U.nsl; ++i, left -= U.ksl) {
                    if (round > 0) tc::mbar_wait(bar_empty + slot, (uint32_t)((round - 1) & 1));
                    if (elect_one()) {
                        const int steps = min(U.ksl, left);
                        const uint32_t bytes = (uint32_t)steps * 4u * piece;
                        mbar_expect_tx(bar_full + slot, bytes);
                        const uint32_t dst = ring0 + (uint32_t)slot * STAGE_BYTES;
                        if (whole) {
                            bulk_g2s(dst, src, bytes, bar_full + slot);
                        } else {
                            for (int p = 0; p < steps * 4; ++p)                        // p = kc_local * 2 + (hi|lo)
                                bulk_g2s(dst + (uint32_t)p * piece, src + (size_t)p * slab, piece, bar_full + slot);
                        }
                    }
                    __syncwarp();
                    src += (size_t)U.ksl * 4 * slab;
                    if (++slot == a.nst) { slot = 0; ++round; }
                }
            }
        }
    } else if (warp == NWORK / 32) {
        // =========================== MMA warp ===========================
        int slot = 0, round = 0;
        work_sync();                                          // operand of unit 0 gathered
        for (int u = 0; u < a.nunits; ++u) {
            const Unit &U = a.U[u];
            const int nk16 = U.K / 16;
            CHAIN_TRACE(40 + 2 * u);
            const uint32_t idesc = tc::instr_desc_f16(TM, U.nc);
            const uint32_t sslab = 2u * (uint32_t)U.nc * 16u;
            // descriptors differ only in the start-address field (address >> 4, low word): bases + 32-bit offsets
            const uint64_t ah0 = tc::smem_desc(a_hi0, 2048u, 128u), al0 = tc::smem_desc(a_lo0, 2048u, 128u);
            const uint64_t bh00 = tc::smem_desc(ring0, sslab, 128u);
            const uint32_t accB = (uint32_t)(U.G * U.nc);
            tc::fence_after_sync();
            for (int j = 0; j < U.nchunks; ++j) {
                const uint32_t tbase = tmem + (uint32_t)(j * (U.G + 1) * U.nc);
                int kk = 0;                                   // k-step; gnext = first k-step of the next hi*hi accumulator
                int g = 0, gnext = U.kb[0];
                uint32_t freshA = 1u;                         // the next hi*hi MMA starts its accumulator
                for (int i = 0; i < U.nsl; ++i) {
                    tc::mbar_wait(bar_full + slot, (uint32_t)(round & 1));
                    // the bookkeeping below is warp-uniform; only the elected lane issues
                    const bool lead = elect_one();
                    const int steps = min(U.ksl, nk16 - kk);
                    uint64_t bh = bh00 + (uint64_t)(uint32_t)(slot * (STAGE_BYTES >> 4));
                    uint64_t ah = ah0 + (uint64_t)(uint32_t)(kk * 256), al = al0 + (uint64_t)(uint32_t)(kk * 256);   // 2 * 2048 / 16
                    for (int s = 0; s < steps; ++s, ++kk, ah += 256, al += 256, bh += sslab >> 3) {
                        if (kk == gnext) { ++g; gnext = U.kb[g]; freshA = 1u; }
                        if (lead) {
                            tc::mma_f16(tbase + (uint32_t)(g * U.nc), ah, bh, idesc, freshA ^ 1u);
                            tc::mma_f16(tbase + accB, ah, bh + (uint64_t)U.nc, idesc, kk > 0);
                            tc::mma_f16(tbase + accB, al, bh, idesc, 1u);
                        }
                        freshA = 0u;
                    }
                    if (lead) tc::mma_commit(bar_empty + slot);   // the stage may be refilled once these MMAs have read it
                    __syncwarp();
                    if (++slot == a.nst) { slot = 0; ++round; }
                }
            }
            if (elect_one()) tc::mma_commit(bar_acc);         // accumulators of the unit complete
            __syncwarp();
            CHAIN_TRACE(41 + 2 * u);
            tc::fence_before_sync();
            work_sync();                                      // epilogue done: operand rewritten, TMEM drained
        }
    } else {
        // =========================== workers ===========================
        const int r = tid & (TM - 1), h = tid >> 7, wq = warp & 3;
        if (warp == 0) CHAIN_TRACE(0);
        long R;                 // global row (rows mode) / row inside the cloud (SA mode)
        int b = 0;
        if (SA) {
            b = blockIdx.y;
            R = (long)blockIdx.x * TM + r;
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
            if (a.N0) {
                // xyz-only input (layer1): the first conv has 3 input channels -- 3 FMAs per output on the CUDA cores in
                // exact f32 instead of a 16-deep tensor-core step plus a whole epilogue round trip through TMEM
                for (int c0 = h * CW; c0 < a.N0; c0 += 2 * CW) {
                    float v[CW];
#pragma unroll
                    for (int i = 0; i < CW / 4; ++i) {
                        const float4 bb = __ldg(reinterpret_cast<const float4 *>(a.b0 + c0) + i);
                        const float4 w0 = __ldg(reinterpret_cast<const float4 *>(a.W0 + c0) + i);
                        const float4 w1 = __ldg(reinterpret_cast<const float4 *>(a.W0 + a.N0 + c0) + i);
                        const float4 w2 = __ldg(reinterpret_cast<const float4 *>(a.W0 + 2 * a.N0 + c0) + i);
                        v[4 * i] = fmaf(rel[2], w2.x, fmaf(rel[1], w1.x, fmaf(rel[0], w0.x, bb.x)));
                        v[4 * i + 1] = fmaf(rel[2], w2.y, fmaf(rel[1], w1.y, fmaf(rel[0], w0.y, bb.y)));
                        v[4 * i + 2] = fmaf(rel[2], w2.z, fmaf(rel[1], w1.z, fmaf(rel[0], w0.z, bb.z)));
                        v[4 * i + 3] = fmaf(rel[2], w2.w, fmaf(rel[1], w1.w, fmaf(rel[0], w0.w, bb.w)));
                    }
                    if (a.relu0) {
#pragma unroll
                        for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    store_operand2<CW>(A_hi, A_lo, r, c0, v);
                }
            } else
            // channel order [features(C), xyz(3), zero pad] (pointnet_util.py:52-57)
            for (int kc = h; kc < a.K0 / 8; kc += 2) {
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
                split8(v, reinterpret_cast<uint4 *>(A_hi + (size_t)kc * 2048 + r * 16), reinterpret_cast<uint4 *>(A_lo + (size_t)kc * 2048 + r * 16));
            }
        } else {
            // ---- rows mode: [X1 or three_interpolate(points2) (C1) | X2 (C2) | zero pad] (pointnet_util.py:223-228) ----
            // Load mapping: 8 lanes per row (a 128-byte line per row and instruction), 4 rows per warp instruction; a warp
            // owns 16 rows.  (One thread per row -- the store-friendly mapping -- touches 32 cache lines per load
            // instruction: the L1 tag stage then bounds the gather, measured 38.8k of fa_layer3's 122.8k cycles per tile,
            // and it slows the co-resident CTA's epilogues down as well.)  Lane (g, j) of the warp holds channels
            // [32 i + 4 j, +4) of row 16 warp + 4 grp + g and stores them as 8-byte halves of the 16-byte operand pieces.
            const long row0 = (long)blockIdx.x * TM;
            const long cloud = row0 / a.rows_per_cloud;
            R = row0 + r;
            if (FP) {
                // (idx, weight) of the tile's rows from the stage's three_nn table (ops.cu three_nn_kernel)
                for (int i = tid; i < TM * 3; i += NWORK) {
                    s_nni[i] = __ldg(a.fp_idx + (size_t)row0 * 3 + i);
                    s_nnw[i] = __ldg(a.fp_w + (size_t)row0 * 3 + i);
                }
                asm volatile("bar.sync 2, %0;" ::"n"(NWORK) : "memory");
            }
            const int g = lane >> 3, j = lane & 7;
            const int nseg = (a.K0 + 31) / 32;
            const bool x2vec = a.X2 && (a.C2 & 3) == 0 && (a.C1 & 3) == 0;
            for (int grp = 0; grp < 4; ++grp) {
                const int rr = warp * 16 + grp * 4 + g;                  // row of the tile this lane works on
                const long Rr = row0 + rr;
                const float *p1, *p2 = nullptr, *p3 = nullptr;
                float w1 = 0.f, w2 = 0.f, w3 = 0.f;
                if (FP) {
                    w1 = s_nnw[rr * 3]; w2 = s_nnw[rr * 3 + 1]; w3 = s_nnw[rr * 3 + 2];
                    p1 = a.fp_points2 + ((size_t)cloud * a.fp_m2 + s_nni[rr * 3 + 0]) * a.C1;
                    p2 = a.fp_points2 + ((size_t)cloud * a.fp_m2 + s_nni[rr * 3 + 1]) * a.C1;
                    p3 = a.fp_points2 + ((size_t)cloud * a.fp_m2 + s_nni[rr * 3 + 2]) * a.C1;
                } else {
                    p1 = a.X1 + (size_t)Rr * a.C1;
                }
                const float *q2 = a.X2 ? a.X2 + (size_t)Rr * a.C2 : nullptr;
                for (int s0 = 0; s0 < nseg; s0 += 4) {                    // 4 segments (128 channels) per trip: up to 12 loads in flight
                    float4 u[4], v[4], w[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int c = (s0 + t) * 32 + 4 * j;
                        if (s0 + t < nseg && c + 4 <= a.C1) {
                            u[t] = ldg4(p1 + c);
                            if (FP) { v[t] = ldg4(p2 + c); w[t] = ldg4(p3 + c); }
                        } else if (s0 + t < nseg && x2vec && c >= a.C1 && c + 4 <= a.C1 + a.C2) {
                            u[t] = ldg4(q2 + (c - a.C1));
                        }
                    }
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int c = (s0 + t) * 32 + 4 * j;
                        if (s0 + t >= nseg || c >= a.K0) continue;
                        float x[4];
                        if (c + 4 <= a.C1) {
                            if (FP) {
                                x[0] = interp3_unfused(u[t].x, v[t].x, w[t].x, w1, w2, w3);
                                x[1] = interp3_unfused(u[t].y, v[t].y, w[t].y, w1, w2, w3);
                                x[2] = interp3_unfused(u[t].z, v[t].z, w[t].z, w1, w2, w3);
                                x[3] = interp3_unfused(u[t].w, v[t].w, w[t].w, w1, w2, w3);
                            } else {
                                x[0] = u[t].x; x[1] = u[t].y; x[2] = u[t].z; x[3] = u[t].w;
                            }
                        } else if (x2vec && c >= a.C1 && c + 4 <= a.C1 + a.C2) {
                            x[0] = u[t].x; x[1] = u[t].y; x[2] = u[t].z; x[3] = u[t].w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const int ce = c + e;
                                x[e] = ce < a.C1 ? (FP ? 0.f : __ldg(p1 + ce)) : (q2 && ce < a.C1 + a.C2 ? __ldg(q2 + (ce - a.C1)) : 0.f);
                            }
                        }
                        // 4 channels -> 8-byte halves of the hi and lo pieces (same split as split8)
                        const __half2 h0 = __floats2half2_rn(x[0], x[1]), h1 = __floats2half2_rn(x[2], x[3]);
                        const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
                        const __half2 l0 = __floats2half2_rn((x[0] - b0.x) * tc::LO_SCALE, (x[1] - b0.y) * tc::LO_SCALE);
                        const __half2 l1 = __floats2half2_rn((x[2] - b1.x) * tc::LO_SCALE, (x[3] - b1.y) * tc::LO_SCALE);
                        const size_t off = (size_t)(c >> 3) * 2048 + rr * 16 + (j & 1) * 8;
                        *reinterpret_cast<uint2 *>(A_hi + off) = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
                        *reinterpret_cast<uint2 *>(A_lo + off) = make_uint2(*reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
                    }
                }
            }
        }
        tc::fence_proxy_async();
        if (warp == 0) CHAIN_TRACE(1);
        work_sync();                                          // operand of unit 0 gathered
        const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);
        for (int u = 0; u < a.nunits; ++u) {
            const Unit &U = a.U[u];
            tc::mbar_wait(bar_acc, (uint32_t)(u & 1));
            if (warp == 0) CHAIN_TRACE(2 + 2 * u);
            tc::fence_after_sync();
            const float *bias = U.bias_stride ? U.bias + (size_t)(R / a.rows_per_cloud) * U.bias_stride : U.bias;
            const int groups = U.nc / CW;                     // CW-column groups per chunk
            if constexpr (FP) {
                if (U.act) {
                    // packed head layer (64 columns, one chunk): BOTH threads of the row read the first 32 columns and share
                    // the activations; the thread of the second half also reads column 32 (the confidence when K = 4, mixed)
                    float x[32];
                    load_acc2(trow, U.G, 0, x, U.nc);
                    const float4 *b4 = reinterpret_cast<const float4 *>(bias + U.n0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 bb = __ldg(b4 + i);
                        x[4 * i] = fmaf(x[4 * i], U.descale, bb.x); x[4 * i + 1] = fmaf(x[4 * i + 1], U.descale, bb.y);
                        x[4 * i + 2] = fmaf(x[4 * i + 2], U.descale, bb.z); x[4 * i + 3] = fmaf(x[4 * i + 3], U.descale, bb.w);
                    }
                    float confi32 = 0.f;
                    if (h == 1 && U.act == TC_ACT_NOCS_HEADS && a.mixed && a.n_parts == 4) {
                        float y[32];
                        load_acc2(trow, U.G, 32, y, U.nc);
                        confi32 = fmaf(y[0], U.descale, __ldg(bias + U.n0 + 32));
                    }
                    head_activations(U.act, h, x, confi32, R, a.n_parts, a.mixed, a.pred);
                    tc::fence_proxy_async();
                    tc::fence_before_sync();
                    if (warp == 0) CHAIN_TRACE(3 + 2 * u);
                    work_sync();
                    continue;
                }
            }
            for (int j = 0; j < U.nchunks; ++j) {
                for (int q = 0; q < groups; ++q) {
                    if (((j * groups + q) & 1) != h) continue;        // the two halves alternate column groups
                    const int c0 = q * CW;
                    float v[CW];
                    load_acc2(trow + (uint32_t)(j * (U.G + 1) * U.nc), U.G, c0, v, U.nc);
                    const int col = U.n0 + j * U.nc + c0;
                    const float4 *b4 = reinterpret_cast<const float4 *>(bias + col);
#pragma unroll
                    for (int i = 0; i < CW / 4; ++i) {
                        const float4 bb = __ldg(b4 + i);
                        v[4 * i] = fmaf(v[4 * i], U.descale, bb.x); v[4 * i + 1] = fmaf(v[4 * i + 1], U.descale, bb.y);
                        v[4 * i + 2] = fmaf(v[4 * i + 2], U.descale, bb.z); v[4 * i + 3] = fmaf(v[4 * i + 3], U.descale, bb.w);
                    }
                    if (U.relu) {
#pragma unroll
                        for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    if (U.pool) {
                        const long g = ((long)blockIdx.x * TM + wq * 32) / a.S;
                        pool_store2<CW>(U.out + ((size_t)b * a.m + g) * U.Nfull + col, a.S, lane, v);
                    } else if (U.image) {
                        // tile image [Nfull/8][hi|lo][128][8] (net_tc.cuh, GemmImgArgs): consecutive rows -> consecutive
                        // 16-byte pieces, coalesced
                        uint8_t *img = reinterpret_cast<uint8_t *>(U.out) + (size_t)blockIdx.x * ((size_t)U.Nfull * 512) + r * 16;
#pragma unroll
                        for (int q8 = 0; q8 < CW / 8; ++q8) {
                            float w[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) w[i] = v[q8 * 8 + i];
                            uint8_t *p = img + (size_t)((col >> 3) + q8) * 4096;
                            split8(w, reinterpret_cast<uint4 *>(p), reinterpret_cast<uint4 *>(p + 2048));
                        }
                    } else if (U.out) {
                        float4 *o = reinterpret_cast<float4 *>(U.out + (size_t)R * U.ldo + col);
#pragma unroll
                        for (int i = 0; i < CW / 4; ++i) o[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                    if (U.inplace) store_operand2<CW>(A_hi, A_lo, r, col, v);
                }
            }
            tc::fence_proxy_async();
            tc::fence_before_sync();
            if (warp == 0) CHAIN_TRACE(3 + 2 * u);
            work_sync();
        }
        if (warp == 0) tc::tmem_dealloc(tmem, a.tmem_cols);
    }
}

// ---- host: unit list + launch ------------------------------------------------------------------------------------
struct LayerSpec {
    TcLayer L;
    int inplace;       // output becomes the next operand
    int pool;          // max-pool over the group (set abstraction's last layer)
    float *out;
    int ldo;
    const float *bias_override;
    long bias_stride;
    int act;
    int image;
};

int build_units(Chain2Args &a, const LayerSpec *spec, int nspec, size_t *smem_out, int *minb_out)
{
    *minb_out = 2;
    int kmax = a.K0;
    for (int i = 0; i < nspec; ++i) {
        const TcLayer &L = spec[i].L;
        // an in-place layer keeps all its chunks' accumulators live (<= 512 TMEM columns); an output layer is cut into
        // independent chunk units and may be as wide as the unit list allows
        if (!L.Wimg || L.K % 16 != 0 || L.N % 32 != 0 || L.N > (spec[i].inplace ? 256 : 1024)) return ANCSH_ERR_INVALID_ARG;
        if (i > 0 && spec[i - 1].inplace && L.K != spec[i - 1].L.N) return ANCSH_ERR_INVALID_ARG;
        kmax = L.K > kmax ? L.K : kmax;
        if (spec[i].inplace && L.N > kmax) kmax = L.N;
    }
    a.kmax8 = kmax / 8;
    const size_t opbytes = (size_t)2 * a.kmax8 * 2048;
    const size_t tail = (2 * MAX_STAGES + 1) * sizeof(uint64_t) + 16 + (a.fp_points2 ? (size_t)TM * 3 * 8 : 0);
    size_t limit = 113 * 1024;                                 // two CTAs per SM when the operand is small enough
    a.tmem_cols = 256;
    a.nst = DEF_STAGES;
    // Narrow chains (every accumulator set fits 128 TMEM columns when the layers are cut into 64-column chunks, operand
    // + ring <= 1/3 of the shared memory): three CTAs per SM -- these chains are bound by the latency of the
    // gather -> MMA -> epilogue sequence of one tile, not by any pipe, so residency is what buys throughput.
    static const bool narrow_off = getenv("ANCSH_CHAIN_NARROW_OFF") != nullptr;   // A/B switch for profiling
    bool narrow = !narrow_off && opbytes + (size_t)DEF_STAGES * STAGE_BYTES + tail <= 74 * 1024;
    for (int i = 0; i < nspec && narrow; ++i) {
        const TcLayer &L = spec[i].L;
        if (L.K > 144 || L.N % 32 != 0 || (spec[i].inplace ? L.N > 64 : (L.N > 64 && L.N % 64 != 0))) narrow = false;
    }
    if (narrow) { limit = 74 * 1024; a.tmem_cols = 128; *minb_out = 3; }
    if (opbytes + (size_t)DEF_STAGES * STAGE_BYTES + tail > limit) {
        limit = 227 * 1024;
        a.tmem_cols = 512;
        a.nst = MAX_STAGES;
        while (a.nst > 2 && opbytes + (size_t)a.nst * STAGE_BYTES + tail > limit) --a.nst;
    }
    const size_t smem = opbytes + (size_t)a.nst * STAGE_BYTES + tail;
    if (smem > limit) return ANCSH_ERR_UNSUPPORTED;
    a.nunits = 0;
    for (int i = 0; i < nspec; ++i) {
        const TcLayer &L = spec[i].L;
        const int nc = narrow ? (L.N >= 64 ? 64 : L.N) : (L.N >= 128 ? 128 : L.N);   // 32 / 64 / 128
        if (L.N % nc != 0) return ANCSH_ERR_UNSUPPORTED;
        const int nch = L.N / nc;
        const int together = spec[i].inplace ? nch : 1;        // chunks whose accumulators are live together
        if (together * 2 * nc > (int)a.tmem_cols) {            // e.g. a 256-wide in-place layer needs all 512 columns
            if (narrow) return ANCSH_ERR_UNSUPPORTED;
            if (limit == 113 * 1024) { a.tmem_cols = 512; }
            if (together * 2 * nc > (int)a.tmem_cols) return ANCSH_ERR_UNSUPPORTED;
        }
        for (int j = 0; j < nch; j += together) {
            if (a.nunits >= MAX_UNITS) return ANCSH_ERR_UNSUPPORTED;
            Unit &U = a.U[a.nunits++];
            U.Wimg = L.Wimg;
            U.bias = spec[i].bias_override ? spec[i].bias_override : L.bias;
            U.bias_stride = spec[i].bias_override ? spec[i].bias_stride : 0;
            U.out = spec[i].out; U.ldo = spec[i].ldo;
            U.K = L.K; U.Nfull = L.N; U.n0 = j * nc; U.nc = nc; U.nchunks = together;
            U.G = 1;                                           // filled below (needs the final tmem_cols)
            U.ksl = STAGE_BYTES / (64 * nc);
            U.nsl = (L.K / 16 + U.ksl - 1) / U.ksl;
            U.relu = L.relu; U.inplace = spec[i].inplace; U.pool = spec[i].pool; U.act = spec[i].act; U.image = spec[i].image;
            U.descale = L.descale;
        }
    }
    for (int u = 0; u < a.nunits; ++u) {
        Unit &U = a.U[u];
        U.G = tc_num_acc(U.K, (int)a.tmem_cols / (U.nc * U.nchunks) - 1);
        if (U.G > 8) return ANCSH_ERR_UNSUPPORTED;
        const int nk16 = U.K / 16;
        for (int g = 0; g < 8; ++g) U.kb[g] = g < U.G - 1 ? ((g + 1) * nk16 + U.G - 1) / U.G : 0x7FFFFFFF;
    }
    *smem_out = smem;
    return ANCSH_OK;
}

template <bool SA, int MINB, bool FP = false>
int launch_chain2(const Chain2Args &a, dim3 grid, size_t smem, cudaStream_t st)
{
    ANCSH_CUDA(cudaFuncSetAttribute(chain2_kernel<SA, MINB, FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ANCSH_CUDA(cudaFuncSetAttribute(chain2_kernel<SA, MINB, FP>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    chain2_kernel<SA, MINB, FP><<<grid, NTHR, smem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

}  // namespace

int sa_tc2_launch(const SaTcArgs &s, int B, cudaStream_t st)
{
    const long rows = (long)s.m * s.S;
    if (rows % TM != 0 || s.S % 32 != 0 || s.C % 8 != 0) return ANCSH_ERR_UNSUPPORTED;
    if (s.L[0].K < s.C + 3 || !s.L[2].relu) return ANCSH_ERR_INVALID_ARG;
    Chain2Args a{};
    a.xyz = s.xyz; a.points = s.points; a.new_xyz = s.new_xyz; a.idx = s.idx;
    a.n = s.n; a.m = s.m; a.S = s.S; a.C = s.C;
    a.rows_per_cloud = 1;
    LayerSpec spec[3] = {};
    int first = 0;
    if (s.C == 0 && s.W0 && s.L[0].N % 64 == 0 && s.L[1].K == s.L[0].N) {   // xyz-only: first conv in the gather
        a.W0 = s.W0; a.b0 = s.L[0].bias; a.N0 = s.L[0].N; a.relu0 = s.L[0].relu;
        first = 1;
    }
    a.K0 = first ? s.L[0].N : s.L[0].K;
    for (int l = first; l < 3; ++l) { spec[l - first].L = s.L[l]; spec[l - first].inplace = l < 2; }
    spec[2 - first].pool = 1; spec[2 - first].out = s.out;
    size_t smem = 0;
    int minb = 2;
    int rc = build_units(a, spec, 3 - first, &smem, &minb);
    if (rc) return rc;
    if (s.S != 32) ANCSH_CUDA(cudaMemsetAsync(s.out, 0, (size_t)B * s.m * s.L[2].N * sizeof(float), st));
    if (minb == 3) return launch_chain2<true, 3>(a, dim3((unsigned)(rows / TM), B), smem, st);
    return launch_chain2<true, 2>(a, dim3((unsigned)(rows / TM), B), smem, st);
}

int chain_tc2_launch(const ChainTcArgs &c, long rows_total, cudaStream_t st)
{
    if (rows_total % TM != 0 || c.C1 % 8 != 0 || c.nsteps < 1 || c.nsteps > 8) return ANCSH_ERR_UNSUPPORTED;
    if (c.S[0].L.K < c.C1 + (c.X2 ? c.C2 : 0)) return ANCSH_ERR_INVALID_ARG;
    Chain2Args a{};
    a.X1 = c.X1; a.X2 = c.X2; a.C1 = c.C1; a.C2 = c.C2;
    a.rows_per_cloud = c.rows_per_cloud > 0 ? c.rows_per_cloud : 1;
    a.K0 = c.S[0].L.K;
    const bool fp = c.fp_points2 != nullptr;
    bool heads = false;
    for (int i = 0; i < c.nsteps; ++i) heads = heads || c.S[i].act != TC_ACT_NONE;
    if (fp) {
        // a tile must not straddle clouds
        if (a.rows_per_cloud % TM != 0 || c.fp_m2 < 1 || !c.fp_idx || !c.fp_w) return ANCSH_ERR_INVALID_ARG;
        a.fp_points2 = c.fp_points2; a.fp_m2 = c.fp_m2; a.fp_idx = c.fp_idx; a.fp_w = c.fp_w;
    }
    if (heads) {
        // the epilogue keeps W | nocs | scale | trans of a row in one 32-column group (gocs = nocs * scale + trans)
        if (!fp || c.n_parts < 2 || (c.mixed ? 8 * c.n_parts : 4 * c.n_parts + 1) > 32) return ANCSH_ERR_UNSUPPORTED;
        a.pred = c.pred; a.n_parts = c.n_parts; a.mixed = c.mixed;
    }
    LayerSpec spec[8] = {};
    for (int i = 0; i < c.nsteps; ++i) {
        spec[i].L = c.S[i].L;
        spec[i].inplace = c.S[i].dst == TC_DST_INPLACE;
        spec[i].out = c.S[i].out; spec[i].ldo = c.S[i].ldo;
        spec[i].act = c.S[i].act;
        spec[i].image = c.S[i].dst == TC_DST_IMAGE;
        if (c.S[i].act != TC_ACT_NONE && (c.S[i].dst != TC_DST_GLOBAL || c.S[i].L.N != 64 || c.S[i].L.relu)) return ANCSH_ERR_INVALID_ARG;
        if (c.S[i].dst != TC_DST_INPLACE && !c.S[i].out && c.S[i].act == TC_ACT_NONE) return ANCSH_ERR_INVALID_ARG;
        if (c.S[i].dst == TC_DST_IMAGE && ((c.pool_S > 0 && i == c.nsteps - 1) || c.S[i].L.N % 16 != 0)) return ANCSH_ERR_INVALID_ARG;
        const bool pooled = c.pool_S > 0 && i == c.nsteps - 1;
        if (!pooled && c.S[i].out && c.S[i].dst != TC_DST_IMAGE && (c.S[i].ldo < c.S[i].L.N || c.S[i].ldo % 4 != 0)) return ANCSH_ERR_INVALID_ARG;
    }
    if (c.bias0) { spec[0].bias_override = c.bias0; spec[0].bias_stride = c.bias0_stride; }
    if (c.pool_S > 0) {
        // group max-pool of the last step (set abstraction with group_all, pointnet_util.py:66-91,156): the group of a
        // warp's 32 rows is (first row) / pool_S, groups never straddle a warp
        const ChainStep &last = c.S[c.nsteps - 1];
        if (c.pool_S % 32 != 0 || rows_total % c.pool_S != 0 || last.dst != TC_DST_GLOBAL || !last.L.relu) return ANCSH_ERR_INVALID_ARG;
        spec[c.nsteps - 1].pool = 1;
        a.S = c.pool_S; a.m = 0;
        if (c.pool_S != 32)
            ANCSH_CUDA(cudaMemsetAsync(last.out, 0, (size_t)(rows_total / c.pool_S) * last.L.N * sizeof(float), st));
    }
    size_t smem = 0;
    int minb = 2;
    int rc = build_units(a, spec, c.nsteps, &smem, &minb);
    if (rc) return rc;
    // profiling aid: ANCSH_CHAIN_TRACE=<file prefix> dumps the phase time stamps of the 12th .. 15th chain launch
    static const char *trace_path = getenv("ANCSH_CHAIN_TRACE");
    static int trace_calls = 0;
    const unsigned nblk = (unsigned)(rows_total / TM);
    long long *trace_dev = nullptr;
    int trace_id = -1;
    if (trace_path && ++trace_calls >= 12 && trace_calls < 16) {
        trace_id = trace_calls;
        if (cudaMalloc(&trace_dev, (size_t)nblk * CTRACE_SLOTS * sizeof(long long)) != cudaSuccess) trace_dev = nullptr;
        else cudaMemsetAsync(trace_dev, 0, (size_t)nblk * CTRACE_SLOTS * sizeof(long long), st);
    }
    a.trace = trace_dev;
    auto dump = [&]() {
        if (!trace_dev) return;
        std::vector<long long> hbuf((size_t)nblk * CTRACE_SLOTS);
        cudaStreamSynchronize(st);
        cudaMemcpy(hbuf.data(), trace_dev, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        cudaFree(trace_dev);
        char name[512];
        snprintf(name, sizeof name, "%s_%d.txt", trace_path, trace_id);
        if (FILE *f = fopen(name, "w")) {
            fprintf(f, "# chain launch %d: %u CTAs, %d units, fp %d, K0 %d; one line per sampled CTA: %d clock64 stamps\n", trace_id, nblk,
                    a.nunits, (int)fp, a.K0, CTRACE_SLOTS);
            for (unsigned c = 0; c < nblk; c += 13) {
                for (int k = 0; k < CTRACE_SLOTS; ++k) fprintf(f, "%lld ", hbuf[(size_t)c * CTRACE_SLOTS + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
    };
    if (fp) {
        if (minb != 2) return ANCSH_ERR_UNSUPPORTED;
        rc = launch_chain2<false, 2, true>(a, dim3(nblk), smem, st);
        dump();
        return rc;
    }
    rc = minb == 3 ? launch_chain2<false, 3>(a, dim3(nblk), smem, st) : launch_chain2<false, 2>(a, dim3(nblk), smem, st);
    dump();
    return rc;
}
