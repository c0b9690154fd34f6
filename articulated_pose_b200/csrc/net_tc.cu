// net_tc.cu -- streaming tcgen05 GEMM for the one layer whose operand does not fit the operand-resident chains
// (net_tc2.cu / net_lean.cu): layer3 conv_2 (group_all, K = 512, N = 1024, pointnet_util.py:94-161 with group_all) with
// the max over the cloud's rows in its epilogue.
//
// Both operands arrive as ready-made fp16 hi/lo images: the weights from the packer (weights.tc_image) and the rows from
// the epilogue of the chain that produced them (TC_DST_IMAGE, net_tc2.cu), which writes per 128-row tile
//     [K/8][hi | lo][128 rows][8] fp16          (tile stride K * 512 bytes)
// so one MMA k-step (16 k) of a tile is ONE contiguous 8 KB block.  One CTA = one 128-row tile x one 128-column chunk:
//   warps 0..3  epilogue: TMEM -> descale + bias + ReLU -> max over the warp's 32 rows -> atomicMax into the pooled row
//   warp 4      MMA: one thread issues 3 tcgen05.mma per k-step (hi*hi -> accumulator A; hi*lo + lo*hi -> accumulator B)
//   warp 5      producer: one thread streams 16 KB stages (8 KB rows + 4 x 2 KB weights) with cp.async.bulk into a
//               6-deep mbarrier ring
// 96 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM, one CTA's epilogue under the other's MMAs.
// The previous version converted the f32 rows in every CTA (8x redundantly, one per column chunk) and alternated
// load / convert / MMA with two buffers (ncu: tensor pipe 12% active).
//
// Numerics.  x = hi + lo * 2^-11 with fp16 pieces (22 bits).  The tensor core adds into its f32 accumulator with
// TRUNCATION, so the dominant hi*hi products and the small cross terms go to separate accumulators and the epilogue adds
// them in round-to-nearest f32 (tc_common.cuh).
#include "common.cuh"
#include "tc_common.cuh"
#include "net_tc.cuh"

namespace {

constexpr int TM = 128;                  // rows per CTA
constexpr int NCH = 128;                 // output columns per CTA = TMEM columns per accumulator
constexpr int NST = 6;                   // ring stages
constexpr int STAGE_A = 8192;            // rows of one k-step: [2 kc][hi|lo][128][8] fp16
constexpr int STAGE_W = 4 * NCH * 16;    // weights of one k-step: [2 kc][hi|lo][NCH][8] fp16
constexpr int STAGE = STAGE_A + STAGE_W;
constexpr int NTHR = 192;
constexpr size_t kGemmSmem = (size_t)NST * STAGE + (2 * NST + 1) * sizeof(uint64_t) + 16;

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

// max over the rows of a group (pointnet_util.py:134).  Values are >= 0 (ReLU), so the unsigned ordering of their bit
// patterns is the float ordering: one redux.sync per column; lane i keeps column i.
__device__ __forceinline__ void pool_store(float *orow, int S, int lane, const float (&v)[32])
{
    uint32_t keep = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const uint32_t mx = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(v[i]));
        if (lane == i) keep = mx;
    }
    if (S == 32) orow[lane] = __uint_as_float(keep);
    else atomicMax(reinterpret_cast<int *>(orow + lane), (int)keep);
}

__global__ void __launch_bounds__(NTHR, 2) gemm_img_kernel(const GemmImgArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *ring = smem;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)NST * STAGE);
    uint64_t *empty = full + NST;
    uint64_t *bar_acc = empty + NST;
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar_acc + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * NCH;
    const int nk16 = a.L.K / 16;
    if (warp == 0) tc::tmem_alloc(s_tmem, 2 * NCH);
    if (tid == 32) {
        for (int i = 0; i < NST; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
        tc::mbar_init(bar_acc, 1);
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *s_tmem;

    if (warp == 5) {
        if (lane == 0) {
            const uint8_t *A = reinterpret_cast<const uint8_t *>(a.Aimg) + (size_t)blockIdx.x * ((size_t)a.L.K * 512);
            const uint8_t *W = reinterpret_cast<const uint8_t *>(a.L.Wimg);
            const size_t wslab = (size_t)a.L.N * 16;                       // one (kc, hi|lo) slab of the weight image
            for (int s = 0; s < nk16; ++s) {
                const int st = s % NST;
                if (s >= NST) tc::mbar_wait(empty + st, (uint32_t)((s / NST - 1) & 1));
                mbar_expect_tx(full + st, STAGE);
                const uint32_t dst = tc::smem_u32(ring + (size_t)st * STAGE);
                bulk_g2s(dst, A + (size_t)s * STAGE_A, STAGE_A, full + st);
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    bulk_g2s(dst + STAGE_A + p * (NCH * 16), W + (size_t)(4 * s + p) * wslab + (size_t)n0 * 16, NCH * 16, full + st);
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            const uint32_t idesc = tc::instr_desc_f16(TM, NCH);
            for (int s = 0; s < nk16; ++s) {
                const int st = s % NST;
                tc::mbar_wait(full + st, (uint32_t)((s / NST) & 1));
                tc::fence_after_sync();
                const uint32_t base = tc::smem_u32(ring + (size_t)st * STAGE);
                const uint64_t ah = tc::smem_desc(base, 4096u, 128u), al = ah + (2048u >> 4);
                const uint64_t wh = tc::smem_desc(base + STAGE_A, 2u * NCH * 16u, 128u), wl = wh + ((NCH * 16u) >> 4);
                tc::mma_f16(tmem, ah, wh, idesc, s > 0);
                tc::mma_f16(tmem + NCH, ah, wl, idesc, s > 0);
                tc::mma_f16(tmem + NCH, al, wh, idesc, 1u);
                tc::mma_commit(empty + st);
            }
            tc::mma_commit(bar_acc);     // tracks every MMA issued before it
        }
    } else {
        tc::mbar_wait(bar_acc, 0);
        tc::fence_after_sync();
        const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
        const long g = ((long)blockIdx.x * TM + warp * 32) / a.pool_S;
        float *orow = a.out + (size_t)g * a.L.N + n0;
        for (int c0 = 0; c0 < NCH; c0 += 32) {
            uint32_t ra[32], rb[32];
            tc::tmem_ld32_issue(trow + c0, ra);
            tc::tmem_ld32_issue(trow + (uint32_t)NCH + c0, rb);
            tc::tmem_wait_ld();
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float acc = fmaf(__uint_as_float(rb[i]), tc::LO_UNSCALE, __uint_as_float(ra[i]));
                v[i] = fmaxf(fmaf(acc, a.L.descale, __ldg(a.L.bias + n0 + c0 + i)), 0.f);
            }
            pool_store(orow + c0, a.pool_S, lane, v);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 2 * NCH);
}

}  // namespace

int gemm_img_launch(const GemmImgArgs &a, long rows_total, cudaStream_t st)
{
    if (rows_total % TM != 0 || a.L.N % NCH != 0) return ANCSH_ERR_UNSUPPORTED;
    if (!a.Aimg || !a.L.Wimg || a.L.K % 16 != 0 || a.L.K < 16 || !a.L.relu || !a.out) return ANCSH_ERR_INVALID_ARG;
    if (a.pool_S < 32 || a.pool_S % 32 != 0 || rows_total % a.pool_S != 0) return ANCSH_ERR_INVALID_ARG;
    ANCSH_CUDA(cudaFuncSetAttribute(gemm_img_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem));
    ANCSH_CUDA(cudaFuncSetAttribute(gemm_img_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (a.pool_S != 32) ANCSH_CUDA(cudaMemsetAsync(a.out, 0, (size_t)(rows_total / a.pool_S) * a.L.N * sizeof(float), st));
    dim3 grid((unsigned)(rows_total / TM), (unsigned)(a.L.N / NCH));
    gemm_img_kernel<<<grid, NTHR, kGemmSmem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}
