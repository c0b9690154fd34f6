// net_tc.cu -- streaming tcgen05 GEMM for the one layer whose operand does not fit the operand-resident chains
// (net_tc2.cu / net_lean.cu): layer3 conv_2 (group_all, K = 512, N = 1024) with the max over the cloud's rows in its
// epilogue.  One CTA = 128 threads = 128 rows (thread t owns row t for the loads and TMEM lane t in the epilogue); the
// operand is converted slice by slice (32 k) into fp16 hi/lo images, the pre-tiled weights are streamed from L2 with
// cp.async, outputs are produced in chunks of <= 128 columns.
//
// Numerics.  x = hi + lo with fp16 pieces (22 bits).  The tensor core adds into its f32 accumulator with TRUNCATION, so
// the error grows linearly with the number of MMAs chained into one accumulator (measured).  Therefore
//   * the dominant hi*hi products go to accumulator(s) A -- K/16 MMAs instead of 3K/16, and for K > 144 the k range is
//     split over up to 3 A accumulators when TMEM allows;
//   * the small cross terms hi*lo + lo*hi go to a separate accumulator B (their truncation error is 2^-11 smaller);
//   * the epilogue adds A.. + B in round-to-nearest f32.
#include "common.cuh"
#include "tc_common.cuh"
#include "net_tc.cuh"

namespace {

constexpr int TM = 128;      // rows per CTA == threads per CTA
constexpr int NCH = 128;     // default output columns per chunk = TMEM columns per accumulator (layer1 uses 64)

struct Engine {
    uint8_t *Wst;            // weight slice buffer: [kc][hi|lo][NC][8] fp16
    uint64_t *bar;
    uint32_t tmem, w0, phase;
    int tid;
    int nch;                 // chunk width = accumulator stride in TMEM columns
};

// copy the weight rows [n0, n0+NC) of k-groups [kc0, kc0+nkc) (hi and lo) into the slice buffer
__device__ __forceinline__ void stage_weights(const Engine &e, const __half *Wimg, int Nfull, int n0, int NC, int kc0, int nkc)
{
    const uint8_t *src = reinterpret_cast<const uint8_t *>(Wimg);
    if (NC == Nfull) {                                      // whole rows: the slice is one contiguous block
        const uint8_t *s0 = src + (size_t)kc0 * (2 * Nfull * 16);
        const uint32_t bytes = (uint32_t)nkc * 2u * NC * 16u;
        for (uint32_t off = e.tid * 16; off < bytes; off += TM * 16) cp_async16(e.Wst + off, s0 + off);
    } else {                                                // NC is a power of two (32 / 64 / 128)
        const int sh = 31 - __clz(NC);
        const int pieces = nkc * 2 * NC;                    // 16-byte pieces
        for (int p = e.tid; p < pieces; p += TM) {
            const int row = p & (NC - 1), part = (p >> sh) & 1, kc = p >> (sh + 1);
            cp_async16(e.Wst + ((size_t)p << 4),
                       src + (size_t)(kc0 + kc) * (2 * Nfull * 16) + (size_t)part * Nfull * 16 + (size_t)(n0 + row) * 16);
        }
    }
    cp_async_commit();
}

// MMAs of one staged slice (`steps` k-steps); a_hi / a_lo are the shared-memory addresses of the operand at the slice's
// first k-step.  hi*hi -> accumulator A_g, cross terms -> accumulator B.
__device__ __forceinline__ void issue_slice(uint32_t tmemA, uint32_t tmemB, int G, int nk16, int kk0, int steps, uint32_t a_hi,
                                            uint32_t a_lo, uint32_t w0, int NC, uint32_t &startedA, uint32_t &startedB,
                                            int nch)
{
    const uint32_t idesc = tc::instr_desc_f16(TM, NC);
    const uint32_t sslab = 2u * NC * 16u;
    for (int s = 0; s < steps; ++s) {
        const uint64_t ah = tc::smem_desc(a_hi + (uint32_t)(2 * s) * 2048u, 2048u, 128u);
        const uint64_t al = tc::smem_desc(a_lo + (uint32_t)(2 * s) * 2048u, 2048u, 128u);
        const uint64_t bh = tc::smem_desc(w0 + (uint32_t)(2 * s) * sslab, sslab, 128u);
        const uint64_t bl = tc::smem_desc(w0 + (uint32_t)(2 * s) * sslab + (uint32_t)NC * 16u, sslab, 128u);
        const int g = (kk0 + s) * G / nk16;
        tc::mma_f16(tmemA + (uint32_t)(g * nch), ah, bh, idesc, (startedA >> g) & 1u);
        startedA |= 1u << g;
        tc::mma_f16(tmemB, ah, bl, idesc, startedB);
        startedB = 1u;
        tc::mma_f16(tmemB, al, bh, idesc, 1u);
    }
}

// 32 accumulator columns of this thread's row: sum of the A accumulators and B, in round-to-nearest f32
__device__ __forceinline__ void load_acc(uint32_t trow, int G, int c0, float (&v)[32], int nch)
{
    if (G == 1) {                            // common case: A and B, both loads in flight before one wait
        uint32_t ra[32], rb[32];
        tc::tmem_ld32_issue(trow + c0, ra);
        tc::tmem_ld32_issue(trow + (uint32_t)nch + c0, rb);
        tc::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(__uint_as_float(rb[i]), tc::LO_UNSCALE, __uint_as_float(ra[i]));
        return;
    }
    tc::tmem_ld32(trow + c0, v);
    for (int g = 1; g <= G; ++g) {           // g == G is the cross-term accumulator B
        float u[32];
        tc::tmem_ld32(trow + (uint32_t)(g * nch) + c0, u);
        const float f = g == G ? tc::LO_UNSCALE : 1.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaf(u[i], f, v[i]);
    }
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

__device__ __forceinline__ Engine engine_setup(uint8_t *Wst, uint64_t *bar, uint32_t *s_tmem, uint32_t tmem_cols, int nch)
{
    const int tid = threadIdx.x;
    if ((tid >> 5) == 0) tc::tmem_alloc(s_tmem, tmem_cols);
    if (tid == 0) tc::mbar_init(bar, 1);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    Engine e;
    e.Wst = Wst; e.bar = bar; e.tmem = *s_tmem; e.w0 = tc::smem_u32(Wst); e.phase = 0; e.tid = tid; e.nch = nch;
    return e;
}


// ============================================================================================================
// Streaming GEMM: neither operand is resident.  Per GK-wide k slice every thread converts its row's GK f32 inputs to
// the fp16 hi/lo images, the weight slice arrives by cp.async.  blockIdx.y selects a 128-column chunk.
// Two slice buffers: the MMAs of slice i (committed to bar[i & 1]) run while the threads load and convert slice i + 1;
// a buffer is rewritten only after the commit of the slice that last used it has arrived.
constexpr int GK = 32;       // k elements per slice of the streaming GEMM (2 MMA k-steps)
constexpr size_t kGemmStageA = (size_t)2 * (GK / 8) * 2048;            // hi + lo operand images of one slice
constexpr size_t kGemmStageW = (size_t)(GK / 8) * 2 * NCH * 16;        // weight slice
constexpr size_t kGemmSmem = 2 * (kGemmStageA + kGemmStageW) + 32;

__global__ void __launch_bounds__(TM) gemm_tc_kernel(const GemmTcArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *Abuf = smem;                                   // [2][hi | lo][(GK/8) * 2048]
    uint8_t *Wbuf = smem + 2 * kGemmStageA;                 // [2][kGemmStageW]
    uint64_t *bar = reinterpret_cast<uint64_t *>(Wbuf + 2 * kGemmStageW);      // [2]
    uint32_t *s_tmem = reinterpret_cast<uint32_t *>(bar + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long R = (long)blockIdx.x * TM + tid;
    const int n0 = blockIdx.y * NCH;
    const int NC = min(NCH, a.L.N - n0);
    Engine e = engine_setup(Wbuf, bar, s_tmem, a.tmem_cols, NCH);
    if (tid == 0) tc::mbar_init(bar + 1, 1);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();

    const float *r1 = a.X1 + (size_t)R * a.C1;
    const float *r2 = a.X2 ? a.X2 + (size_t)R * a.C2 : nullptr;
    const int nk16 = a.L.K / 16;
    const int G = a.max_acc;
    uint32_t startedA = 0, startedB = 0;
    uint32_t phase[2] = {0, 0};
    int it = 0;
    for (int k16 = 0; k16 < nk16; k16 += GK / 16, ++it) {
        const int st = it & 1;
        const int steps = min(GK / 16, nk16 - k16);
        uint8_t *A_hi = Abuf + (size_t)st * kGemmStageA, *A_lo = A_hi + (GK / 8) * 2048;
        if (it >= 2) {                                      // the MMAs of slice it - 2 have read this buffer
            tc::mbar_wait(bar + st, phase[st]);
            phase[st] ^= 1;
        }
        e.Wst = Wbuf + (size_t)st * kGemmStageW;
        stage_weights(e, a.L.Wimg, a.L.N, n0, NC, k16 * 2, steps * 2);
#pragma unroll
        for (int kc = 0; kc < GK / 8; ++kc) {
            if (kc >= steps * 2) break;
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
            issue_slice(e.tmem, e.tmem + (uint32_t)(G * NCH), G, nk16, k16, steps, tc::smem_u32(A_hi), tc::smem_u32(A_lo),
                        tc::smem_u32(e.Wst), NC, startedA, startedB, NCH);
            tc::mma_commit(bar + st);
        }
    }
    // tcgen05.commit tracks every MMA issued before it: the last commit covers the whole accumulation
    if (it >= 1) tc::mbar_wait(bar + ((it - 1) & 1), phase[(it - 1) & 1]);
    tc::fence_after_sync();
    const uint32_t trow = e.tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < NC; c0 += 32) {
        float v[32];
        load_acc(trow, G, c0, v, NCH);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float x = fmaf(v[i], a.L.descale, __ldg(a.L.bias + n0 + c0 + i));
            v[i] = a.L.relu ? fmaxf(x, 0.f) : x;
        }
        if (a.pool_S == 0) {
            float4 *o = reinterpret_cast<float4 *>(a.out + (size_t)R * a.ldo + n0 + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        } else {
            const long g = ((long)blockIdx.x * TM + warp * 32) / a.pool_S;
            pool_store(a.out + (size_t)g * a.L.N + n0 + c0, a.pool_S, lane, v);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(e.tmem, a.tmem_cols);
}

inline uint32_t pow2_cols(int cols) { return cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512; }

}  // namespace

int gemm_tc_launch(const GemmTcArgs &a0, long rows_total, cudaStream_t st)
{
    GemmTcArgs a = a0;
    if (rows_total % TM != 0 || a.C1 % 8 != 0) return ANCSH_ERR_UNSUPPORTED;
    if (!a.L.Wimg || a.L.K % 16 != 0 || a.L.N % 32 != 0 || a.L.K < a.C1 + (a.X2 ? a.C2 : 0)) return ANCSH_ERR_INVALID_ARG;
    if (a.pool_S && (a.pool_S % 32 != 0 || !a.L.relu)) return ANCSH_ERR_INVALID_ARG;
    if (!a.pool_S && (a.ldo < a.L.N || a.ldo % 4 != 0)) return ANCSH_ERR_INVALID_ARG;
    a.max_acc = 1;      // one hi*hi + one cross-term accumulator = 256 columns: two CTAs per SM (3+1 accumulators were
                        // measured: 1.6x slower for 15% less error on layer3)
    a.tmem_cols = pow2_cols((a.max_acc + 1) * NCH);
    const size_t smem = kGemmSmem;
    ANCSH_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ANCSH_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    if (a.pool_S && a.pool_S != 32)
        ANCSH_CUDA(cudaMemsetAsync(a.out, 0, (size_t)(rows_total / a.pool_S) * a.L.N * sizeof(float), st));
    dim3 grid((unsigned)(rows_total / TM), (unsigned)((a.L.N + NCH - 1) / NCH));
    gemm_tc_kernel<<<grid, TM, smem, st>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}
