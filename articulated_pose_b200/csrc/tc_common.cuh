// tc_common.cuh -- thin inline-PTX layer for Blackwell tensor cores (sm_100a): tcgen05.mma with the accumulator in
// tensor memory (TMEM), shared-memory matrix descriptors, mbarrier completion, TMEM load for the epilogue.
//
// Operand layout used throughout (no swizzle, K-major "core matrices"): a [ROWS][K] fp16 operand is stored as
//     byte_addr(r, k) = (k / 8) * (ROWS * 16) + r * 16 + (k % 8) * 2
// i.e. for every group of 8 consecutive k the 16-byte pieces of all rows are contiguous.  A core matrix (8 rows x 16
// bytes) is then 128 contiguous bytes; descriptor strides: SBO (next 8 rows) = 128 B, LBO (next 8 k) = ROWS * 16 B.
// One thread per row writes 16-byte pieces with consecutive lanes on consecutive rows -> conflict-free stores.
//
// Split precision (hard part 3): a = a_hi + a_lo, b = b_hi + b_lo with hi = fp16(x),
// lo = fp16(x - hi); the product is a_hi*b_hi + a_hi*b_lo + a_lo*b_hi accumulated in f32 in TMEM (3 MMAs per k-step).
// fp16 keeps 11 significant bits per piece, so hi + lo carries ~22 bits (error ~2^-21 per product, f32-class); the
// bf16 split the survey proposed carries only ~16 bits and measured 1.2e-4 on the network outputs -- over the 1e-4
// parity bar.  Range: |x| must stay below 65504 (BN-normalised PointNet++ activations are O(1..100)).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- tensor memory ---------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols)   // one full warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)     // same warp that allocated
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier --------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// arrive on `bar` once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- descriptors -----------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, K-major (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);            // start address  [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // leading byte offset [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    return d;                                            // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor: D f32 (c_format 1), A/B fp16 (a_format = b_format = 0), both K-major, dense
// (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N)
{
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; one elected thread
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- epilogue: 32 consecutive f32 accumulator columns of this thread's TMEM lane ---------------------
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// issue only: the registers are valid after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- fp16 hi/lo split ---------------------------------------------------------------------------------
// x = hi + lo * 2^-11 with hi = fp16(x), lo = fp16((x - hi) * 2^11): the lo piece is stored SCALED by 2^11 (exact), so
// that it is a normal fp16 number whenever hi is (|x| >= 6.1e-5) -- unscaled it goes subnormal below |x| = 0.125 and the
// split degrades from ~22 to ~11 + log2(|x| / 6e-8) bits (measured: 6e-4 output error with hidden activations of 1e-3).
// Weights are split the same way (weights.tc_image), so the cross-term accumulator holds 2^11 * (hi*lo + lo*hi) and
// every epilogue folds the factor into the add it does anyway: acc_hh + LO_UNSCALE * acc_cross.
// Range contract: |v| < 65504.  Larger magnitudes saturate (hi = +-65504, lo = the clamped remainder) instead of turning
// into inf - inf = NaN: the element is clipped, the rest of the row stays exact.
constexpr float LO_SCALE = 2048.f;
constexpr float LO_UNSCALE = 1.f / 2048.f;
__device__ __forceinline__ __half f2h_sat(float v)
{
    unsigned short h;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return __ushort_as_half(h);
}
__device__ __forceinline__ void split_f16(float v, __half &hi, __half &lo)
{
    hi = f2h_sat(v);
    lo = f2h_sat((v - __half2float(hi)) * LO_SCALE);
}
// 8 consecutive k of one row -> one 16-byte piece each for the hi and lo operand images (fp16)
__device__ __forceinline__ void store_split8(const float (&v)[8], uint4 *dst_hi, uint4 *dst_lo)
{
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half h0, l0, h1, l1;
        split_f16(v[2 * i], h0, l0);
        split_f16(v[2 * i + 1], h1, l1);
        h[i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        l[i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
    }
    *dst_hi = make_uint4(h[0], h[1], h[2], h[3]);
    *dst_lo = make_uint4(l[0], l[1], l[2], l[3]);
}

}  // namespace tc
