// common.cuh -- shared helpers for the ANCSH B200 hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ancsh_b200.h"

// every kernel launch of the library goes through this macro: it also feeds the diagnostic launch counter
// (ancsh_launch_count, a relaxed atomic; the only process-wide state of the library)
void ancsh_count_launch();
#define ANCSH_CHECK_LAUNCH()                                   \
    do {                                                       \
        ancsh_count_launch();                                  \
        cudaError_t e__ = cudaGetLastError();                  \
        if (e__ != cudaSuccess) return ANCSH_ERR_CUDA;         \
    } while (0)

#define ANCSH_CUDA(call)                                       \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return ANCSH_ERR_CUDA;         \
    } while (0)

static inline int ancsh_cdiv(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
