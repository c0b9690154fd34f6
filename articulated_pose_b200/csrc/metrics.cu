// metrics.cu -- SURVEY 8(f) row 2: the heavy kernels of the metric scripts that follow the pose stage.
//   * ancsh_amodal_extent : per-part amodal box extents from predicted NOCS (evaluation/compute_miou.py:196-200)
//   * ancsh_box_iou_3d    : lib/d3_utils.py:55-69 iou_3d -- nres^3 sample grid over the joint bounding box of two
//                           oriented boxes, counts of samples inside both / either box (pts_inside_box, :40-53)
// Both are integer-counting / max-reduction work over data that fits on chip: one CTA per (cloud, part) or box pair,
// warp-shuffle reductions, no atomics on global memory, results deterministic.
#include "common.cuh"

namespace {

constexpr int MT = 256;

// ---------------------------------------------------------------------------------------------------------------------
// amodal extents.  scale_pred = 2 * max_i |nocs[i, 3j:3j+3] - 0.5| over the points whose argmax(mask) == j, all in
// f32 like NumPy evaluates it on the f32 h5 arrays (compute_miou.py:198-199); argmax keeps the first maximum
// (np.argmax, :187).  Empty part: NumPy's max over an empty axis raises (swallowed by the script's bare except, :250);
// here extent = NaN and count = 0.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MT) amodal_extent_kernel(int N, int K, const float *__restrict__ nocs,
                                                           const float *__restrict__ mask, float *__restrict__ extent,
                                                           int *__restrict__ count)
{
    const int b = blockIdx.x / K, j = blockIdx.x % K;
    const float *w = mask + (size_t)b * N * K;
    const float *q = nocs + (size_t)b * N * 3 * K + 3 * j;
    float m0 = 0.f, m1 = 0.f, m2 = 0.f;
    int cnt = 0;
    for (int i = threadIdx.x; i < N; i += MT) {
        const float *wi = w + (size_t)i * K;
        int best = 0;
        float bv = wi[0];
        for (int k = 1; k < K; ++k) {
            const float v = wi[k];
            if (v > bv) { bv = v; best = k; }
        }
        if (best == j) {
            const float *p = q + (size_t)i * 3 * K;
            m0 = fmaxf(m0, fabsf(p[0] - 0.5f));
            m1 = fmaxf(m1, fabsf(p[1] - 0.5f));
            m2 = fmaxf(m2, fabsf(p[2] - 0.5f));
            ++cnt;
        }
    }
    __shared__ float s_m[3][MT / 32];
    __shared__ int s_c[MT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m0 = fmaxf(m0, __shfl_xor_sync(0xFFFFFFFFu, m0, o));
        m1 = fmaxf(m1, __shfl_xor_sync(0xFFFFFFFFu, m1, o));
        m2 = fmaxf(m2, __shfl_xor_sync(0xFFFFFFFFu, m2, o));
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    }
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_m[0][wid] = m0; s_m[1][wid] = m1; s_m[2][wid] = m2; s_c[wid] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < MT / 32; ++k) {
            m0 = fmaxf(m0, s_m[0][k]); m1 = fmaxf(m1, s_m[1][k]); m2 = fmaxf(m2, s_m[2][k]);
            cnt += s_c[k];
        }
        float *e = extent + (size_t)blockIdx.x * 3;
        if (cnt == 0) {
            e[0] = e[1] = e[2] = __int_as_float(0x7fc00000);
        } else {
            e[0] = 2.f * m0; e[1] = 2.f * m1; e[2] = 2.f * m2;
        }
        count[blockIdx.x] = cnt;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// box IoU on a sample grid.  f64 throughout (the boxes reach iou_3d as f64, compute_miou.py:224-231).
// Grid coordinate i of an axis is np.linspace's: i * ((hi - lo) / (nres - 1)) + lo with two roundings (no FMA), the
// last one exactly hi.  A sample is inside a box iff 0 < (p - c4).u_k < u_k.u_k for the three edges u1 = c5 - c4,
// u2 = c7 - c4, u3 = c0 - c4 (d3_utils.py:43-52; strict comparisons).  Dot products are accumulated left to right
// without contraction.
// ---------------------------------------------------------------------------------------------------------------------
struct BoxFrame {
    double o[3], u[3][3], uu[3];
};

__device__ __forceinline__ double dot3_nofma(const double *a, const double *b)
{
    return __dadd_rn(__dadd_rn(__dmul_rn(a[0], b[0]), __dmul_rn(a[1], b[1])), __dmul_rn(a[2], b[2]));
}

__device__ void box_frame(const double *c, BoxFrame &f)
{
    for (int d = 0; d < 3; ++d) {
        f.o[d] = c[4 * 3 + d];
        f.u[0][d] = c[5 * 3 + d] - c[4 * 3 + d];
        f.u[1][d] = c[7 * 3 + d] - c[4 * 3 + d];
        f.u[2][d] = c[0 * 3 + d] - c[4 * 3 + d];
    }
    for (int k = 0; k < 3; ++k) f.uu[k] = dot3_nofma(f.u[k], f.u[k]);
}

__device__ __forceinline__ bool inside(const BoxFrame &f, const double *p)
{
    const double up[3] = {p[0] - f.o[0], p[1] - f.o[1], p[2] - f.o[2]};
    bool in = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double v = dot3_nofma(up, f.u[k]);
        in = in && (v > 0.0) && (v < f.uu[k]);
    }
    return in;
}

__global__ void __launch_bounds__(MT) box_iou_kernel(int nres, const double *__restrict__ bbox1,
                                                     const double *__restrict__ bbox2, double *__restrict__ iou,
                                                     int *__restrict__ inter_out, int *__restrict__ union_out)
{
    __shared__ BoxFrame s_f[2];
    __shared__ double s_lo[3], s_step[3], s_hi[3];
    __shared__ int s_i[MT / 32], s_u[MT / 32];
    const double *c1 = bbox1 + (size_t)blockIdx.x * 24, *c2 = bbox2 + (size_t)blockIdx.x * 24;
    if (threadIdx.x < 2) box_frame(threadIdx.x == 0 ? c1 : c2, s_f[threadIdx.x]);
    if (threadIdx.x >= 32 && threadIdx.x < 35) {
        const int d = threadIdx.x - 32;
        double lo = c1[d], hi = c1[d];
        for (int k = 0; k < 8; ++k) {
            lo = fmin(lo, fmin(c1[3 * k + d], c2[3 * k + d]));
            hi = fmax(hi, fmax(c1[3 * k + d], c2[3 * k + d]));
        }
        s_lo[d] = lo; s_hi[d] = hi;
        s_step[d] = nres > 1 ? (hi - lo) / (double)(nres - 1) : 0.0;
    }
    __syncthreads();
    const BoxFrame f1 = s_f[0], f2 = s_f[1];
    const int total = nres * nres * nres;
    int ni = 0, nu = 0;
    for (int g = threadIdx.x; g < total; g += MT) {
        const int iz = g % nres, iy = (g / nres) % nres, ix = g / (nres * nres);
        const int id[3] = {ix, iy, iz};
        double p[3];
#pragma unroll
        for (int d = 0; d < 3; ++d)
            p[d] = (id[d] == nres - 1 && nres > 1) ? s_hi[d] : __dadd_rn(__dmul_rn((double)id[d], s_step[d]), s_lo[d]);
        const bool a = inside(f1, p), b = inside(f2, p);
        ni += (a && b);
        nu += (a || b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ni += __shfl_xor_sync(0xFFFFFFFFu, ni, o);
        nu += __shfl_xor_sync(0xFFFFFFFFu, nu, o);
    }
    if ((threadIdx.x & 31) == 0) { s_i[threadIdx.x >> 5] = ni; s_u[threadIdx.x >> 5] = nu; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < MT / 32; ++k) { ni += s_i[k]; nu += s_u[k]; }
        iou[blockIdx.x] = nu == 0 ? 1.0 : (double)ni / (double)nu;     // union == 0 -> 1 (d3_utils.py:66-67)
        if (inter_out) inter_out[blockIdx.x] = ni;
        if (union_out) union_out[blockIdx.x] = nu;
    }
}

}  // namespace

extern "C" int ancsh_amodal_extent(int B, int N, int K, const float *nocs, const float *mask, float *extent, int *count,
                                   void *stream)
{
    if (B < 0 || N <= 0 || K <= 0 || K > 64) return ANCSH_ERR_INVALID_ARG;
    if (B == 0) return ANCSH_OK;
    if (!nocs || !mask || !extent || !count) return ANCSH_ERR_INVALID_ARG;
    amodal_extent_kernel<<<B * K, MT, 0, (cudaStream_t)stream>>>(N, K, nocs, mask, extent, count);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

extern "C" int ancsh_box_iou_3d(int npairs, int nres, const double *bbox1, const double *bbox2, double *iou, int *inter,
                                int *uni, void *stream)
{
    if (npairs < 0 || nres < 1 || nres > 1024) return ANCSH_ERR_INVALID_ARG;
    if (npairs == 0) return ANCSH_OK;
    if (!bbox1 || !bbox2 || !iou) return ANCSH_ERR_INVALID_ARG;
    box_iou_kernel<<<npairs, MT, 0, (cudaStream_t)stream>>>(nres, bbox1, bbox2, iou, inter, uni);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}
