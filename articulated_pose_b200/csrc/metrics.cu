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


// ---------------------------------------------------------------------------------------------------------------------
// joint parameter voting (evaluation/eval_joint_params.py:178-190).  Per cloud and joint j = 1..K-1 over the points
// whose argmax(index_per_point) == j:
//     joint axis  = median(joint_axis_per_point[idx], axis=0)
//     joint point = median((gn + unitvec * (1 - heatmap) * thres_r)[idx], axis=0)
// with gn[i] = gocs_per_point[i, 3c:3c+3], c = argmax(instance_per_point[i]) (:160-166; gn_width == 3: gocs[i, :3]).
// All f32 like NumPy on the f32 h5 arrays: ((u * (1 - h)) * 0.2f) then the sum, un-contracted; np.median = middle element or
// the f32 mean (a + b) / 2 of the two middle ones; an empty joint gives NaN (np.median of an empty slice).
// One CTA per cloud; the six value columns of a joint are sorted together by a bitonic network in shared memory.
// ---------------------------------------------------------------------------------------------------------------------
struct VoteArgs {
    int N, K, gn_width, n_index;
    const float *gocs, *mask, *unitvec, *heatmap, *joint_axis, *index;
    float thres_r;
    float *axis_out, *pt_out;
    int *count;
};

__global__ void __launch_bounds__(MT) joint_vote_kernel(const VoteArgs a)
{
    extern __shared__ float s_val[];                    // 6 * npow2
    __shared__ int s_n;
    const int N = a.N, K = a.K, b = blockIdx.x, tid = threadIdx.x;
    int npow2 = 1;
    while (npow2 < N) npow2 <<= 1;
    for (int j = 1; j < K; ++j) {
        __syncthreads();
        if (tid == 0) s_n = 0;
        for (int i = tid; i < 6 * npow2; i += MT) s_val[i] = __int_as_float(0x7f800000);   // +inf padding sorts last
        __syncthreads();
        for (int i = tid; i < N; i += MT) {
            const size_t g = (size_t)b * N + i;
            const float *q = a.index + g * a.n_index;
            int jc = 0;
            float bv = q[0];
            for (int k = 1; k < a.n_index; ++k)
                if (q[k] > bv) { bv = q[k]; jc = k; }
            if (jc != j) continue;
            int c = 0;
            if (a.gn_width != 3) {
                const float *m = a.mask + g * K;
                float mv = m[0];
                for (int k = 1; k < K; ++k)
                    if (m[k] > mv) { mv = m[k]; c = k; }
            }
            const float *gn = a.gocs + g * a.gn_width + 3 * c;
            const float om = __fsub_rn(1.0f, a.heatmap[g]);
            const int p = atomicAdd(&s_n, 1);           // order is irrelevant: the columns are sorted independently
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                s_val[d * npow2 + p] = a.joint_axis[g * 3 + d];
                const float off = __fmul_rn(__fmul_rn(a.unitvec[g * 3 + d], om), a.thres_r);
                s_val[(3 + d) * npow2 + p] = __fadd_rn(gn[d], off);
            }
        }
        __syncthreads();
        const int nj = s_n;
        for (int k = 2; k <= npow2; k <<= 1)
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                for (int t = tid; t < 6 * npow2; t += MT) {
                    const int c = t / npow2, i = t - c * npow2, ixj = i ^ jj;
                    if (ixj > i) {
                        float *v = s_val + c * npow2;
                        const bool up = (i & k) == 0;
                        const float x = v[i], y = v[ixj];
                        if ((x > y) == up) { v[i] = y; v[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        if (tid < 6) {
            const float *v = s_val + tid * npow2;
            float med;
            if (nj == 0) med = __int_as_float(0x7fc00000);
            else if (nj & 1) med = v[nj / 2];
            else med = __fdiv_rn(__fadd_rn(v[nj / 2 - 1], v[nj / 2]), 2.0f);
            float *o = (tid < 3 ? a.axis_out : a.pt_out) + ((size_t)b * (K - 1) + (j - 1)) * 3 + (tid % 3);
            *o = med;
        }
        if (tid == 0) a.count[(size_t)b * (K - 1) + (j - 1)] = nj;
    }
}

// ---- create_unit_data_from_hdf5's sampling / normalisation (lib/dataset.py:290-317, 346-372) -----------------------------------
// one thread per output point: gathers every per-point array through perm (positions in the cloud tiled up to at least
// num_points entries, dataset.py:290-317, i.e. index modulo n_total), scales the coordinates by the cloud's norm factor
// (:351), builds the one-hot part mask (:360) and the joint mask (:356-358); `rot` = the sapien joint_rpy rotation applied
// about 0.5 to the NOCS arrays and to the joint vectors (:369-377, f64 like np.dot with the f64 euler matrix).
struct UnitArgs {
    int n_max, num_points, K;
    const int *n_total, *perm;
    const float *norm;
    const double *rot;
    ancsh_unit_in_t in;
    ancsh_unit_out_t out;
};

__device__ __forceinline__ void rot3(const double *R, const float *v, float shift, float *o)
{
    // np.dot(v - shift, R.T) + shift
    const double a = (double)(v[0] - shift), b = (double)(v[1] - shift), c = (double)(v[2] - shift);
    for (int i = 0; i < 3; ++i) o[i] = (float)(a * R[3 * i] + b * R[3 * i + 1] + c * R[3 * i + 2] + (double)shift);
}

__global__ void __launch_bounds__(256) unit_data_kernel(const UnitArgs a)
{
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.num_points) return;
    const int nt = a.n_total[b];
    const int src = nt > 0 ? a.perm[(size_t)b * a.num_points + i] % nt : 0;
    const size_t s = (size_t)b * a.n_max + src, d = (size_t)b * a.num_points + i;
    const double *R = a.rot ? a.rot + (size_t)b * 9 : nullptr;
    const float nf = a.norm[b];
    for (int c = 0; c < 3; ++c) a.out.P[d * 3 + c] = __fmul_rn(a.in.pts[s * 3 + c], nf);
    const float cls = a.in.cls[s];
    a.out.cls_gt[d] = cls;
    if (a.out.mask_array) {
        const int k = (int)(signed char)(int)cls;                                   // cls_arr.astype(np.int8), :360
        for (int j = 0; j < a.K; ++j) a.out.mask_array[d * a.K + j] = (j == (k < 0 ? k + a.K : k)) ? 1.f : 0.f;
    }
    auto vec3 = [&](const float *in, float *out, float shift, bool rotate) {
        if (!in || !out) return;
        if (rotate && R) rot3(R, in + s * 3, shift, out + d * 3);
        else for (int c = 0; c < 3; ++c) out[d * 3 + c] = in[s * 3 + c];
    };
    vec3(a.in.nocs_p, a.out.nocs_gt, 0.5f, true);
    vec3(a.in.nocs_g, a.out.nocs_gt_g, 0.5f, true);
    vec3(a.in.unitvec, a.out.unitvec_gt, 0.f, true);
    vec3(a.in.orient, a.out.orient_gt, 0.f, true);
    if (a.in.heatmap && a.out.heatmap_gt) a.out.heatmap_gt[d] = a.in.heatmap[s];
    if (a.in.joint_cls) {
        const float jc = a.in.joint_cls[s];
        if (a.out.joint_cls_gt) a.out.joint_cls_gt[d] = jc;
        if (a.out.joint_cls_mask) a.out.joint_cls_mask[d] = jc > 0.f ? 1.f : 0.f;
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

extern "C" int ancsh_joint_vote(int B, int N, int K, int gn_width, int n_index, const float *gocs, const float *mask,
                                const float *unitvec, const float *heatmap, const float *joint_axis, const float *index_per_point,
                                float thres_r, float *axis_out, float *pt_out, int *count, void *stream)
{
    if (B < 0 || N <= 0 || N > 8192 || K < 2 || K > 64 || n_index < 1 || (gn_width != 3 && gn_width != 3 * K))
        return ANCSH_ERR_INVALID_ARG;
    if (B == 0) return ANCSH_OK;
    if (!gocs || !mask || !unitvec || !heatmap || !joint_axis || !index_per_point || !axis_out || !pt_out || !count)
        return ANCSH_ERR_INVALID_ARG;
    int npow2 = 1;
    while (npow2 < N) npow2 <<= 1;
    const size_t smem = (size_t)6 * npow2 * sizeof(float);
    ANCSH_CUDA(cudaFuncSetAttribute(joint_vote_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VoteArgs a{N, K, gn_width, n_index, gocs, mask, unitvec, heatmap, joint_axis, index_per_point, thres_r, axis_out, pt_out, count};
    joint_vote_kernel<<<B, MT, smem, (cudaStream_t)stream>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

extern "C" int ancsh_unit_data(int B, int n_max, int num_points, int n_parts, const int *n_total, const int *perm,
                               const float *norm_factor, const double *rot, const ancsh_unit_in_t *in, const ancsh_unit_out_t *out,
                               void *stream)
{
    if (B < 0 || n_max <= 0 || num_points <= 0 || n_parts < 1 || n_parts > 127 || !in || !out) return ANCSH_ERR_INVALID_ARG;
    if (B == 0) return ANCSH_OK;
    if (B > 65535) return ANCSH_ERR_UNSUPPORTED;
    if (!n_total || !perm || !norm_factor || !in->pts || !in->cls || !out->P || !out->cls_gt) return ANCSH_ERR_INVALID_ARG;
    UnitArgs a{n_max, num_points, n_parts, n_total, perm, norm_factor, rot, *in, *out};
    unit_data_kernel<<<dim3(ancsh_cdiv(num_points, 256), B), 256, 0, (cudaStream_t)stream>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}
