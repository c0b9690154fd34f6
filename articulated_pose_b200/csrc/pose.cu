// pose.cu -- per-part RANSAC + joint-constrained nonlinear solve on the GPU (f64 on the FP64 pipe).
//
//   partition_kernel       argmax part labels, ordered per-part point lists, per-joint median axis
//                          (evaluation/parallel_ancsh_pose.py:238-247, 259-260, 295)
//   single_score_kernel    one thread per hypothesis: 3-point Kabsch/scale/translation + inlier count over the
//                          part's points staged in shared memory as f64            (:35-54)
//   single_refit_kernel    first-max hypothesis, its inlier mask, refit on the inliers incl. the O(n^2)
//                          pairwise-distance scale                                   (:28-32, d3_utils.py:223-246)
//   joint_init / joint_lm / joint_model kernels   per hypothesis: 3+3 samples, MINPACK-lmder LM as a SIMT state machine
//                          (pose_math.cuh / lm_tick.cuh), model assembly                                (:106-184)
//   joint_verify_kernel    inlier counts of every hypothesis, warp per hypothesis                     (:186-194)
//   joint_refit_kernel     first-max hypothesis, masks, block-cooperative LM refit on all inliers
//   umeyama_kernel         lib/aligning.py:580-622 (GT poses, compute_gt_pose.py:87)
#include <stdlib.h>
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "pose_math.cuh"

namespace {

constexpr int RT = 256;   // threads of the cooperative kernels
constexpr int RJ = 128;   // threads of joint_refit_kernel: ~240 registers (LM control code) -> two blocks per SM (64 threads x 4 blocks: same 1.1 ms -- the serial LM control of warp 0 bounds it)

// ------------------------------------------------------------------------------------------------
// block reduction of NV doubles per thread (fixed order -> deterministic); result broadcast to all threads.
// s_red: (NTH/32 + 1) * NV doubles.  Warp shuffles -> one partial per warp -> thread k adds the partials of value k ->
// everybody reads the NV totals (two barriers; the totals live in their own slice, so the next call may start at once).
// ------------------------------------------------------------------------------------------------
template <int NV, int NTH>
__device__ void block_sum(double *v, double *s_red)
{
    constexpr int NW = NTH / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, off);
        if (lane == 0) s_red[warp * NV + k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double x = 0.0;
        for (int w = 0; w < NW; ++w) x += s_red[w * NV + threadIdx.x];
        s_red[NW * NV + threadIdx.x] = x;
    }
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < NV; ++k) v[k] = s_red[NW * NV + k];
}

// ordered list of the indices i < n with mask[i] != 0 -> list[0 .. count); returns count (same in every thread).
// s_scan: NTH/32 ints.  Ends with a barrier.
template <int NTH>
__device__ int block_compact_indices(const unsigned char *mask, int n, int *list, int *s_scan)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = (n + NTH - 1) / NTH, i0 = min(n, tid * chunk), i1 = min(n, i0 + chunk);
    int c = 0;
    for (int i = i0; i < i1; ++i) c += mask[i] ? 1 : 0;
    int incl = c;
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < NTH / 32; ++w) {
        const int x = s_scan[w];
        if (w < warp) base += x;
        total += x;
    }
    int pos = base + incl - c;
    for (int i = i0; i < i1; ++i)
        if (mask[i]) list[pos++] = i;
    __syncthreads();
    return total;
}

// Per-thread partial sums over the unordered pairs {i, j} of the listed points (scale_pts, d3_utils.py:237-246):
// ab += |s_i - s_j| |t_i - t_j|, aa += |s_i - s_j|^2, bb += |t_i - t_j|^2.  Row r is paired with row n-1-r so that every
// thread walks the same number of pairs; |a||b| is evaluated as sqrt(|a|^2 |b|^2) (one f64 square root per pair).
__device__ __forceinline__ void pair_row(const double *S, const double *T, const int *list, int n, int r, double &ab, double &aa,
                                         double &bb)
{
    const double *si = S + 3 * list[r], *ti = T + 3 * list[r];
    const double s0 = si[0], s1 = si[1], s2 = si[2], t0 = ti[0], t1 = ti[1], t2 = ti[2];
    for (int c = r + 1; c < n; ++c) {
        const double *sj = S + 3 * list[c], *tj = T + 3 * list[c];
        const double d0 = s0 - sj[0], d1 = s1 - sj[1], d2 = s2 - sj[2];
        const double e0 = t0 - tj[0], e1 = t1 - tj[1], e2 = t2 - tj[2];
        const double A2 = d0 * d0 + d1 * d1 + d2 * d2, B2 = e0 * e0 + e1 * e1 + e2 * e2;
        ab += sqrt(A2 * B2); aa += A2; bb += B2;
    }
}
template <int NTH>
__device__ void pair_sums(const double *S, const double *T, const int *list, int n, double &ab, double &aa, double &bb)
{
    ab = 0.0; aa = 0.0; bb = 0.0;
    const int half = (n + 1) / 2;
    for (int r = threadIdx.x; r < half; r += NTH) {
        pair_row(S, T, list, n, r, ab, aa, bb);
        const int r2 = n - 1 - r;
        if (r2 != r) pair_row(S, T, list, n, r2, ab, aa, bb);
    }
}

__device__ __forceinline__ bool is_inlier(const double *R, double s, const double *t, const double *src, const double *tgt,
                                          double th2)
{
    // res = target - scale * (R @ source) - translation ; sqrt(sum(res^2)) < th   (parallel_ancsh_pose.py:51-52)
    double rs[3];
    pm::matvec3(R, src, rs);
    const double r0 = tgt[0] - s * rs[0] - t[0];
    const double r1 = tgt[1] - s * rs[1] - t[1];
    const double r2 = tgt[2] - s * rs[2] - t[2];
    return (r0 * r0 + r1 * r1) + r2 * r2 < th2;
}

// ================================================================================================
// partition
// ================================================================================================
struct PartitionArgs {
    const float *P, *nocs, *mask, *joint_axis;
    const int *joint_cls;
    int N, K;
    int *part_idx;
    float *part_src, *part_tgt;
    double *axis_med;
    int *part_count;
};

__global__ void __launch_bounds__(RT) partition_kernel(const PartitionArgs a)
{
    extern __shared__ unsigned char s_raw[];
    const int N = a.N, K = a.K, b = blockIdx.x, tid = threadIdx.x;
    int npow2 = 1;
    while (npow2 < N) npow2 <<= 1;
    unsigned char *s_label = s_raw;                                      // N
    int *s_cnt = reinterpret_cast<int *>(s_raw + ((N + 15) & ~15));      // RT * 8
    float *s_val = reinterpret_cast<float *>(s_cnt + RT * 8);            // 3 * npow2
    __shared__ int s_tot[8];
    __shared__ int s_nj;

    // labels = argmax(mask, axis=1), first maximum (np.argmax) -- parallel_ancsh_pose.py:238
    for (int i = tid; i < N; i += RT) {
        const float *m = a.mask + ((size_t)b * N + i) * K;
        int best = 0;
        float bv = m[0];
        for (int k = 1; k < K; ++k)
            if (m[k] > bv) { bv = m[k]; best = k; }
        s_label[i] = (unsigned char)best;
    }
    __syncthreads();
    // ordered compaction per part: each thread owns a contiguous chunk of points
    const int chunk = (N + RT - 1) / RT;
    const int i0 = tid * chunk, i1 = min(N, i0 + chunk);
    int cnt[8];
    for (int k = 0; k < 8; ++k) cnt[k] = 0;
    for (int i = i0; i < i1; ++i) cnt[s_label[i]]++;
    for (int k = 0; k < 8; ++k) s_cnt[k * RT + tid] = cnt[k];
    __syncthreads();
    if (tid < K) {                      // exclusive scan per part (RT entries, serial: tiny)
        int run = 0;
        for (int t = 0; t < RT; ++t) { int c = s_cnt[tid * RT + t]; s_cnt[tid * RT + t] = run; run += c; }
        s_tot[tid] = run;
        a.part_count[(size_t)b * K + tid] = run;
    }
    __syncthreads();
    int pos[8];
    for (int k = 0; k < 8; ++k) pos[k] = s_cnt[k * RT + tid];
    for (int i = i0; i < i1; ++i) {
        const int j = s_label[i];
        const size_t o = ((size_t)b * K + j) * N + pos[j]++;
        a.part_idx[o] = i;
        const float *np_ = a.nocs + ((size_t)b * N + i) * 3 * K + 3 * j;      // nocs_pred[partidx[j], 3j:3j+3], :259
        const float *pp = a.P + ((size_t)b * N + i) * 3;
        a.part_src[o * 3 + 0] = np_[0]; a.part_src[o * 3 + 1] = np_[1]; a.part_src[o * 3 + 2] = np_[2];
        a.part_tgt[o * 3 + 0] = pp[0]; a.part_tgt[o * 3 + 1] = pp[1]; a.part_tgt[o * 3 + 2] = pp[2];
    }
    // per-joint median of the predicted axis over the points with joint_cls == j (:295)
    for (int j = 1; j < K; ++j) {
        __syncthreads();
        if (tid == 0) s_nj = 0;
        for (int i = tid; i < 3 * npow2; i += RT) s_val[i] = __int_as_float(0x7f800000);   // +inf padding
        __syncthreads();
        for (int i = tid; i < N; i += RT) {
            if (a.joint_cls[(size_t)b * N + i] == j) {
                const int p = atomicAdd(&s_nj, 1);          // order is irrelevant: the values get sorted
                const float *ax = a.joint_axis + ((size_t)b * N + i) * 3;
                s_val[p] = ax[0]; s_val[npow2 + p] = ax[1]; s_val[2 * npow2 + p] = ax[2];
            }
        }
        __syncthreads();
        const int nj = s_nj;
        // bitonic sort of the first npj = 2^lgj >= nj entries of each component (the rest is +inf padding)
        int npj = 2, lgj = 1;
        while (npj < nj) { npj <<= 1; ++lgj; }
        for (int k = 2; k <= npj; k <<= 1)
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                for (int t = tid; t < 3 * npj; t += RT) {
                    const int c = t >> lgj, i = t & (npj - 1), ixj = i ^ jj;
                    if (ixj > i) {
                        float *v = s_val + c * npow2;
                        const bool up = (i & k) == 0;
                        const float x = v[i], y = v[ixj];
                        if ((x > y) == up) { v[i] = y; v[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        if (tid < 3) {
            const float *v = s_val + tid * npow2;
            double med;
            if (nj == 0) med = nan("");
            else if (nj & 1) med = (double)v[nj / 2];
            else med = ((double)v[nj / 2 - 1] + (double)v[nj / 2]) / 2.0;
            a.axis_med[((size_t)b * (K - 1) + (j - 1)) * 3 + tid] = med;
        }
    }
}

// ================================================================================================
// single-part RANSAC
// ================================================================================================
struct SingleArgs {
    const float *part_src, *part_tgt;   // (nprob, N, 3)
    const int *part_count;              // (nprob)
    const int *idx;                     // NULL or (nprob, niter, 3)
    int *scores;                        // (nprob, niter)
    int *best;                          // (nprob)
    int N, niter;
    double th2;
    unsigned long long seed;
    // refit outputs
    double *R, *s, *t;
    int *score_out;
    unsigned char *inliers;
    int *status;
};

__device__ __forceinline__ void load_part_f64(const float *src, const float *tgt, int n, double *s_src, double *s_tgt)
{
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x) { s_src[i] = (double)src[i]; s_tgt[i] = (double)tgt[i]; }
}

__device__ __forceinline__ void fetch_sample(const int *idx, unsigned long long seed, unsigned prob, unsigned hyp,
                                             unsigned stream, int niter, int n, int *out)
{
    if (idx) {
        const int *p = idx + ((size_t)prob * niter + hyp) * 3;
        for (int i = 0; i < 3; ++i) out[i] = min(max(p[i], 0), n - 1);
    } else {
        pm::sample3(seed, prob, hyp, stream, n, out);
    }
}

__global__ void __launch_bounds__(RT, 2) single_score_kernel(const SingleArgs a)
{
    extern __shared__ double s_pts[];
    const int prob = blockIdx.y;
    const int n = a.part_count[prob];
    double *s_src = s_pts, *s_tgt = s_pts + (size_t)a.N * 3;
    const int h = blockIdx.x * RT + threadIdx.x;
    if (n <= 0) {
        if (h < a.niter) a.scores[(size_t)prob * a.niter + h] = 0;
        return;
    }
    load_part_f64(a.part_src + (size_t)prob * a.N * 3, a.part_tgt + (size_t)prob * a.N * 3, n, s_src, s_tgt);
    __syncthreads();
    if (h >= a.niter) return;
    int id[3];
    fetch_sample(a.idx, a.seed, prob, h, 0u, a.niter, n, id);
    double S[9], T[9], R[9], sc, t[3];
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) { S[3 * i + c] = s_src[3 * id[i] + c]; T[3 * i + c] = s_tgt[3 * id[i] + c]; }
    pm::transform3(S, T, R, &sc, t);
    int cnt = 0;
    for (int i = 0; i < n; ++i) cnt += is_inlier(R, sc, t, s_src + 3 * i, s_tgt + 3 * i, a.th2) ? 1 : 0;
    a.scores[(size_t)prob * a.niter + h] = cnt;
}

// first maximum of an int score array (ransac keeps the FIRST best: strict '>' at parallel_ancsh_pose.py:28)
template <int NTH>
__device__ int block_first_argmax_int(const int *scores, int n, unsigned long long *s_key)
{
    unsigned long long best = 0ull;
    for (int h = threadIdx.x; h < n; h += NTH) {
        const unsigned long long key = ((unsigned long long)(unsigned)(scores[h] + 1) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)h);
        best = key > best ? key : best;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_down_sync(0xFFFFFFFFu, best, off);
        best = o > best ? o : best;
    }
    if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = best;
    __syncthreads();
    best = 0ull;
    for (int w = 0; w < NTH / 32; ++w) best = s_key[w] > best ? s_key[w] : best;
    __syncthreads();
    return (int)(0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull));
}

// Threads per block: the kernel alternates block-wide passes with thread 0's 3x3 SVDs (45 % of the warp time of a 256-thread
// block was spent at barriers behind them, profiles/r02_single_refit_sass_stalls.txt); smaller blocks -- more of them resident
// per SM -- overlap one block's serial section with the others' passes.
constexpr int RSR = 128;
__global__ void __launch_bounds__(RSR, 4) single_refit_kernel(const SingleArgs a)
{
    extern __shared__ double s_pts[];
    __shared__ double s_red[(RSR / 32 + 1) * 11];
    __shared__ unsigned long long s_key[RSR / 32];
    __shared__ double s_model[13];
    __shared__ int s_scan[RSR / 32];
    const int prob = blockIdx.x, tid = threadIdx.x;
    const int n = a.part_count[prob];
    double *s_src = s_pts, *s_tgt = s_pts + (size_t)a.N * 3;
    int *s_list = reinterpret_cast<int *>(s_pts + (size_t)a.N * 6);                 // indices of the inliers, ascending
    unsigned char *s_inl = reinterpret_cast<unsigned char *>(s_list + a.N);
    unsigned char *g_inl = a.inliers + (size_t)prob * a.N;
    double *oR = a.R + (size_t)prob * 9, *ot = a.t + (size_t)prob * 3;
    if (n <= 0) {
        for (int i = tid; i < a.N; i += RSR) g_inl[i] = 0;
        if (tid == 0) {
            for (int i = 0; i < 9; ++i) oR[i] = nan("");
            for (int i = 0; i < 3; ++i) ot[i] = nan("");
            a.s[prob] = nan("");
            a.score_out[prob] = 0;
            a.best[prob] = -1;
            a.status[prob] |= ANCSH_POSE_EMPTY_PART;
        }
        return;
    }
    load_part_f64(a.part_src + (size_t)prob * a.N * 3, a.part_tgt + (size_t)prob * a.N * 3, n, s_src, s_tgt);
    const int hbest = block_first_argmax_int<RSR>(a.scores + (size_t)prob * a.niter, a.niter, s_key);   // syncs inside
    if (tid == 0) {
        int id[3];
        fetch_sample(a.idx, a.seed, prob, hbest, 0u, a.niter, n, id);
        double S[9], T[9];
        for (int i = 0; i < 3; ++i)
            for (int c = 0; c < 3; ++c) { S[3 * i + c] = s_src[3 * id[i] + c]; T[3 * i + c] = s_tgt[3 * id[i] + c]; }
        pm::transform3(S, T, s_model, s_model + 9, s_model + 10);
        a.best[prob] = hbest;
    }
    __syncthreads();
    // inlier mask of the best hypothesis + sums over the inliers
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < a.N; i += RSR) {
        unsigned char in = 0;
        if (i < n) {
            in = is_inlier(s_model, s_model[9], s_model + 10, s_src + 3 * i, s_tgt + 3 * i, a.th2) ? 1 : 0;
            s_inl[i] = in;
            if (in) {
                for (int c = 0; c < 3; ++c) { acc[c] += s_src[3 * i + c]; acc[3 + c] += s_tgt[3 * i + c]; }
                acc[6] += 1.0;
            }
        }
        g_inl[i] = in;
    }
    block_sum<7, RSR>(acc, s_red);
    const int nin = (int)acc[6];
    if (nin == 0) {
        if (tid == 0) {
            for (int i = 0; i < 9; ++i) oR[i] = nan("");
            for (int i = 0; i < 3; ++i) ot[i] = nan("");
            a.s[prob] = nan("");
            a.score_out[prob] = 0;
            a.status[prob] |= ANCSH_POSE_NO_INLIERS;
        }
        return;
    }
    double ms[3], mt[3];
    for (int c = 0; c < 3; ++c) { ms[c] = acc[c] / nin; mt[c] = acc[3 + c] / nin; }
    // centre in place (transform_pts, d3_utils.py:226-227)
    for (int i = tid; i < n; i += RSR)
        for (int c = 0; c < 3; ++c) { s_src[3 * i + c] -= ms[c]; s_tgt[3 * i + c] -= mt[c]; }
    __syncthreads();
    // M = target_c^T source_c (9) and the pairwise sums of scale_pts over unordered inlier pairs (2)
    block_compact_indices<RSR>(s_inl, n, s_list, s_scan);
    double v[11];
    for (int k = 0; k < 9; ++k) v[k] = 0.0;
    for (int c = tid; c < nin; c += RSR) {
        const double *si = s_src + 3 * s_list[c], *ti = s_tgt + 3 * s_list[c];
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) v[3 * p + q] += ti[p] * si[q];
    }
    {
        double bb;
        pair_sums<RSR>(s_src, s_tgt, s_list, nin, v[9], v[10], bb);
    }
    block_sum<11, RSR>(v, s_red);
    if (tid == 0) {
        double R[9];
        pm::kabsch_rotation(v, R);
        const double scale = (2.0 * v[9]) / (2.0 * v[10] + 1e-6);           // all ordered pairs, d3_utils.py:241-245
        double rs[3];
        pm::matvec3(R, ms, rs);
        for (int i = 0; i < 9; ++i) oR[i] = R[i];
        a.s[prob] = scale;
        for (int c = 0; c < 3; ++c) ot[c] = mt[c] - scale * rs[c];          // mean(target - s R source), :233
        a.score_out[prob] = nin;
    }
}

// ================================================================================================
// joint RANSAC
// ================================================================================================
struct JointArgs {
    const float *part_src, *part_tgt;   // (nparts_total, N, 3)
    const int *part_count;              // (nparts_total)
    const double *axis_med;             // (nprob, 3)
    const int *idx0, *idx1;             // NULL or (nprob, niter, 3)
    double *scores;                     // (nprob, niter)
    int *nfev;                          // (nprob, niter)
    pm::JointModel *models;             // (nprob, niter) per-hypothesis models
    int nprob;
    int *tail_count;                    // work counter of joint_lm_kernel (next unclaimed hypothesis)
    int *best;                          // (nprob)
    int N, K, niter;
    double th2;
    unsigned long long seed;
    // refit outputs (nprob, ...)
    double *R0, *s0, *t0, *R1, *s1, *t1, *score_out;
    unsigned char *inl0, *inl1;
    int *status;                        // (B,K)
};

// problem p = b*(K-1) + (j-1) couples part (b,0) with part (b,j)
__device__ __forceinline__ void joint_parts(int prob, int K, int &pa, int &pb)
{
    const int b = prob / (K - 1), j = prob - b * (K - 1) + 1;
    pa = b * K;
    pb = b * K + j;
}

__device__ __forceinline__ void gather_joint_samples(const JointArgs &a, int prob, int h, int pa, int pb, int n0, int n1,
                                                     double *S0, double *T0, double *S1, double *T1)
{
    int i0[3], i1[3];
    fetch_sample(a.idx0, a.seed, prob, h, 1u, a.niter, n0, i0);
    fetch_sample(a.idx1, a.seed, prob, h, 2u, a.niter, n1, i1);
    const float *gs0 = a.part_src + (size_t)pa * a.N * 3, *gt0 = a.part_tgt + (size_t)pa * a.N * 3;
    const float *gs1 = a.part_src + (size_t)pb * a.N * 3, *gt1 = a.part_tgt + (size_t)pb * a.N * 3;
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) {
            S0[3 * i + c] = (double)gs0[3 * i0[i] + c]; T0[3 * i + c] = (double)gt0[3 * i0[i] + c];
            S1[3 * i + c] = (double)gs1[3 * i1[i] + c]; T1[3 * i + c] = (double)gt1[3 * i1[i] + c];
        }
}

// ---- joint estimation = three kernels --------------------------------------------------------------------------
//   joint_init_kernel   thread per hypothesis: 3+3 samples, scale_pts / centring / Kabsch start (:121-148) -> JointRec
//   joint_lm_kernel     the LM solves (:154-155) as a SIMT state machine (lm_tick.cuh): persistent warps, one solve
//                       per lane, every loop trip = one trial step for all lanes; a lane whose solve terminated takes
//                       the next unsolved record, so short solves (median 12 evaluations) and long ones (MINPACK's
//                       maxfev = 600) share warps without idling lanes
//   joint_model_kernel  thread per hypothesis: rotation vectors -> {R0,s0,t0,R1,s1,t1} (:156-184)
struct JointRec {
    double pts[36];                            // x0[9] y0[9] x1[9] y1[9]: centred sources, pre-scaled centred targets
    double x[6];                               // rotation vectors: Kabsch start, overwritten with the LM solution
    double fnorm;                              // |f(x)|
    double s0, s1;                             // scale_pts of each part (:121; never refined, :174)
    double mS0[3], mT0[3], mS1[3], mT1[3];     // sample means: t = mean(T - s R S) = mean(T) - s R mean(S)
    double par, delta, xnorm;                  // lmder state of a suspended solve (valid when iter > 1)
    int iter, nfev, njev;                      // iter == 1: fresh record
    int valid;
    double pad[2];
};
static_assert(sizeof(JointRec) == 512, "ancsh_pose_plan reserves 512 bytes per joint hypothesis");

constexpr int JIT = 128;   // threads per block of the init / model kernels

__global__ void __launch_bounds__(JIT) joint_init_kernel(const JointArgs a, JointRec *recs)
{
    const long t = (long)blockIdx.x * JIT + threadIdx.x;
    if (t >= (long)a.nprob * a.niter) return;
    const int prob = (int)(t / a.niter), h = (int)(t - (long)prob * a.niter);
    int pa, pb;
    joint_parts(prob, a.K, pa, pb);
    const int n0 = a.part_count[pa], n1 = a.part_count[pb];
    JointRec &r = recs[t];
    if (n0 <= 0 || n1 <= 0) { r.valid = 0; a.nfev[t] = 0; return; }
    double S0[9], T0[9], S1[9], T1[9];
    gather_joint_samples(a, prob, h, pa, pb, n0, n1, S0, T0, S1, T1);
    double S0c[9], T0c[9], S1c[9], T1c[9], s0, s1, R[9], x[6];
    pm::centre_scale3(S0, T0, S0c, T0c, &s0);
    pm::centre_scale3(S1, T1, S1c, T1c, &s1);
    pm::kabsch_points(S0c, T0c, 3, R);                         // :138-139
    pm::matrix_to_rotvec(R, x);                                // :147-148
    pm::kabsch_points(S1c, T1c, 3, R);
    pm::matrix_to_rotvec(R, x + 3);
    pm::SerialProb P;
    P.x0 = S0c; P.y0 = T0c; P.n0 = 3; P.x1 = S1c; P.y1 = T1c; P.n1 = 3;
    const double *u = a.axis_med + (size_t)prob * 3;
    P.u[0] = u[0]; P.u[1] = u[1]; P.u[2] = u[2];
    P.nj = 3.0;                                                // min(n0,n1) copies of the joint direction, :134
    r.fnorm = sqrt(P.cost(x));
    r.par = 0.0; r.delta = 0.0; r.xnorm = 0.0;
    r.iter = 1; r.nfev = 1; r.njev = 0;
    for (int i = 0; i < 9; ++i) { r.pts[i] = S0c[i]; r.pts[9 + i] = T0c[i]; r.pts[18 + i] = S1c[i]; r.pts[27 + i] = T1c[i]; }
    for (int i = 0; i < 6; ++i) r.x[i] = x[i];
    r.s0 = s0; r.s1 = s1;
    for (int c = 0; c < 3; ++c) {
        r.mS0[c] = (S0[c] + S0[3 + c] + S0[6 + c]) / 3.0; r.mT0[c] = (T0[c] + T0[3 + c] + T0[6 + c]) / 3.0;
        r.mS1[c] = (S1[c] + S1[3 + c] + S1[6 + c]) / 3.0; r.mT1[c] = (T1[c] + T1[3 + c] + T1[6 + c]) / 3.0;
    }
    r.valid = 1;
}

// Threads per block of joint_lm_kernel (template parameter).  The lanes are latency bound and independent, so the block
// shape does not change a solve; it decides how many SMs hold an LM block while the forward kernels of the next batches
// run beside it: a forward CTA needs ~31k registers and ~100 KB of shared memory, two fit an SM only when no LM block
// (216 registers per lane) is resident.  Fewer, larger blocks leave more SMs entirely to the forward kernels.
constexpr int LMT_DEFAULT = 256;
constexpr int LM_SOLVES_PER_LANE = 3;
constexpr int LM_PHASES = 3;
constexpr int LM_BUDGET[LM_PHASES] = {32, 160, 0x7fffffff};   // evaluations after which a solve moves to the next phase
constexpr int LM_PHASE_SHARE[LM_PHASES] = {1, 4, 32};         // phase p is sized for 1/share of the solves
// 4 blocks x 64 threads -> 255 registers per thread, no spills (6 blocks / 168 registers spilled 400 bytes per thread:
// measured 8.5 -> 6.9 ms for the joint stage alone; the pipelined throughput moves by 2% only, see DESIGN.md "overlap")
constexpr int LM_BLOCKS_PER_SM = 4;
constexpr int LM_SLOTS = 39;     // doubles per lane in shared memory: 36 point coordinates + joint direction
// statistics slots inside the 64-int counter block of the workspace (joint_tail): totals over all solves of a call
constexpr int LM_STAT_NJEV = 60, LM_STAT_NLM = 61;

// objective_eval over one lane's 3+3 points; element e of the lane lives at pts[e * LMT] (conflict-free)
template <int LMT>
struct LaneProb {
    const double *pts;
    __device__ __forceinline__ void load3(int e, double *v) const
    {
        v[0] = pts[e * LMT]; v[1] = pts[(e + 1) * LMT]; v[2] = pts[(e + 2) * LMT];
    }
    __device__ __forceinline__ double cost(const double *p) const
    {
        pm::RotVec r0, r1;
        r0.set(p);
        r1.set(p + 3);
        double fsq = 0.0, x[3], y[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { load3(3 * i, x); load3(9 + 3 * i, y); pm::accum_part(r0, 0, x, y, nullptr, fsq); }
#pragma unroll
        for (int i = 0; i < 3; ++i) { load3(18 + 3 * i, x); load3(27 + 3 * i, y); pm::accum_part(r1, 1, x, y, nullptr, fsq); }
        load3(36, x);
        pm::accum_joint(r0, r1, x, 3.0, nullptr, fsq);
        return fsq;
    }
    __device__ __forceinline__ void normal(const double *p, pm::Normal6 &N) const
    {
        pm::RotVec r0, r1;
        r0.set(p);
        r1.set(p + 3);
        N.zero();
        double fsq = 0.0, x[3], y[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) { load3(3 * i, x); load3(9 + 3 * i, y); pm::accum_part(r0, 0, x, y, &N, fsq); }
#pragma unroll
        for (int i = 0; i < 3; ++i) { load3(18 + 3 * i, x); load3(27 + 3 * i, y); pm::accum_part(r1, 1, x, y, &N, fsq); }
        load3(36, x);
        pm::accum_joint(r0, r1, x, 3.0, &N, fsq);
        N.fsq = fsq;
    }
};

// Phase `ph` of the solves.  Phase 0 claims the fresh records 0..total-1; a solve that has used LM_BUDGET[ph]
// evaluations without terminating is suspended at its next outer-iteration boundary (its lmder state goes back into the
// record -- 3 doubles + 3 counters, the Jacobian products are recomputed) and its index is appended to the work list of
// phase ph+1.  Later phases claim from the list the previous phase wrote.  Re-packing keeps the lanes of a warp busy:
// without it the few solves that run to MINPACK's maxfev = 600 each pin a warp at 1/32 utilisation (measured: 9.8 of 32
// lanes active on average, 8.3 ms per 256 clouds; the median solve needs 12 evaluations).
template <int LMT>
__global__ void __launch_bounds__(LMT, 256 / LMT) joint_lm_kernel(const JointArgs a, JointRec *recs, int total, int ph,
                                                                  int budget, const int *in_list, int *out_list)
{
    extern __shared__ double s_pts[];            // [LM_SLOTS][LMT]
    double *my = s_pts + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    int *claim = a.tail_count + 2 * ph;          // [2 ph] claim counter of this phase, [2 ph + 1] length of its work list
    const int avail = in_list ? a.tail_count[2 * ph + 1] : total;
    LaneProb<LMT> P;
    P.pts = my;
    pm::LmTick s;
    int t = -1;             // record this lane is solving
    int njev0 = 0;          // Jacobian evaluations the record had when this lane took it
    bool more = true;       // unclaimed work may remain
    for (;;) {
        const bool want = t < 0 && more;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, want);
        if (m) {                                              // warp-aggregated claim of the next items
            const int leader = __ffs(m) - 1;
            int base = 0;
            if ((int)lane == leader) base = atomicAdd(claim, __popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (want) {
                const int item = base + __popc(m & ((1u << lane) - 1u));
                if (item >= avail) {
                    more = false;
                } else {
                    const int mine = in_list ? in_list[item] : item;
                    const JointRec &r = recs[mine];
                    if (r.valid) {
                        t = mine;
                        // all 36 coordinates in flight at once (18 x 16-byte loads): a claim used to pay 9 dependent L2 / DRAM
                        // round trips here -- 12 % of the kernel's stall samples (profiles/r02_joint_lm_sass_stalls.txt)
                        double2 pv[18];
#pragma unroll
                        for (int e = 0; e < 18; ++e) pv[e] = __ldg(reinterpret_cast<const double2 *>(r.pts) + e);
#pragma unroll
                        for (int e = 0; e < 18; ++e) { my[(2 * e) * LMT] = pv[e].x; my[(2 * e + 1) * LMT] = pv[e].y; }
                        const double *u = a.axis_med + (size_t)(mine / a.niter) * 3;
                        my[36 * LMT] = u[0]; my[37 * LMT] = u[1]; my[38 * LMT] = u[2];
                        pm::lm_tick_init(s, r.x, r.fnorm);
                        if (r.iter > 1) {                     // resume a suspended solve
                            s.par = r.par; s.delta = r.delta; s.xnorm = r.xnorm;
                            s.iter = r.iter; s.nfev = r.nfev; s.njev = r.njev;
                        }
                        njev0 = s.njev;
                    }
                }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, t >= 0)) {
            if (!__any_sync(0xFFFFFFFFu, more)) break;
            continue;
        }
        if (t >= 0) {
            pm::lm_tick(P, s, 1e-4, 1e-8, 1e-8, 600, 100.0);  // ftol=1e-4 (:155), xtol=gtol=1e-8, max_nfev=100 n, factor=100
            if (s.info != 0) {
                JointRec &r = recs[t];
#pragma unroll
                for (int j = 0; j < 6; ++j) r.x[j] = s.x[j];
                a.nfev[t] = s.nfev;
                atomicAdd(a.tail_count + LM_STAT_NJEV, s.njev - njev0);
                atomicAdd(a.tail_count + LM_STAT_NLM, s.nlm);
                t = -1;
            } else if (s.need_jac && s.nfev >= budget) {      // out of budget: hand the solve to the next phase
                atomicAdd(a.tail_count + LM_STAT_NJEV, s.njev - njev0);
                atomicAdd(a.tail_count + LM_STAT_NLM, s.nlm);
                JointRec &r = recs[t];
#pragma unroll
                for (int j = 0; j < 6; ++j) r.x[j] = s.x[j];
                r.fnorm = s.fnorm; r.par = s.par; r.delta = s.delta; r.xnorm = s.xnorm;
                r.iter = s.iter; r.nfev = s.nfev; r.njev = s.njev;
                out_list[atomicAdd(a.tail_count + 2 * (ph + 1) + 1, 1)] = t;
                t = -1;
            }
        }
    }
}

// Launch shape of joint_lm_kernel: threads per block and the most blocks a phase uses (see the measurements at the launch).
static long lm_launch_shape(int sms, int *threads)
{
    static const int lm_cap = getenv("ANCSH_LM_BLOCKS_PER_SM") ? atoi(getenv("ANCSH_LM_BLOCKS_PER_SM")) : 1;
    static const int lmt_env = getenv("ANCSH_LM_THREADS") ? atoi(getenv("ANCSH_LM_THREADS")) : LMT_DEFAULT;
    static const int lane_pct = getenv("ANCSH_LM_LANE_PCT") ? atoi(getenv("ANCSH_LM_LANE_PCT")) : 50;
    const int lmt = lmt_env == 64 || lmt_env == 128 ? lmt_env : 256;
    long cap = (long)sms * 64 * (lm_cap >= 1 && lm_cap <= LM_BLOCKS_PER_SM ? lm_cap : 1) * (lane_pct >= 10 ? lane_pct : 100) / 100 / lmt;
    *threads = lmt;
    return cap < 1 ? 1 : cap;
}

__global__ void __launch_bounds__(JIT) joint_model_kernel(const JointArgs a, const JointRec *recs)
{
    const long t = (long)blockIdx.x * JIT + threadIdx.x;
    if (t >= (long)a.nprob * a.niter) return;
    const JointRec &r = recs[t];
    if (!r.valid) return;
    pm::JointModel m;
    pm::rotvec_to_matrix(r.x, m.R0);                           // :156-157
    pm::rotvec_to_matrix(r.x + 3, m.R1);
    m.s0 = r.s0; m.s1 = r.s1;
    double rs[3];
    pm::matvec3(m.R0, r.mS0, rs);
    for (int c = 0; c < 3; ++c) m.t0[c] = r.mT0[c] - m.s0 * rs[c];   // mean(T - s R S), :174-175 (un-refined scale)
    pm::matvec3(m.R1, r.mS1, rs);
    for (int c = 0; c < 3; ++c) m.t1[c] = r.mT1[c] - m.s1 * rs[c];
    a.models[t] = m;
}

// joint_transformation_verifier (:186-194): one block per problem, points staged in shared memory as f64, one
// warp per hypothesis at a time, lanes over points.
__global__ void __launch_bounds__(RT) joint_verify_kernel(const JointArgs a)
{
    extern __shared__ double s_pts[];
    const int prob = blockIdx.x;
    int pa, pb;
    joint_parts(prob, a.K, pa, pb);
    const int n0 = a.part_count[pa], n1 = a.part_count[pb];
    if (n0 <= 0 || n1 <= 0) {
        for (int h = threadIdx.x; h < a.niter; h += RT) a.scores[(size_t)prob * a.niter + h] = 0.0;
        return;
    }
    double *s_src0 = s_pts, *s_tgt0 = s_src0 + 3 * n0, *s_src1 = s_tgt0 + 3 * n0, *s_tgt1 = s_src1 + 3 * n1;
    load_part_f64(a.part_src + (size_t)pa * a.N * 3, a.part_tgt + (size_t)pa * a.N * 3, n0, s_src0, s_tgt0);
    load_part_f64(a.part_src + (size_t)pb * a.N * 3, a.part_tgt + (size_t)pb * a.N * 3, n1, s_src1, s_tgt1);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int h = warp; h < a.niter; h += RT / 32) {
        const pm::JointModel &m = a.models[(size_t)prob * a.niter + h];
        int c0 = 0, c1 = 0;
        for (int i = lane; i < n0; i += 32) c0 += is_inlier(m.R0, m.s0, m.t0, s_src0 + 3 * i, s_tgt0 + 3 * i, a.th2) ? 1 : 0;
        for (int i = lane; i < n1; i += 32) c1 += is_inlier(m.R1, m.s1, m.t1, s_src1 + 3 * i, s_tgt1 + 3 * i, a.th2) ? 1 : 0;
        c0 = __reduce_add_sync(0xFFFFFFFFu, c0);
        c1 = __reduce_add_sync(0xFFFFFFFFu, c1);
        // score = (sum(inl0)/res0.shape[0] + sum(inl1)/res1.shape[0]) / 2 with shape[0] == 3   (:193)
        if (lane == 0) a.scores[(size_t)prob * a.niter + h] = ((double)c0 / 3.0 + (double)c1 / 3.0) / 2.0;
    }
}

// block-cooperative evaluator of objective_eval over the masked (inlier) points in shared memory
// Only warp 0 runs the LM control code (~230 KB of SASS); the other warps of the block sit in serve() and take part
// in the residual / normal-equation sums when warp 0 posts a command.  Barrier protocol per evaluation: one
// __syncthreads publishing (cmd, x), then block_sum's two.
struct BlockProb {
    const double *x0, *y0, *x1, *y1;
    const int *l0, *l1;          // ascending indices of the inliers of each part
    int n0, n1;                  // list lengths
    double u[3];
    double nj;
    double *s_red;
    double *s_cross;   // [9] joint-row cross block of J^T J (only thread 0 contributes)
    int *s_cmd;        // 0 = exit, 1 = cost, 2 = normal equations
    double *s_x;       // [6] evaluation point

    __device__ void part_cost(const double *p) const
    {
        pm::RotVec r0, r1;
        r0.set(p);
        r1.set(p + 3);
        double fsq[1] = {0.0};
        for (int c = threadIdx.x; c < n0; c += RJ) pm::accum_part(r0, 0, x0 + 3 * l0[c], y0 + 3 * l0[c], nullptr, fsq[0]);
        for (int c = threadIdx.x; c < n1; c += RJ) pm::accum_part(r1, 1, x1 + 3 * l1[c], y1 + 3 * l1[c], nullptr, fsq[0]);
        if (threadIdx.x == 0) pm::accum_joint(r0, r1, u, nj, nullptr, fsq[0]);
        block_sum<1, RJ>(fsq, s_red);
        s_result = fsq[0];
    }
    // The point rows of part k only touch the diagonal block (k,k) of J^T J; the off-diagonal block comes from the joint
    // rows alone (thread 0).  So 6 + 6 + 6 + 1 = 19 values are reduced over the block instead of all 43.
    __device__ void part_normal(const double *p, pm::Normal6 &N) const
    {
        pm::RotVec r0, r1;
        r0.set(p);
        r1.set(p + 3);
        N.zero();
        double fsq = 0.0;
        for (int c = threadIdx.x; c < n0; c += RJ) pm::accum_part(r0, 0, x0 + 3 * l0[c], y0 + 3 * l0[c], &N, fsq);
        for (int c = threadIdx.x; c < n1; c += RJ) pm::accum_part(r1, 1, x1 + 3 * l1[c], y1 + 3 * l1[c], &N, fsq);
        if (threadIdx.x == 0) {
            pm::accum_joint(r0, r1, u, nj, &N, fsq);
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) s_cross[3 * a + b] = N.JtJ[a * 6 + 3 + b];
        }
        double v[19];
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = a; b < 3; ++b) v[6 * blk + 3 * a - (a * (a - 1)) / 2 + (b - a)] = N.JtJ[(3 * blk + a) * 6 + 3 * blk + b];
#pragma unroll
        for (int a = 0; a < 6; ++a) v[12 + a] = N.Jtf[a];
        v[18] = fsq;
        block_sum<19, RJ>(v, s_red);          // its barriers also publish s_cross
#pragma unroll
        for (int blk = 0; blk < 2; ++blk)
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = a; b < 3; ++b) {
                    const double x = v[6 * blk + 3 * a - (a * (a - 1)) / 2 + (b - a)];
                    N.JtJ[(3 * blk + a) * 6 + 3 * blk + b] = x;
                    N.JtJ[(3 * blk + b) * 6 + 3 * blk + a] = x;
                }
#pragma unroll
        for (int a = 0; a < 6; ++a) N.Jtf[a] = v[12 + a];
        N.fsq = v[18];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b) { N.JtJ[a * 6 + 3 + b] = s_cross[3 * a + b]; N.JtJ[(3 + b) * 6 + a] = s_cross[3 * a + b]; }
    }
    mutable double s_result;
    // ---- called by warp 0 only (through pm::lm_solve_tick) ----
    __host__ __device__ double cost(const double *p) const
    {
#ifdef __CUDA_ARCH__
        if (threadIdx.x == 0) { *s_cmd = 1; for (int j = 0; j < 6; ++j) s_x[j] = p[j]; }
        __syncthreads();
        part_cost(p);
        return s_result;
#else
        return 0.0;
#endif
    }
    __host__ __device__ void normal(const double *p, pm::Normal6 &N) const
    {
#ifdef __CUDA_ARCH__
        if (threadIdx.x == 0) { *s_cmd = 2; for (int j = 0; j < 6; ++j) s_x[j] = p[j]; }
        __syncthreads();
        part_normal(p, N);
#endif
    }
    // ---- warps 1.. ----
    __device__ void serve() const
    {
        for (;;) {
            __syncthreads();
            const int cmd = *s_cmd;
            if (cmd == 0) break;
            double p[6];
            for (int j = 0; j < 6; ++j) p[j] = s_x[j];
            if (cmd == 1) {
                part_cost(p);
            } else {
                pm::Normal6 N;
                part_normal(p, N);
            }
        }
    }
    __device__ void finish() const       // warp 0, after the solve
    {
        if (threadIdx.x == 0) *s_cmd = 0;
        __syncthreads();
    }
};
static_assert(sizeof(pm::Normal6) == 43 * sizeof(double), "Normal6 must be 43 contiguous doubles");

__device__ int block_first_argmax_f64(const double *scores, int n, double *s_val, int *s_idx)
{
    double bv = -1.0;
    int bi = 0x7FFFFFFF;
    for (int h = threadIdx.x; h < n; h += RJ) {
        const double v = scores[h];
        if (v > bv) { bv = v; bi = h; }       // ascending h per thread: strict '>' keeps the first
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xFFFFFFFFu, bv, off);
        const int oi = __shfl_down_sync(0xFFFFFFFFu, bi, off);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = bv; s_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    bv = -1.0; bi = 0x7FFFFFFF;
    for (int w = 0; w < RJ / 32; ++w)
        if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
    __syncthreads();
    return bi;
}

__global__ void __launch_bounds__(RJ, 2) joint_refit_kernel(const JointArgs a)
{
    extern __shared__ double s_pts[];
    __shared__ double s_red[(RJ / 32 + 1) * 19];
    __shared__ double s_val[RJ / 32];
    __shared__ int s_idx[RJ / 32];
    __shared__ pm::JointModel s_model;
    __shared__ int s_cmd;
    __shared__ double s_x[6];
    __shared__ double s_cross[9];
    const int prob = blockIdx.x, tid = threadIdx.x;
    int pa, pb;
    joint_parts(prob, a.K, pa, pb);
    const int n0 = a.part_count[pa], n1 = a.part_count[pb];
    const int jpart = pb - pa;
    unsigned char *g0 = a.inl0 + (size_t)prob * a.N, *g1 = a.inl1 + (size_t)prob * a.N;
    auto write_nan = [&](int flag) {
        if (tid == 0) {
            for (int i = 0; i < 9; ++i) { a.R0[(size_t)prob * 9 + i] = nan(""); a.R1[(size_t)prob * 9 + i] = nan(""); }
            for (int i = 0; i < 3; ++i) { a.t0[(size_t)prob * 3 + i] = nan(""); a.t1[(size_t)prob * 3 + i] = nan(""); }
            a.s0[prob] = nan(""); a.s1[prob] = nan("");
            a.status[pa + jpart] |= flag;
        }
    };
    if (n0 <= 0 || n1 <= 0) {
        for (int i = tid; i < a.N; i += RJ) { g0[i] = 0; g1[i] = 0; }
        if (tid == 0) { a.best[prob] = -1; a.score_out[prob] = 0.0; }
        write_nan(ANCSH_POSE_EMPTY_PART);
        return;
    }
    double *s_src0 = s_pts, *s_tgt0 = s_src0 + 3 * n0, *s_src1 = s_tgt0 + 3 * n0, *s_tgt1 = s_src1 + 3 * n1;
    int *s_l0 = reinterpret_cast<int *>(s_tgt1 + 3 * n1), *s_l1 = s_l0 + n0;           // inlier index lists
    unsigned char *s_m0 = reinterpret_cast<unsigned char *>(s_l1 + n1), *s_m1 = s_m0 + n0;
    load_part_f64(a.part_src + (size_t)pa * a.N * 3, a.part_tgt + (size_t)pa * a.N * 3, n0, s_src0, s_tgt0);
    load_part_f64(a.part_src + (size_t)pb * a.N * 3, a.part_tgt + (size_t)pb * a.N * 3, n1, s_src1, s_tgt1);
    const double *axis = a.axis_med + (size_t)prob * 3;
    const int hbest = block_first_argmax_f64(a.scores + (size_t)prob * a.niter, a.niter, s_val, s_idx);   // syncs
    if (tid == 0) {
        s_model = a.models[(size_t)prob * a.niter + hbest];       // the winning hypothesis' model (joint_model_kernel)
        a.best[prob] = hbest;
        a.score_out[prob] = a.scores[(size_t)prob * a.niter + hbest];
    }
    __syncthreads();
    // masks of the best hypothesis and the sums needed for centring: sum S, sum T, count per part
    double acc[14];
    for (int k = 0; k < 14; ++k) acc[k] = 0.0;
    for (int i = tid; i < a.N; i += RJ) {
        unsigned char in = 0;
        if (i < n0) {
            in = is_inlier(s_model.R0, s_model.s0, s_model.t0, s_src0 + 3 * i, s_tgt0 + 3 * i, a.th2) ? 1 : 0;
            s_m0[i] = in;
            if (in) { for (int c = 0; c < 3; ++c) { acc[c] += s_src0[3 * i + c]; acc[3 + c] += s_tgt0[3 * i + c]; } acc[6] += 1.0; }
        }
        g0[i] = in;
        in = 0;
        if (i < n1) {
            in = is_inlier(s_model.R1, s_model.s1, s_model.t1, s_src1 + 3 * i, s_tgt1 + 3 * i, a.th2) ? 1 : 0;
            s_m1[i] = in;
            if (in) { for (int c = 0; c < 3; ++c) { acc[7 + c] += s_src1[3 * i + c]; acc[10 + c] += s_tgt1[3 * i + c]; } acc[13] += 1.0; }
        }
        g1[i] = in;
    }
    block_sum<14, RJ>(acc, s_red);
    const int nin0 = (int)acc[6], nin1 = (int)acc[13];
    if (nin0 == 0 || nin1 == 0) { write_nan(ANCSH_POSE_NO_INLIERS); return; }
    double mS0[3], mT0[3], mS1[3], mT1[3];
    for (int c = 0; c < 3; ++c) {
        mS0[c] = acc[c] / nin0; mT0[c] = acc[3 + c] / nin0;
        mS1[c] = acc[7 + c] / nin1; mT1[c] = acc[10 + c] / nin1;
    }
    // scale_pts(S,T) and scale_pts(T,S) over the inliers of each part (:121-124): pairwise sums
    block_compact_indices<RJ>(s_m0, n0, s_l0, s_idx);
    block_compact_indices<RJ>(s_m1, n1, s_l1, s_idx);
    double pw[6];
    pair_sums<RJ>(s_src0, s_tgt0, s_l0, nin0, pw[0], pw[1], pw[2]);
    pair_sums<RJ>(s_src1, s_tgt1, s_l1, nin1, pw[3], pw[4], pw[5]);
    block_sum<6, RJ>(pw, s_red);
    const double sc0 = 2.0 * pw[0] / (2.0 * pw[1] + 1e-6), sinv0 = 2.0 * pw[0] / (2.0 * pw[2] + 1e-6);
    const double sc1 = 2.0 * pw[3] / (2.0 * pw[4] + 1e-6), sinv1 = 2.0 * pw[3] / (2.0 * pw[5] + 1e-6);
    // centre / pre-scale in place (:126-132): x = S - mean(S); y = sinv*T - mean(sinv*T)
    for (int i = tid; i < n0; i += RJ)
        for (int c = 0; c < 3; ++c) { s_src0[3 * i + c] -= mS0[c]; s_tgt0[3 * i + c] = sinv0 * s_tgt0[3 * i + c] - sinv0 * mT0[c]; }
    for (int i = tid; i < n1; i += RJ)
        for (int c = 0; c < 3; ++c) { s_src1[3 * i + c] -= mS1[c]; s_tgt1[3 * i + c] = sinv1 * s_tgt1[3 * i + c] - sinv1 * mT1[c]; }
    __syncthreads();
    // Kabsch initialisation (:138-139)
    double M[18];
    for (int k = 0; k < 18; ++k) M[k] = 0.0;
    for (int c = tid; c < nin0; c += RJ) {
        const int i = s_l0[c];
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) M[3 * p + q] += s_tgt0[3 * i + p] * s_src0[3 * i + q];
    }
    for (int c = tid; c < nin1; c += RJ) {
        const int i = s_l1[c];
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) M[9 + 3 * p + q] += s_tgt1[3 * i + p] * s_src1[3 * i + q];
    }
    block_sum<18, RJ>(M, s_red);
    double R0[9], R1[9], x[6];
    BlockProb P;
    P.x0 = s_src0; P.y0 = s_tgt0; P.x1 = s_src1; P.y1 = s_tgt1; P.l0 = s_l0; P.l1 = s_l1; P.n0 = nin0; P.n1 = nin1;
    P.u[0] = axis[0]; P.u[1] = axis[1]; P.u[2] = axis[2];
    P.nj = (double)min(nin0, nin1);
    P.s_red = s_red;
    P.s_cmd = &s_cmd;
    P.s_x = s_x;
    P.s_cross = s_cross;
    if (tid < 32) {
        pm::kabsch_rotation(M, R0);
        pm::kabsch_rotation(M + 9, R1);
        pm::matrix_to_rotvec(R0, x);
        pm::matrix_to_rotvec(R1, x + 3);
        pm::lm_solve_tick(P, x, 1e-4, 1e-8, 1e-8, 600, 100.0);   // warp 0: LM control; all its lanes see the same sums
        P.finish();
    } else {
        P.serve();
    }
    if (tid == 0) {
        pm::rotvec_to_matrix(x, R0);
        pm::rotvec_to_matrix(x + 3, R1);
        double rs[3];
        for (int i = 0; i < 9; ++i) { a.R0[(size_t)prob * 9 + i] = R0[i]; a.R1[(size_t)prob * 9 + i] = R1[i]; }
        a.s0[prob] = sc0; a.s1[prob] = sc1;
        pm::matvec3(R0, mS0, rs);
        for (int c = 0; c < 3; ++c) a.t0[(size_t)prob * 3 + c] = mT0[c] - sc0 * rs[c];     // :174
        pm::matvec3(R1, mS1, rs);
        for (int c = 0; c < 3; ++c) a.t1[(size_t)prob * 3 + c] = mT1[c] - sc1 * rs[c];     // :175
    }
}

// ================================================================================================
// Umeyama (lib/aligning.py:580-622)
// ================================================================================================
__global__ void __launch_bounds__(RT) umeyama_kernel(int nmax, const float *__restrict__ src, const float *__restrict__ tgt,
                                                     const int *__restrict__ cnt, double *scale, double *Rout, double *tout)
{
    __shared__ double s_red[(RT / 32) * 16];
    const int prob = blockIdx.x, tid = threadIdx.x;
    const int n = cnt[prob];
    const float *S = src + (size_t)prob * nmax * 3, *T = tgt + (size_t)prob * nmax * 3;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += RT)
        for (int c = 0; c < 3; ++c) { acc[c] += (double)S[3 * i + c]; acc[3 + c] += (double)T[3 * i + c]; }
    block_sum<6, RT>(acc, s_red);
    if (n <= 0) {
        if (tid == 0) {
            scale[prob] = nan("");
            for (int i = 0; i < 9; ++i) Rout[(size_t)prob * 9 + i] = nan("");
            for (int i = 0; i < 3; ++i) tout[(size_t)prob * 3 + i] = nan("");
        }
        return;
    }
    double ms[3], mt[3];
    for (int c = 0; c < 3; ++c) { ms[c] = acc[c] / n; mt[c] = acc[3 + c] / n; }
    double v[12];
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    for (int i = tid; i < n; i += RT) {
        double sc[3], tc[3];
        for (int c = 0; c < 3; ++c) { sc[c] = (double)S[3 * i + c] - ms[c]; tc[c] = (double)T[3 * i + c] - mt[c]; }
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) v[3 * p + q] += tc[p] * sc[q];        // CovMatrix = T_c S_c^T / n
        for (int c = 0; c < 3; ++c) v[9 + c] += sc[c] * sc[c];               // np.var(Source, axis=1) * n
    }
    block_sum<12, RT>(v, s_red);
    if (tid == 0) {
        double C[9], R[9], sv[3];
        for (int i = 0; i < 9; ++i) C[i] = v[i] / n;
        pm::kabsch_rotation(C, R);                 // U Vh with the reflection fix
        pm::singular_values3(C, sv);
        double dsum = sv[0] + sv[1] + sv[2];
        if (pm::det3(C) < 0.0) dsum -= 2.0 * sv[2];     // D[-1] = -D[-1] when det(U)det(Vh) < 0  (:599-601)
        const double varP = (v[9] + v[10] + v[11]) / n;
        const double sf = 1.0 / varP * dsum;
        scale[prob] = sf;
        // Rotation = (U Vh)^T (:606); Translation = T_mean - S_mean . (sf * Rotation)  (:613)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Rout[(size_t)prob * 9 + 3 * i + j] = R[3 * j + i];
        for (int j = 0; j < 3; ++j) {
            double d = 0.0;
            for (int i = 0; i < 3; ++i) d += ms[i] * sf * R[3 * j + i];      // (S_mean . Rot)[j] = sum_i S_mean[i] Rot[i][j]
            tout[(size_t)prob * 3 + j] = mt[j] - d;
        }
    }
}

// ================================================================================================
// estimateSimilarityTransform (lib/aligning.py:17-33): thresholds from the norm ratio (set_config, :88-103), 5-point
// Umeyama RANSAC with the reference's bookkeeping (getRANSACInliers :485-507, evaluateModel :540-547), Umeyama refit
// on the best hypothesis' inliers (:580-622).  SURVEY 8(a) row a-21.  One block per problem:
//   1. PassT / StopT: block reduction of the point norms;
//   2. thread h < niter: Umeyama of hypothesis h's 5 samples -> (sR | t) in shared memory;
//   3. warp per hypothesis: residuals of all points, sum of squares and the count of inliers with a NON-ZERO INDEX
//      (np.count_nonzero over the index array, :545 -- point 0 never counts);
//   4. thread 0 replays the sequential loop: strict `>` on the ratio keeps the first best, the loop stops after the
//      first iteration that leaves BestResidual < StopT;
//   5. mask of the best hypothesis (all points when no hypothesis ever scored), block-cooperative Umeyama on it.
// ================================================================================================
constexpr int SIM_MAX_ITER = 256;

struct SimArgs {
    int nmax, niter;
    const float *src, *tgt;
    const int *cnt, *idx;
    double *scale, *R, *t, *ratio;
    unsigned char *inliers;
    int *iters, *status;
};

__device__ void umeyama5(const float *S, const float *T, const int *id, double *M)
{
    double ms[3] = {0, 0, 0}, mt[3] = {0, 0, 0};
    for (int k = 0; k < 5; ++k)
        for (int c = 0; c < 3; ++c) { ms[c] += (double)S[3 * id[k] + c]; mt[c] += (double)T[3 * id[k] + c]; }
    for (int c = 0; c < 3; ++c) { ms[c] /= 5.0; mt[c] /= 5.0; }
    double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, var = 0.0;
    for (int k = 0; k < 5; ++k) {
        double sc[3], tc[3];
        for (int c = 0; c < 3; ++c) { sc[c] = (double)S[3 * id[k] + c] - ms[c]; tc[c] = (double)T[3 * id[k] + c] - mt[c]; }
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) C[3 * p + q] += tc[p] * sc[q];
        var += sc[0] * sc[0] + sc[1] * sc[1] + sc[2] * sc[2];
    }
    for (int i = 0; i < 9; ++i) C[i] /= 5.0;
    double R[9], sv[3];
    pm::kabsch_rotation(C, R);
    pm::singular_values3(C, sv);
    double dsum = sv[0] + sv[1] + sv[2];
    if (pm::det3(C) < 0.0) dsum -= 2.0 * sv[2];
    const double sf = 1.0 / (var / 5.0) * dsum;
    double rs[3];
    pm::matvec3(R, ms, rs);
    for (int i = 0; i < 9; ++i) M[i] = sf * R[i];                       // OutTransform[:3,:3] = diag(s) Rotation^T = s U Vh
    for (int c = 0; c < 3; ++c) M[9 + c] = mt[c] - sf * rs[c];           // Translation (:613)
}

__device__ __forceinline__ double sim_residual(const double *M, const float *S, const float *T, int i)
{
    const double x = (double)S[3 * i], y = (double)S[3 * i + 1], z = (double)S[3 * i + 2];
    const double d0 = (double)T[3 * i] - (M[0] * x + M[1] * y + M[2] * z + M[9]);
    const double d1 = (double)T[3 * i + 1] - (M[3] * x + M[4] * y + M[5] * z + M[10]);
    const double d2 = (double)T[3 * i + 2] - (M[6] * x + M[7] * y + M[8] * z + M[11]);
    return sqrt(d0 * d0 + d1 * d1 + d2 * d2);
}

__global__ void __launch_bounds__(RT) similarity_ransac_kernel(const SimArgs a)
{
    extern __shared__ __align__(16) unsigned char sim_smem[];
    float *s_src = reinterpret_cast<float *>(sim_smem);                  // n*3
    float *s_tgt = s_src + (size_t)a.nmax * 3;
    __shared__ double s_M[SIM_MAX_ITER * 12];
    __shared__ double s_res[SIM_MAX_ITER];
    __shared__ int s_nin[SIM_MAX_ITER];
    __shared__ double s_red[(RT / 32) * 16 + 16];
    __shared__ int s_best, s_stop;
    const int prob = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = a.cnt[prob];
    unsigned char *mask = a.inliers + (size_t)prob * a.nmax;
    for (int i = tid; i < a.nmax; i += RT) mask[i] = 0;
    if (n <= 0) {
        if (tid == 0) {
            a.scale[prob] = nan(""); a.ratio[prob] = 0.0; a.iters[prob] = 0; a.status[prob] = ANCSH_POSE_EMPTY_PART;
            for (int i = 0; i < 9; ++i) a.R[(size_t)prob * 9 + i] = nan("");
            for (int i = 0; i < 3; ++i) a.t[(size_t)prob * 3 + i] = nan("");
        }
        return;
    }
    const float *gs = a.src + (size_t)prob * a.nmax * 3, *gt = a.tgt + (size_t)prob * a.nmax * 3;
    double nrm[2] = {0.0, 0.0};
    for (int i = tid; i < n; i += RT) {
        double q = 0.0, w = 0.0;
        for (int c = 0; c < 3; ++c) {
            const float u = gs[3 * i + c], v = gt[3 * i + c];
            s_src[3 * i + c] = u; s_tgt[3 * i + c] = v;
            q += (double)u * (double)u; w += (double)v * (double)v;
        }
        nrm[0] += sqrt(q); nrm[1] += sqrt(w);
    }
    block_sum<2, RT>(nrm, s_red);                                         // ends with a barrier: s_src / s_tgt visible
    const double SourceNorm = nrm[0] / n, TargetNorm = nrm[1] / n;
    const double RatioTS = TargetNorm / SourceNorm, RatioST = SourceNorm / TargetNorm;
    const double PassT = RatioST > RatioTS ? RatioST : RatioTS;
    const double StopT = PassT / 100;
    for (int h = tid; h < a.niter; h += RT) umeyama5(s_src, s_tgt, a.idx + ((size_t)prob * a.niter + h) * 5, s_M + 12 * h);
    __syncthreads();
    for (int h = warp; h < a.niter; h += RT / 32) {
        double M[12];
        for (int k = 0; k < 12; ++k) M[k] = s_M[12 * h + k];
        double ss = 0.0;
        int cntin = 0;
        for (int i = lane; i < n; i += 32) {
            const double r = sim_residual(M, s_src, s_tgt, i);
            ss += r * r;
            cntin += (r < PassT && i != 0);
        }
        for (int off = 16; off > 0; off >>= 1) {
            ss += __shfl_xor_sync(0xFFFFFFFFu, ss, off);
            cntin += __shfl_xor_sync(0xFFFFFFFFu, cntin, off);
        }
        if (lane == 0) { s_res[h] = sqrt(ss); s_nin[h] = cntin; }
    }
    __syncthreads();
    if (tid == 0) {
        double best_res = 1e10;
        int best_nin = 0, best = -1, it = 0;
        for (int h = 0; h < a.niter; ++h) {
            ++it;
            if (s_nin[h] > best_nin) { best_nin = s_nin[h]; best_res = s_res[h]; best = h; }
            if (best_res < StopT) break;
        }
        s_best = best; s_stop = it;
        a.iters[prob] = it;
        a.ratio[prob] = (double)best_nin / (double)n;
    }
    __syncthreads();
    const int best = s_best;
    double M[12];
    if (best >= 0)
        for (int k = 0; k < 12; ++k) M[k] = s_M[12 * best + k];
    // Umeyama over the inliers of the best hypothesis (BestInlierIdx; arange(n) when nothing scored, :488)
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += RT) {
        const bool in = best < 0 || sim_residual(M, s_src, s_tgt, i) < PassT;
        mask[i] = in;
        if (in) {
            for (int c = 0; c < 3; ++c) { acc[c] += (double)s_src[3 * i + c]; acc[3 + c] += (double)s_tgt[3 * i + c]; }
            acc[6] += 1.0;
        }
    }
    block_sum<7, RT>(acc, s_red);
    const int nin = (int)acc[6];
    const double ratio = a.ratio[prob];
    if (nin == 0 || ratio < 0.1) {            // the reference returns four Nones below 0.1 (:25-27)
        if (tid == 0) {
            a.status[prob] = ANCSH_POSE_NO_INLIERS;
            a.scale[prob] = nan("");
            for (int i = 0; i < 9; ++i) a.R[(size_t)prob * 9 + i] = nan("");
            for (int i = 0; i < 3; ++i) a.t[(size_t)prob * 3 + i] = nan("");
        }
        return;
    }
    double ms[3], mt[3];
    for (int c = 0; c < 3; ++c) { ms[c] = acc[c] / nin; mt[c] = acc[3 + c] / nin; }
    double v[12];
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    for (int i = tid; i < n; i += RT) {
        if (!mask[i]) continue;
        double sc[3], tc[3];
        for (int c = 0; c < 3; ++c) { sc[c] = (double)s_src[3 * i + c] - ms[c]; tc[c] = (double)s_tgt[3 * i + c] - mt[c]; }
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) v[3 * p + q] += tc[p] * sc[q];
        for (int c = 0; c < 3; ++c) v[9 + c] += sc[c] * sc[c];
    }
    block_sum<12, RT>(v, s_red);
    if (tid == 0) {
        double C[9], R[9], sv[3];
        for (int i = 0; i < 9; ++i) C[i] = v[i] / nin;
        pm::kabsch_rotation(C, R);
        pm::singular_values3(C, sv);
        double dsum = sv[0] + sv[1] + sv[2];
        if (pm::det3(C) < 0.0) dsum -= 2.0 * sv[2];
        const double sf = 1.0 / ((v[9] + v[10] + v[11]) / nin) * dsum;
        a.scale[prob] = sf;
        a.status[prob] = 0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) a.R[(size_t)prob * 9 + 3 * i + j] = R[3 * j + i];      // Rotation = (U Vh)^T (:606)
        double rs[3];
        pm::matvec3(R, ms, rs);
        for (int j = 0; j < 3; ++j) a.t[(size_t)prob * 3 + j] = mt[j] - sf * rs[j];
    }
}

__global__ void sample_indices_kernel(unsigned long long seed, int stream_id, int nprob, int niter, const int *n_per,
                                      int *idx)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)nprob * niter) return;
    const int prob = (int)(t / niter), h = (int)(t - (long)prob * niter);
    int out[3] = {0, 0, 0};
    if (n_per[prob] > 0) pm::sample3(seed, (unsigned)prob, (unsigned)h, (unsigned)stream_id, n_per[prob], out);
    idx[t * 3 + 0] = out[0]; idx[t * 3 + 1] = out[1]; idx[t * 3 + 2] = out[2];
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" int ancsh_pose_plan(const ancsh_pose_cfg_t *cfg, int B, int N, ancsh_pose_ws_t *L)
{
    if (!cfg || !L || B <= 0 || N <= 0) return ANCSH_ERR_INVALID_ARG;
    const size_t K = cfg->n_parts, b = B, n = N;
    if (K < 1 || K > 8 || N > 4096 || cfg->niter_single < 1 || cfg->niter_joint < 1) return ANCSH_ERR_UNSUPPORTED;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align256(off + bytes); return o; };
    L->part_idx = take(b * K * n * 4);
    L->part_src = take(b * K * n * 3 * 4);
    L->part_tgt = take(b * K * n * 3 * 4);
    L->axis_med = take(b * (K > 1 ? K - 1 : 1) * 3 * 8);
    L->single_scores = take(b * K * (size_t)cfg->niter_single * 4);
    L->joint_scores = take(b * (K > 1 ? K - 1 : 1) * (size_t)cfg->niter_joint * 8);
    L->single_best = take(b * K * 4);
    L->joint_best = take(b * (K > 1 ? K - 1 : 1) * 4);
    L->joint_nfev = take(b * (K > 1 ? K - 1 : 1) * (size_t)cfg->niter_joint * 4);
    L->joint_models = take(b * (K > 1 ? K - 1 : 1) * (size_t)cfg->niter_joint * sizeof(pm::JointModel));
    L->joint_tail = take(256 + b * (K > 1 ? K - 1 : 1) * (size_t)cfg->niter_joint * (512 + 8));   /* counters + JointRec[] + 2 work lists */
    L->total_bytes = off;
    return ANCSH_OK;
}

extern "C" int ancsh_pose_solve(const ancsh_pose_cfg_t *cfg, const ancsh_pose_in_t *in, int B, int N, void *workspace,
                                size_t workspace_bytes, const ancsh_pose_out_t *out, void *const *stage_events,
                                void *stream)
{
    if (!cfg || !in || !out || !workspace) return ANCSH_ERR_INVALID_ARG;
    if (!in->P || !in->nocs || !in->mask) return ANCSH_ERR_INVALID_ARG;
    ancsh_pose_ws_t L;
    int rc = ancsh_pose_plan(cfg, B, N, &L);
    if (rc) return rc;
    if (workspace_bytes < L.total_bytes) return ANCSH_ERR_WORKSPACE;
    const int K = cfg->n_parts;
    if (K > 1 && (!in->joint_axis || !in->joint_cls)) return ANCSH_ERR_INVALID_ARG;
    if ((long)B * K > 65535) return ANCSH_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    int *part_idx = (int *)(ws + L.part_idx);
    float *part_src = (float *)(ws + L.part_src), *part_tgt = (float *)(ws + L.part_tgt);
    double *axis_med = (double *)(ws + L.axis_med);
    int *single_scores = (int *)(ws + L.single_scores);
    double *joint_scores = (double *)(ws + L.joint_scores);
    int *single_best = (int *)(ws + L.single_best), *joint_best = (int *)(ws + L.joint_best);

    int stage = 0;
#define STAGE_MARK()                                                                        \
    do {                                                                                    \
        if (stage_events) ANCSH_CUDA(cudaEventRecord((cudaEvent_t)stage_events[stage], st)); \
        ++stage;                                                                            \
    } while (0)
    STAGE_MARK();
    ANCSH_CUDA(cudaMemsetAsync(out->status, 0, (size_t)B * K * sizeof(int), st));
    {
        PartitionArgs a{in->P, in->nocs, in->mask, in->joint_axis, in->joint_cls, N, K, part_idx, part_src, part_tgt,
                        axis_med, out->part_count};
        int npow2 = 1;
        while (npow2 < N) npow2 <<= 1;
        size_t smem = ((N + 15) & ~15) + (size_t)RT * 8 * 4 + (size_t)3 * npow2 * 4;
        ANCSH_CUDA(cudaFuncSetAttribute(partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ANCSH_CUDA(cudaFuncSetAttribute(partition_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        partition_kernel<<<B, RT, smem, st>>>(a);
        ANCSH_CHECK_LAUNCH();
    }
    const double th2 = cfg->inlier_th * cfg->inlier_th;
    STAGE_MARK();
    {
        SingleArgs a{};
        a.part_src = part_src; a.part_tgt = part_tgt; a.part_count = out->part_count; a.idx = in->idx_single;
        a.scores = single_scores; a.best = single_best; a.N = N; a.niter = cfg->niter_single; a.th2 = th2;
        a.seed = cfg->seed;
        a.R = out->single_R; a.s = out->single_s; a.t = out->single_t; a.score_out = out->single_score;
        a.inliers = out->single_inliers; a.status = out->status;
        size_t smem = (size_t)N * 6 * sizeof(double);
        ANCSH_CUDA(cudaFuncSetAttribute(single_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ANCSH_CUDA(cudaFuncSetAttribute(single_score_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        dim3 grid(ancsh_cdiv(cfg->niter_single, RT), B * K);
        single_score_kernel<<<grid, RT, smem, st>>>(a);
        ANCSH_CHECK_LAUNCH();
        STAGE_MARK();
        size_t smem2 = smem + (size_t)5 * N;           // + inlier index list (int) + mask (byte)
        ANCSH_CUDA(cudaFuncSetAttribute(single_refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        ANCSH_CUDA(cudaFuncSetAttribute(single_refit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        single_refit_kernel<<<B * K, RSR, smem2, st>>>(a);
        ANCSH_CHECK_LAUNCH();
    }
    STAGE_MARK();
    if (K > 1) {
        JointArgs a{};
        a.part_src = part_src; a.part_tgt = part_tgt; a.part_count = out->part_count; a.axis_med = axis_med;
        a.idx0 = in->idx_joint0; a.idx1 = in->idx_joint1; a.scores = joint_scores; a.best = joint_best;
        a.nfev = (int *)(ws + L.joint_nfev);
        a.models = (pm::JointModel *)(ws + L.joint_models);
        a.nprob = B * (K - 1);
        a.tail_count = (int *)(ws + L.joint_tail);
        JointRec *recs = (JointRec *)(ws + L.joint_tail + 256);
        ANCSH_CUDA(cudaMemsetAsync(a.tail_count, 0, 64 * sizeof(int), st));
        a.N = N; a.K = K; a.niter = cfg->niter_joint; a.th2 = th2; a.seed = cfg->seed;
        a.R0 = out->joint_R0; a.s0 = out->joint_s0; a.t0 = out->joint_t0;
        a.R1 = out->joint_R1; a.s1 = out->joint_s1; a.t1 = out->joint_t1; a.score_out = out->joint_score;
        a.inl0 = out->joint_inliers0; a.inl1 = out->joint_inliers1; a.status = out->status;
        size_t smem = (size_t)N * 6 * sizeof(double);          // n0 + n1 <= N
        const long nsolves = (long)a.nprob * cfg->niter_joint;
        if (nsolves > 0x7fffffffL) return ANCSH_ERR_UNSUPPORTED;
        const unsigned gsolve = (unsigned)((nsolves + JIT - 1) / JIT);
        joint_init_kernel<<<gsolve, JIT, 0, st>>>(a, recs);
        ANCSH_CHECK_LAUNCH();
        {
            // persistent lanes: about one lane per LM_SOLVES_PER_LANE items, so that refilling finished lanes evens out
            // the uneven solve lengths; never more blocks than can be resident (SM count x LM_BLOCKS_PER_SM).  Later
            // phases see a fraction of the solves (LM_PHASE_SHARE: generous upper estimates; the claim loop copes with
            // any list length, surplus blocks exit at once).
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            // Lanes in flight and their packing.  The lanes are independent serial solves (latency bound: ~20% issue
            // utilisation), so the joint stage alone gets faster with more lanes -- but an SM that holds LM lanes (216
            // registers each) has no room for the second forward CTA of the batches that overlap this stage, and the
            // pipelined throughput is what counts.  Measured (B200, 256 clouds per batch, bench.py default):
            //   64 / 128 / 256 threads per block at 64 lanes per SM:   32.0k / 33.5k / 34.9k clouds/s
            //   256 threads, 135 / 100 / 75 / 50 / 35 / 25 % of those lanes:  34.6k / 34.8k / 35.4k / 35.9k / 35.8k / 35.1k
            //   (joint stage alone: 8.5 / 8.8 / 9.6 / 11.2 / 13.3 / 15.6 ms)
            //   512 threads at 128 registers (2 KB of spills per lane): 35.1k -- rejected; 384 threads at 168 registers: 37.4k vs
            //   36.9k in the same session (+1 %, joint stage 14.4 instead of 11.4 ms) -- not worth the spills
            // Default: 256-thread blocks, 32 lanes per SM -> 18 blocks on a 148-SM part.  ANCSH_LM_THREADS (64 / 128 / 256),
            // ANCSH_LM_LANE_PCT and ANCSH_LM_BLOCKS_PER_SM (multiplier 1..4) override -- more lanes for latency, fewer for
            // throughput.
            int lmt = 0;
            const long cap = lm_launch_shape(sms, &lmt);
            int *lists = (int *)(recs + nsolves);              // 2 x nsolves work-list entries behind the records
            // debugging aid: ANCSH_LM_TRACE=1 prints the duration of every LM phase (synchronises the stream)
            static const bool lm_trace = getenv("ANCSH_LM_TRACE") != nullptr;
            cudaEvent_t tev[LM_PHASES + 1];
            if (lm_trace) { for (auto &e : tev) cudaEventCreate(&e); cudaEventRecord(tev[0], st); }
            for (int ph = 0; ph < LM_PHASES; ++ph) {
                const long items = ph == 0 ? nsolves : nsolves / LM_PHASE_SHARE[ph] + 1;
                const long per_lane = ph == 0 ? LM_SOLVES_PER_LANE : 1;     // later phases: few, long solves -> one lane each
                long blocks = (items + (long)lmt * per_lane - 1) / ((long)lmt * per_lane);
                blocks = blocks < 1 ? 1 : (blocks > cap ? cap : blocks);
                const int *in_list = ph == 0 ? nullptr : lists + (size_t)((ph - 1) & 1) * nsolves;
                int *out_list = lists + (size_t)(ph & 1) * nsolves;
                const size_t lm_smem = (size_t)LM_SLOTS * lmt * sizeof(double);
                if (lmt == 64) {
                    joint_lm_kernel<64><<<(unsigned)blocks, 64, lm_smem, st>>>(a, recs, (int)nsolves, ph, LM_BUDGET[ph], in_list, out_list);
                } else if (lmt == 128) {
                    joint_lm_kernel<128><<<(unsigned)blocks, 128, lm_smem, st>>>(a, recs, (int)nsolves, ph, LM_BUDGET[ph], in_list, out_list);
                } else {
                    ANCSH_CUDA(cudaFuncSetAttribute(joint_lm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lm_smem));
                    joint_lm_kernel<256><<<(unsigned)blocks, 256, lm_smem, st>>>(a, recs, (int)nsolves, ph, LM_BUDGET[ph], in_list, out_list);
                }
                ANCSH_CHECK_LAUNCH();
                if (lm_trace) cudaEventRecord(tev[ph + 1], st);
            }
            if (lm_trace) {
                cudaStreamSynchronize(st);
                int h_tail[64];
                cudaMemcpy(h_tail, a.tail_count, sizeof h_tail, cudaMemcpyDeviceToHost);
                fprintf(stderr, "[lm trace] solves %ld:", nsolves);
                for (int ph = 0; ph < LM_PHASES; ++ph) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, tev[ph], tev[ph + 1]);
                    fprintf(stderr, "  phase %d: %d items %.3f ms", ph, ph == 0 ? (int)nsolves : h_tail[2 * ph + 1], ms);
                }
                fprintf(stderr, "  njev %d lmpar iterations %d\n", h_tail[LM_STAT_NJEV], h_tail[LM_STAT_NLM]);
                for (auto &e : tev) cudaEventDestroy(e);
            }
        }
        joint_model_kernel<<<gsolve, JIT, 0, st>>>(a, recs);
        ANCSH_CHECK_LAUNCH();
        ANCSH_CUDA(cudaFuncSetAttribute(joint_verify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ANCSH_CUDA(cudaFuncSetAttribute(joint_verify_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        joint_verify_kernel<<<a.nprob, RT, smem, st>>>(a);
        ANCSH_CHECK_LAUNCH();
        STAGE_MARK();
        size_t smem2 = smem + (size_t)5 * N + 16;      // + inlier index lists (int) + masks (byte)
        ANCSH_CUDA(cudaFuncSetAttribute(joint_refit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        ANCSH_CUDA(cudaFuncSetAttribute(joint_refit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        joint_refit_kernel<<<B * (K - 1), RJ, smem2, st>>>(a);
        ANCSH_CHECK_LAUNCH();
    } else {
        STAGE_MARK();
    }
    STAGE_MARK();
#undef STAGE_MARK
    return ANCSH_OK;
}

// ---- ransac() on caller datasets (parallel_ancsh_pose.py:20) for hosts without the Python mirror ---------------------------
// Both entries express the dataset as a one-cloud problem of ancsh_pose_solve: the dataset's points become the parts of a
// cloud (part membership through a 0/1 mask), so the kernels, the sampling streams and the results are the ones of the
// per-cloud path.
namespace {
__global__ void ransac_dataset_kernel(int n0, int n1, const float *__restrict__ s0, const float *__restrict__ t0,
                                      const float *__restrict__ s1, const float *__restrict__ t1, double dx, double dy, double dz,
                                      float *P, float *nocs, float *mask, float *axis, int *jcls)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, n = n0 + n1, K = n1 > 0 ? 2 : 1;
    if (i >= n) return;
    const bool first = i < n0;
    const float *s = first ? s0 + 3 * i : s1 + 3 * (i - n0), *t = first ? t0 + 3 * i : t1 + 3 * (i - n0);
    for (int c = 0; c < 3; ++c) P[3 * i + c] = t[c];
    for (int c = 0; c < 3 * K; ++c) nocs[3 * K * i + c] = 0.f;
    for (int c = 0; c < 3; ++c) nocs[3 * K * i + (first ? 0 : 3) + c] = s[c];
    for (int k = 0; k < K; ++k) mask[K * i + k] = (k == (first ? 0 : 1)) ? 1.f : 0.f;
    if (K == 2) { axis[3 * i] = (float)dx; axis[3 * i + 1] = (float)dy; axis[3 * i + 2] = (float)dz; jcls[i] = 1; }
}

struct DatasetLayout { size_t P, nocs, mask, axis, jcls, out, pose_ws, total; };

int dataset_layout(int n, int K, int niter_single, int niter_joint, DatasetLayout *L, ancsh_pose_ws_t *pw)
{
    ancsh_pose_cfg_t cfg{K, niter_single, niter_joint, 0.1, 0ull};
    int rc = ancsh_pose_plan(&cfg, 1, n, pw);
    if (rc) return rc;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align256(off + bytes); return o; };
    L->P = take((size_t)n * 3 * 4); L->nocs = take((size_t)n * 3 * K * 4); L->mask = take((size_t)n * K * 4);
    L->axis = take((size_t)n * 3 * 4); L->jcls = take((size_t)n * 4);
    L->out = take(4096 + (size_t)4 * n);                 // scratch: outputs the caller did not ask for, N-strided inlier rows
    L->pose_ws = take(pw->total_bytes);
    L->total = off;
    return ANCSH_OK;
}
}  // namespace

extern "C" int ancsh_ransac_workspace_bytes(int n_total, int n_parts, int niter, size_t *bytes)
{
    if (!bytes || n_total < 1 || (n_parts != 1 && n_parts != 2) || niter < 1) return ANCSH_ERR_INVALID_ARG;
    DatasetLayout L;
    ancsh_pose_ws_t pw;
    int rc = dataset_layout(n_total, n_parts, n_parts == 1 ? niter : 1, n_parts == 2 ? niter : 1, &L, &pw);
    if (rc) return rc;
    *bytes = L.total;
    return ANCSH_OK;
}

extern "C" int ancsh_ransac_single(int n, const float *source, const float *target, double inlier_th, int niter, const int *sample_idx,
                                   unsigned long long seed, void *workspace, size_t workspace_bytes, double *R, double *scale,
                                   double *t, int *score, unsigned char *inliers, int *status, void *stream)
{
    if (n < 1 || !source || !target || !workspace || !R || !scale || !t || !inliers || !status) return ANCSH_ERR_INVALID_ARG;
    DatasetLayout L;
    ancsh_pose_ws_t pw;
    int rc = dataset_layout(n, 1, niter, 1, &L, &pw);
    if (rc) return rc;
    if (workspace_bytes < L.total) return ANCSH_ERR_WORKSPACE;
    char *ws = (char *)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    ransac_dataset_kernel<<<ancsh_cdiv(n, 256), 256, 0, st>>>(n, 0, source, target, nullptr, nullptr, 0, 0, 0, (float *)(ws + L.P),
                                                              (float *)(ws + L.nocs), (float *)(ws + L.mask), nullptr, nullptr);
    ANCSH_CHECK_LAUNCH();
    ancsh_pose_cfg_t cfg{1, niter, 1, inlier_th, seed};
    ancsh_pose_in_t in{};
    in.P = (float *)(ws + L.P); in.nocs = (float *)(ws + L.nocs); in.mask = (float *)(ws + L.mask); in.idx_single = sample_idx;
    ancsh_pose_out_t out{};
    char *o = ws + L.out;
    out.single_R = R; out.single_s = scale; out.single_t = t; out.single_inliers = inliers; out.status = status;
    out.single_score = score ? score : (int *)(o + 0);
    out.part_count = (int *)(o + 64);
    // K == 1: the joint outputs are never written (ancsh_pose_solve skips the joint stage)
    return ancsh_pose_solve(&cfg, &in, 1, n, ws + L.pose_ws, pw.total_bytes, &out, nullptr, stream);
}

extern "C" int ancsh_ransac_joint(int n0, const float *source0, const float *target0, int n1, const float *source1,
                                  const float *target1, const double *joint_direction_host, double inlier_th, int niter,
                                  const int *sample_idx0, const int *sample_idx1, unsigned long long seed, void *workspace,
                                  size_t workspace_bytes, double *R0, double *s0, double *t0, double *R1, double *s1, double *t1,
                                  double *score, unsigned char *inliers0, unsigned char *inliers1, int *status, void *stream)
{
    if (n0 < 1 || n1 < 1 || !source0 || !target0 || !source1 || !target1 || !joint_direction_host || !workspace || !R0 || !s0 ||
        !t0 || !R1 || !s1 || !t1 || !inliers0 || !inliers1 || !status)
        return ANCSH_ERR_INVALID_ARG;
    const int n = n0 + n1;
    DatasetLayout L;
    ancsh_pose_ws_t pw;
    int rc = dataset_layout(n, 2, 1, niter, &L, &pw);
    if (rc) return rc;
    if (workspace_bytes < L.total) return ANCSH_ERR_WORKSPACE;
    char *ws = (char *)workspace;
    cudaStream_t st = (cudaStream_t)stream;
    ransac_dataset_kernel<<<ancsh_cdiv(n, 256), 256, 0, st>>>(n0, n1, source0, target0, source1, target1, joint_direction_host[0],
                                                              joint_direction_host[1], joint_direction_host[2], (float *)(ws + L.P),
                                                              (float *)(ws + L.nocs), (float *)(ws + L.mask), (float *)(ws + L.axis),
                                                              (int *)(ws + L.jcls));
    ANCSH_CHECK_LAUNCH();
    ancsh_pose_cfg_t cfg{2, 1, niter, inlier_th, seed};
    ancsh_pose_in_t in{};
    in.P = (float *)(ws + L.P); in.nocs = (float *)(ws + L.nocs); in.mask = (float *)(ws + L.mask);
    in.joint_axis = (float *)(ws + L.axis); in.joint_cls = (int *)(ws + L.jcls);
    in.idx_joint0 = sample_idx0; in.idx_joint1 = sample_idx1;
    ancsh_pose_out_t out{};
    char *o = ws + L.out;
    // single-part outputs of the two parts (one hypothesis each) go to scratch
    out.single_R = (double *)(o + 256); out.single_s = (double *)(o + 512); out.single_t = (double *)(o + 640);
    out.single_score = (int *)(o + 768); out.single_inliers = (unsigned char *)(o + 4096);
    out.part_count = (int *)(o + 64);
    out.joint_R0 = R0; out.joint_s0 = s0; out.joint_t0 = t0; out.joint_R1 = R1; out.joint_s1 = s1; out.joint_t1 = t1;
    out.joint_score = score ? score : (double *)(o + 896);
    // the solver writes inlier rows of the cloud's length N = n0 + n1: scratch rows, then the leading n0 / n1 entries
    unsigned char *row0 = (unsigned char *)(o + 4096 + (size_t)2 * n), *row1 = row0 + n;
    out.joint_inliers0 = row0; out.joint_inliers1 = row1; out.status = status;
    rc = ancsh_pose_solve(&cfg, &in, 1, n, ws + L.pose_ws, pw.total_bytes, &out, nullptr, stream);
    if (rc) return rc;
    ANCSH_CUDA(cudaMemcpyAsync(inliers0, row0, (size_t)n0, cudaMemcpyDeviceToDevice, st));
    ANCSH_CUDA(cudaMemcpyAsync(inliers1, row1, (size_t)n1, cudaMemcpyDeviceToDevice, st));
    return ANCSH_OK;
}

extern "C" int ancsh_pose_sample_indices(unsigned long long seed, int stream_id, int nprob, int niter,
                                         const int *n_per_problem, int *idx, void *stream)
{
    if (nprob < 0 || niter < 0 || !n_per_problem || !idx) return ANCSH_ERR_INVALID_ARG;
    const long total = (long)nprob * niter;
    if (total == 0) return ANCSH_OK;
    sample_indices_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, stream_id, nprob, niter,
                                                                                             n_per_problem, idx);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

extern "C" int ancsh_umeyama(int nprob, int nmax, const float *src, const float *tgt, const int *cnt, double *scale,
                             double *R, double *t, void *stream)
{
    if (nprob < 0 || nmax <= 0 || !src || !tgt || !cnt || !scale || !R || !t) return ANCSH_ERR_INVALID_ARG;
    if (nprob == 0) return ANCSH_OK;
    umeyama_kernel<<<nprob, RT, 0, (cudaStream_t)stream>>>(nmax, src, tgt, cnt, scale, R, t);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

extern "C" int ancsh_similarity_ransac(int nprob, int nmax, int niter, const float *src, const float *tgt, const int *cnt,
                                       const int *idx, double *scale, double *R, double *t, double *inlier_ratio,
                                       unsigned char *inliers, int *iters_run, int *status, void *stream)
{
    if (nprob < 0 || nmax <= 0 || nmax > 8192 || niter <= 0 || niter > SIM_MAX_ITER) return ANCSH_ERR_INVALID_ARG;
    if (nprob == 0) return ANCSH_OK;
    if (!src || !tgt || !cnt || !idx || !scale || !R || !t || !inlier_ratio || !inliers || !iters_run || !status)
        return ANCSH_ERR_INVALID_ARG;
    SimArgs a{nmax, niter, src, tgt, cnt, idx, scale, R, t, inlier_ratio, inliers, iters_run, status};
    const size_t smem = (size_t)nmax * 6 * sizeof(float);
    ANCSH_CUDA(cudaFuncSetAttribute(similarity_ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    similarity_ransac_kernel<<<nprob, RT, smem, (cudaStream_t)stream>>>(a);
    ANCSH_CHECK_LAUNCH();
    return ANCSH_OK;
}

// Launch shape of the joint LM solve on the current device (diagnostics: the stage is deliberately narrow, bench.py reports
// the SMs it occupies next to its roofline fraction).
extern "C" int ancsh_pose_lm_shape(int *threads_per_block, int *max_blocks)
{
    if (!threads_per_block || !max_blocks) return ANCSH_ERR_INVALID_ARG;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *max_blocks = (int)lm_launch_shape(sms, threads_per_block);
    return ANCSH_OK;
}
