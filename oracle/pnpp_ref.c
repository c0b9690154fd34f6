/*
 * oracle/pnpp_ref.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Plain-C restatement of the reference's PointNet++ native ops and layer arithmetic for the
 * ANCSH hot path.  Each function cites the reference file:line it follows (paths relative to
 * /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from this file.
 *
 * Parity status: the FPS / ball-query / group restatements are pinned against the reference's
 * own CUDA kernels compiled for sm_100a (oracle/_ref/libref_tfops.so, see oracle/build.py and
 * tests/test_ops_gpu.py) and against goldens those kernels produced on a B200
 * (tests/golden/ops_ref_b200.npz).  three_nn / three_interpolate are pinned against the
 * reference's CPU loops extracted at build time (oracle/_ref/libref_interp.so).  The conv/BN
 * arithmetic restates TF semantics (TF1 cannot be installed): "parity unpinned" for that part.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/build.py).  -ffp-contract=off is
 * REQUIRED: fused multiply-adds are written explicitly with fmaf() where the reference's GPU
 * SASS fuses (SURVEY.md section 8a "arithmetic contract"), and must not appear anywhere else
 * (the reference's CPU three_nn is un-contracted f32).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ---------------------------------------------------------------------------------------------
 * farthest point sampling -- pointnet_plusplus/utils/tf_ops/sampling/tf_sampling_g.cu:105-170
 *
 * The kernel runs 512 threads; thread t scans k = t, t+512, ... keeping the FIRST strict
 * maximum (:146-149), then a shared-memory tree keeps the LEFT operand on ties (:158-161).
 * Net tie rule: among equal maxima the winner has the smallest (k mod 512, k div 512).
 * Squared distance is contracted by nvcc to fma(dz,dz, fma(dy,dy, dx*dx)).
 * ------------------------------------------------------------------------------------------- */
#define ORC_FPS_BLOCK 512

void orc_fps(int b, int n, int m, const float *xyz, int *idx)
{
    if (m <= 0) return;
    float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    float best_t[ORC_FPS_BLOCK];
    int besti_t[ORC_FPS_BLOCK];
    for (int i = 0; i < b; ++i) {
        const float *p = xyz + (size_t)i * n * 3;
        int old = 0;
        idx[(size_t)i * m + 0] = old;
        for (int j = 0; j < n; ++j) temp[j] = 1e38f;              /* :116-118 */
        for (int j = 1; j < m; ++j) {
            float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int t = 0; t < ORC_FPS_BLOCK; ++t) {             /* per-thread scan :131-150 */
                float best = -1.0f;
                int besti = 0;
                for (int k = t; k < n; k += ORC_FPS_BLOCK) {
                    float td = temp[k];
                    float dx = p[k * 3 + 0] - x1;
                    float dy = p[k * 3 + 1] - y1;
                    float dz = p[k * 3 + 2] - z1;
                    float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    float d2 = d < td ? d : td;                   /* min(d,td) */
                    if (d2 != td) temp[k] = d2;
                    if (d2 > best) { best = d2; besti = k; }
                }
                best_t[t] = best;
                besti_t[t] = besti;
            }
            /* tree reduce :153-163, left operand kept on ties */
            for (int u = 0; (1 << u) < ORC_FPS_BLOCK; ++u) {
                for (int t = 0; t < (ORC_FPS_BLOCK >> (u + 1)); ++t) {
                    int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
                    if (best_t[i1] < best_t[i2]) { best_t[i1] = best_t[i2]; besti_t[i1] = besti_t[i2]; }
                }
            }
            old = besti_t[0];
            idx[(size_t)i * m + j] = old;
        }
    }
    free(temp);
}

/* gather_point -- tf_sampling_g.cu:172-181 */
void orc_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out)
{
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)i * m + j];
            for (int c = 0; c < 3; ++c) out[((size_t)i * m + j) * 3 + c] = inp[((size_t)i * n + a) * 3 + c];
        }
}

/* ---------------------------------------------------------------------------------------------
 * query_ball_point -- pointnet_plusplus/utils/tf_ops/grouping/tf_grouping_g.cu:3-36
 * first-nsample-in-index-order, d = max(sqrtf(.),1e-20f) < radius, first hit pre-fills all slots.
 * The reference leaves idx rows with no hit uninitialised; we define them as 0.
 * ------------------------------------------------------------------------------------------- */
void orc_ball_query(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                    int *idx, int *pts_cnt)
{
    for (int bi = 0; bi < b; ++bi) {
        const float *p1 = xyz1 + (size_t)bi * n * 3;
        const float *p2 = xyz2 + (size_t)bi * m * 3;
        int *id = idx + (size_t)bi * m * nsample;
        int *pc = pts_cnt + (size_t)bi * m;
        for (int j = 0; j < m; ++j) {
            int cnt = 0;
            for (int l = 0; l < nsample; ++l) id[j * nsample + l] = 0;
            float x2 = p2[j * 3 + 0], y2 = p2[j * 3 + 1], z2 = p2[j * 3 + 2];
            for (int k = 0; k < n; ++k) {
                if (cnt == nsample) break;
                float dx = x2 - p1[k * 3 + 0];
                float dy = y2 - p1[k * 3 + 1];
                float dz = z2 - p1[k * 3 + 2];
                float d = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                d = d > 1e-20f ? d : 1e-20f;
                if (d < radius) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) id[j * nsample + l] = k;
                    id[j * nsample + cnt] = k;
                    cnt += 1;
                }
            }
            pc[j] = cnt;
        }
    }
}

/* group_point -- tf_grouping_g.cu:40-57 */
void orc_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out)
{
    for (int bi = 0; bi < b; ++bi) {
        const float *p = points + (size_t)bi * n * c;
        const int *id = idx + (size_t)bi * m * nsample;
        float *o = out + (size_t)bi * m * nsample * c;
        for (int j = 0; j < m; ++j)
            for (int k = 0; k < nsample; ++k) {
                int ii = id[j * nsample + k];
                memcpy(o + ((size_t)j * nsample + k) * c, p + (size_t)ii * c, sizeof(float) * c);
            }
    }
}

/* ---------------------------------------------------------------------------------------------
 * three_nn -- pointnet_plusplus/utils/tf_ops/3d_interpolation/tf_interpolate.cpp:60-103
 * f32 un-contracted squared distance promoted to double, strict '<' chain.
 * ------------------------------------------------------------------------------------------- */
void orc_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx)
{
    for (int i = 0; i < b; ++i) {
        for (int j = 0; j < n; ++j) {
            float x1 = xyz1[j * 3 + 0], y1 = xyz1[j * 3 + 1], z1 = xyz1[j * 3 + 2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                float x2 = xyz2[k * 3 + 0], y2 = xyz2[k * 3 + 1], z2 = xyz2[k * 3 + 2];
                float df = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
                double d = df;
                if (d < best1) {
                    best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2; best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            dist[j * 3 + 0] = (float)best1; idx[j * 3 + 0] = besti1;
            dist[j * 3 + 1] = (float)best2; idx[j * 3 + 1] = besti2;
            dist[j * 3 + 2] = (float)best3; idx[j * 3 + 2] = besti3;
        }
        xyz1 += n * 3; xyz2 += m * 3; dist += n * 3; idx += n * 3;
    }
}

/* three_interpolate -- tf_interpolate.cpp:107-127 (f32, left-to-right, no fma) */
void orc_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight,
                           float *out)
{
    for (int i = 0; i < b; ++i) {
        for (int j = 0; j < n; ++j) {
            float w1 = weight[j * 3], w2 = weight[j * 3 + 1], w3 = weight[j * 3 + 2];
            int i1 = idx[j * 3], i2 = idx[j * 3 + 1], i3 = idx[j * 3 + 2];
            for (int l = 0; l < c; ++l)
                out[(size_t)j * c + l] =
                    points[(size_t)i1 * c + l] * w1 + points[(size_t)i2 * c + l] * w2 + points[(size_t)i3 * c + l] * w3;
        }
        points += (size_t)m * c; idx += n * 3; weight += n * 3; out += (size_t)n * c;
    }
}

/* inverse-distance weights -- pointnet_plusplus/utils/pointnet_util.py:219-222
 * dist=max(dist,1e-10); norm=sum(1/dist); weight=(1/dist)/norm   (all f32) */
void orc_three_weights(int rows, const float *dist, float *weight)
{
    for (int r = 0; r < rows; ++r) {
        float d0 = dist[r * 3 + 0], d1 = dist[r * 3 + 1], d2 = dist[r * 3 + 2];
        d0 = d0 > 1e-10f ? d0 : 1e-10f;
        d1 = d1 > 1e-10f ? d1 : 1e-10f;
        d2 = d2 > 1e-10f ? d2 : 1e-10f;
        float r0 = 1.0f / d0, r1 = 1.0f / d1, r2 = 1.0f / d2;
        float norm = (r0 + r1) + r2;
        weight[r * 3 + 0] = r0 / norm;
        weight[r * 3 + 1] = r1 / norm;
        weight[r * 3 + 2] = r2 / norm;
    }
}

/* ---------------------------------------------------------------------------------------------
 * 1x1 convolution + bias (+ inference batch-norm, eps 1e-3) (+ ReLU)
 *   pointnet_plusplus/utils/tf_util.py:120-185 (conv2d), :52-115 (conv1d), :527-531 (BN)
 * x: (rows,cin) row-major, W: (cin,cout) (TF [1,1,cin,cout] / [1,cin,cout]), f32 accumulation in
 * natural k order.  bn == NULL -> no batch norm; else bn = {gamma,beta,mean,var} each (cout).
 * ------------------------------------------------------------------------------------------- */
void orc_conv1x1(long rows, int cin, int cout, const float *x, const float *W, const float *bias,
                 const float *gamma, const float *beta, const float *mean, const float *var, int relu, float *y)
{
    float *scale = NULL;
    if (gamma) {
        scale = (float *)malloc(sizeof(float) * cout);
        for (int o = 0; o < cout; ++o) scale[o] = gamma[o] / sqrtf(var[o] + 1e-3f);
    }
#pragma omp parallel for schedule(static)
    for (long r = 0; r < rows; ++r) {
        float *yr = y + (size_t)r * cout;
        const float *xr = x + (size_t)r * cin;
        for (int o = 0; o < cout; ++o) yr[o] = 0.0f;
        for (int k = 0; k < cin; ++k) {
            float xv = xr[k];
            const float *wk = W + (size_t)k * cout;
            for (int o = 0; o < cout; ++o) yr[o] = yr[o] + xv * wk[o];
        }
        for (int o = 0; o < cout; ++o) {
            float v = yr[o] + bias[o];
            if (gamma) v = (v - mean[o]) * scale[o] + beta[o];
            if (relu) v = v > 0.0f ? v : 0.0f;
            yr[o] = v;
        }
    }
    free(scale);
}

/* max over the nsample axis -- pointnet_util.py:134 ; x: (groups, nsample, c) -> (groups, c) */
void orc_group_max(long groups, int nsample, int c, const float *x, float *y)
{
#pragma omp parallel for schedule(static)
    for (long g = 0; g < groups; ++g) {
        const float *xg = x + (size_t)g * nsample * c;
        float *yg = y + (size_t)g * c;
        for (int o = 0; o < c; ++o) yg[o] = xg[o];
        for (int s = 1; s < nsample; ++s)
            for (int o = 0; o < c; ++o) {
                float v = xg[(size_t)s * c + o];
                if (v > yg[o]) yg[o] = v;
            }
    }
}

/* thread control for the cpu_baseline leg of bench.py (OpenMP over rows; 1 = scalar port) */
#ifdef _OPENMP
#include <omp.h>
void orc_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
int orc_get_max_threads(void) { return omp_get_max_threads(); }
#else
void orc_set_threads(int n) { (void)n; }
int orc_get_max_threads(void) { return 1; }
#endif
