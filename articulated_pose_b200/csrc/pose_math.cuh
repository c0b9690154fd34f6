// pose_math.cuh -- f64 geometry + Levenberg-Marquardt for the ANCSH pose stage.  __host__ __device__ so that
// the very same code is unit-tested on the CPU (tests/hostsim) against numpy / scipy and runs inside the
// RANSAC kernels (pose.cu).
//
//   kabsch_rotation     lib/d3_utils.py:206-220  rotate_pts     (numpy.linalg.svd -> one-sided Jacobi)
//   pair_scale          lib/d3_utils.py:237-246  scale_pts      (all ordered pairs, +1e-6)
//   rodrigues / d_rodrigues  lib/d3_utils.py:150-163 rotate_points_with_rotvec and its analytic derivative
//   matrix_to_rotvec / rotvec_to_matrix   scipy Rotation.from_matrix().as_rotvec() / from_rotvec().as_matrix()
//   lm_solve            scipy.optimize.least_squares(method='lm', x_scale=1.0) = MINPACK lmder, mode 2,
//                       diag = 1, factor = 100 (evaluation/parallel_ancsh_pose.py:154-155), restated on the
//                       6x6 normal equations: R from a pivoted Cholesky of J^T J equals qrfac's R up to row
//                       signs, (Q^T f)[:n] = R^-T P^T J^T f, so lmpar / qrsolv run unchanged.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define PM_HD __host__ __device__ __forceinline__
#define PM_HDN __host__ __device__
#define PM_NOINLINE __host__ __device__ __noinline__
#else
#define PM_HD inline
#define PM_HDN
#define PM_NOINLINE
#endif

namespace pm {

constexpr double EPSMCH = 2.220446049250313e-16;
constexpr double DWARF = 2.2250738585072014e-308;

PM_HD void cross3(const double *a, const double *b, double *c)
{
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
PM_HD double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
PM_HD double norm3(const double *a) { return sqrt(dot3(a, a)); }
// y = R x, R row-major 3x3
PM_HD void matvec3(const double *R, const double *x, double *y)
{
    y[0] = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
    y[1] = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
    y[2] = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
}

// ---------------------------------------------------------------------------------------------------
// Rotation of rotate_pts (d3_utils.py:212-219) from M = target_c^T source_c (row-major 3x3):
//   U,D,Vh = svd(M); if det(U)det(Vh) < 0: U[:,-1] *= -1; R = U Vh.
// With u3' = u1 x u2 and v3' = v1 x v2 the reflection fix is implicit:  R = u1 v1^T + u2 v2^T + u3' v3'^T.
// One-sided (Hestenes) Jacobi on the columns of M.  Rank-2 M (every 3-point sample) is the normal case and
// is well defined; rank <= 1 (repeated sample indices) is arbitrary in LAPACK too -- we return a
// deterministic completion.  Returns the numerical rank (0..3) for diagnostics.
// ---------------------------------------------------------------------------------------------------
PM_HDN inline int kabsch_rotation(const double *M, double *R)
{
    double A[9], V[9];
    for (int i = 0; i < 9; ++i) { A[i] = M[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 12; ++sweep) {
        int rotated = 0;
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
            const double a = A[p] * A[p] + A[3 + p] * A[3 + p] + A[6 + p] * A[6 + p];
            const double b = A[q] * A[q] + A[3 + q] * A[3 + q] + A[6 + q] * A[6 + q];
            const double c = A[p] * A[q] + A[3 + p] * A[3 + q] + A[6 + p] * A[6 + q];
            if (c == 0.0 || fabs(c) <= 1e-17 * sqrt(a * b)) continue;
            rotated = 1;
            const double zeta = (b - a) / (2.0 * c);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
            for (int i = 0; i < 3; ++i) {
                const double ap = A[3 * i + p], aq = A[3 * i + q];
                A[3 * i + p] = cs * ap - sn * aq;
                A[3 * i + q] = sn * ap + cs * aq;
                const double vp = V[3 * i + p], vq = V[3 * i + q];
                V[3 * i + p] = cs * vp - sn * vq;
                V[3 * i + q] = sn * vp + cs * vq;
            }
        }
        if (!rotated) break;
    }
    double s[3];
    for (int j = 0; j < 3; ++j) s[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
    int i1 = 0;
    if (s[1] > s[i1]) i1 = 1;
    if (s[2] > s[i1]) i1 = 2;
    int i2 = (i1 == 0) ? 1 : 0;
    for (int j = 0; j < 3; ++j)
        if (j != i1 && s[j] > s[i2]) i2 = j;
    const int i3 = 3 - i1 - i2;
    double u1[3], u2[3], u3[3], v1[3], v2[3], v3[3];
    int rank = 0;
    if (s[i1] > 0.0) {
        rank = 1;
        for (int i = 0; i < 3; ++i) { u1[i] = A[3 * i + i1] / s[i1]; v1[i] = V[3 * i + i1]; }
    } else {
        u1[0] = 1; u1[1] = 0; u1[2] = 0; v1[0] = 1; v1[1] = 0; v1[2] = 0;
    }
    if (s[i2] > 1e-14 * s[i1] && s[i2] > 0.0) {
        rank = (s[i3] > 1e-14 * s[i1]) ? 3 : 2;
        for (int i = 0; i < 3; ++i) { u2[i] = A[3 * i + i2] / s[i2]; v2[i] = V[3 * i + i2]; }
        // re-orthogonalise u2 against u1 (V is orthogonal by construction)
        const double d = dot3(u1, u2);
        for (int i = 0; i < 3; ++i) u2[i] -= d * u1[i];
        const double nn = norm3(u2);
        for (int i = 0; i < 3; ++i) u2[i] /= nn;
    } else {
        // degenerate: any unit vectors orthogonal to u1 / v1 (deterministic choice)
        for (int pass = 0; pass < 2; ++pass) {
            const double *w = pass ? v1 : u1;
            double *o = pass ? v2 : u2;
            int k = 0;
            if (fabs(w[1]) < fabs(w[k])) k = 1;
            if (fabs(w[2]) < fabs(w[k])) k = 2;
            double e[3] = {0, 0, 0};
            e[k] = 1.0;
            const double d = dot3(w, e);
            for (int i = 0; i < 3; ++i) o[i] = e[i] - d * w[i];
            const double nn = norm3(o);
            for (int i = 0; i < 3; ++i) o[i] /= nn;
        }
    }
    cross3(u1, u2, u3);
    cross3(v1, v2, v3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j];
    return rank;
}

// singular values (descending) of a 3x3 matrix via the same Jacobi iteration -- used by Umeyama's scale
PM_HDN inline void singular_values3(const double *M, double *sv)
{
    double A[9];
    for (int i = 0; i < 9; ++i) A[i] = M[i];
    for (int sweep = 0; sweep < 12; ++sweep) {
        int rotated = 0;
        for (int pq = 0; pq < 3; ++pq) {
            const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
            const double a = A[p] * A[p] + A[3 + p] * A[3 + p] + A[6 + p] * A[6 + p];
            const double b = A[q] * A[q] + A[3 + q] * A[3 + q] + A[6 + q] * A[6 + q];
            const double c = A[p] * A[q] + A[3 + p] * A[3 + q] + A[6 + p] * A[6 + q];
            if (c == 0.0 || fabs(c) <= 1e-17 * sqrt(a * b)) continue;
            rotated = 1;
            const double zeta = (b - a) / (2.0 * c);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
            for (int i = 0; i < 3; ++i) {
                const double ap = A[3 * i + p], aq = A[3 * i + q];
                A[3 * i + p] = cs * ap - sn * aq;
                A[3 * i + q] = sn * ap + cs * aq;
            }
        }
        if (!rotated) break;
    }
    double s[3];
    for (int j = 0; j < 3; ++j) s[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
    // sort descending
    if (s[0] < s[1]) { double t = s[0]; s[0] = s[1]; s[1] = t; }
    if (s[1] < s[2]) { double t = s[1]; s[1] = s[2]; s[2] = t; }
    if (s[0] < s[1]) { double t = s[0]; s[0] = s[1]; s[1] = t; }
    sv[0] = s[0]; sv[1] = s[1]; sv[2] = s[2];
}

PM_HD double det3(const double *M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// ---------------------------------------------------------------------------------------------------
// scipy Rotation.from_matrix(R).as_rotvec()
// ---------------------------------------------------------------------------------------------------
PM_HDN inline void matrix_to_rotvec(const double *R, double *rv)
{
    const double tr = R[0] + R[4] + R[8];
    double dec[4] = {R[0], R[4], R[8], tr};
    int choice = 0;
    for (int i = 1; i < 4; ++i)
        if (dec[i] > dec[choice]) choice = i;
    double q[4];
    if (choice != 3) {
        const int i = choice, j = (i + 1) % 3, k = (j + 1) % 3;
        q[i] = 1.0 - tr + 2.0 * R[3 * i + i];
        q[j] = R[3 * j + i] + R[3 * i + j];
        q[k] = R[3 * k + i] + R[3 * i + k];
        q[3] = R[3 * k + j] - R[3 * j + k];
    } else {
        q[0] = R[7] - R[5];
        q[1] = R[2] - R[6];
        q[2] = R[3] - R[1];
        q[3] = 1.0 + tr;
    }
    const double qn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= qn;
    if (q[3] < 0.0)
        for (int i = 0; i < 4; ++i) q[i] = -q[i];
    const double nv = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double angle = 2.0 * atan2(nv, q[3]);
    double scale;
    if (angle <= 1e-3) {
        const double a2 = angle * angle;
        scale = 2.0 + a2 / 12.0 + 7.0 * a2 * a2 / 2880.0;
    } else {
        scale = angle / sin(angle / 2.0);
    }
    rv[0] = scale * q[0]; rv[1] = scale * q[1]; rv[2] = scale * q[2];
}

// scipy Rotation.from_rotvec(rv).as_matrix()
PM_HDN inline void rotvec_to_matrix(const double *rv, double *R)
{
    const double angle = norm3(rv);
    double scale;
    if (angle <= 1e-3) {
        const double a2 = angle * angle;
        scale = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0;
    } else {
        scale = sin(angle / 2.0) / angle;
    }
    const double x = scale * rv[0], y = scale * rv[1], z = scale * rv[2], w = cos(angle / 2.0);
    const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    R[0] = x2 - y2 - z2 + w2; R[1] = 2 * (xy - zw);       R[2] = 2 * (xz + yw);
    R[3] = 2 * (xy + zw);     R[4] = -x2 + y2 - z2 + w2;  R[5] = 2 * (yz - xw);
    R[6] = 2 * (xz - yw);     R[7] = 2 * (yz + xw);       R[8] = -x2 - y2 + z2 + w2;
}

// ---------------------------------------------------------------------------------------------------
// rotate_points_with_rotvec (d3_utils.py:150-163) for one point, and precomputed per-rotvec constants
// ---------------------------------------------------------------------------------------------------
struct RotVec {
    double v[3], theta, c, s;
    PM_HD void set(const double *r)
    {
        theta = norm3(r);
        if (theta > 0.0) { const double it = 1.0 / theta; v[0] = r[0] * it; v[1] = r[1] * it; v[2] = r[2] * it; }
        else { v[0] = v[1] = v[2] = 0.0; }                 // nan_to_num(r / 0) == 0
        sincos(theta, &s, &c);
    }
    PM_HD void rotate(const double *p, double *o) const
    {
        double vxp[3];
        cross3(v, p, vxp);
        const double d = dot3(p, v) * (1.0 - c);
        o[0] = c * p[0] + s * vxp[0] + d * v[0];
        o[1] = c * p[1] + s * vxp[1] + d * v[1];
        o[2] = c * p[2] + s * vxp[2] + d * v[2];
    }
    // D[3*i+j] = d(rotate(p))_i / d r_j   (analytic; theta -> 0 limit is -[p]x)
    PM_HD void jacobian(const double *p, double *D) const
    {
        if (theta < 1e-12) {
            D[0] = 0;     D[1] = p[2];  D[2] = -p[1];
            D[3] = -p[2]; D[4] = 0;     D[5] = p[0];
            D[6] = p[1];  D[7] = -p[0]; D[8] = 0;
            return;
        }
        double vxp[3];
        cross3(v, p, vxp);
        const double vp = dot3(v, p);
        double a[3];   // d/dtheta
        for (int i = 0; i < 3; ++i) a[i] = -s * p[i] + c * vxp[i] + vp * s * v[i];
        // G = s * (-[p]x) + (1-c) * (v p^T + (v.p) I)      (d/dv)
        double G[9];
        const double omc = 1.0 - c;
        G[0] = omc * (v[0] * p[0] + vp);        G[1] = s * p[2] + omc * v[0] * p[1];   G[2] = -s * p[1] + omc * v[0] * p[2];
        G[3] = -s * p[2] + omc * v[1] * p[0];   G[4] = omc * (v[1] * p[1] + vp);       G[5] = s * p[0] + omc * v[1] * p[2];
        G[6] = s * p[1] + omc * v[2] * p[0];    G[7] = -s * p[0] + omc * v[2] * p[1];  G[8] = omc * (v[2] * p[2] + vp);
        // D = a v^T + G (I - v v^T) / theta
        const double it = 1.0 / theta;
        for (int i = 0; i < 3; ++i) {
            const double gv = G[3 * i] * v[0] + G[3 * i + 1] * v[1] + G[3 * i + 2] * v[2];
            for (int j = 0; j < 3; ++j) D[3 * i + j] = a[i] * v[j] + (G[3 * i + j] - gv * v[j]) * it;
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// Accumulator of the 6x6 normal equations of objective_eval (parallel_ancsh_pose.py:56-68, isweight=False):
//   residual = [ y0 - Rod(r0) x0 ; y1 - Rod(r1) x1 ; Rod(r0) u - Rod(r1) u  (nj identical rows) ]
// JtJ is stored as a full symmetric 6x6, row-major.
// ---------------------------------------------------------------------------------------------------
struct Normal6 {
    double JtJ[36], Jtf[6], fsq;
    PM_HD void zero()
    {
        for (int i = 0; i < 36; ++i) JtJ[i] = 0.0;
        for (int i = 0; i < 6; ++i) Jtf[i] = 0.0;
        fsq = 0.0;
    }
    // residual block res (3) with d res / d r_blk = sign * D (3x3), blk in {0,1}, weight w
    PM_HD void add_block(int blk, double sign, const double *D, const double *res, double w)
    {
        const int o = 3 * blk;
        for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b)
                JtJ[(o + a) * 6 + o + b] += w * (D[a] * D[b] + D[3 + a] * D[3 + b] + D[6 + a] * D[6 + b]);
            Jtf[o + a] += w * sign * (D[a] * res[0] + D[3 + a] * res[1] + D[6 + a] * res[2]);
        }
    }
    // cross term of the joint rows: J = [ +D0 | -D1 ]
    PM_HD void add_cross(const double *D0, const double *D1, double w)
    {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                const double v = -w * (D0[a] * D1[b] + D0[3 + a] * D1[3 + b] + D0[6 + a] * D1[6 + b]);
                JtJ[a * 6 + 3 + b] += v;
                JtJ[(3 + b) * 6 + a] += v;
            }
    }
};

// point residuals of one part: adds sum |y - Rod(r) x|^2 (and, if N != nullptr, its normal-equation terms)
PM_HD void accum_part(const RotVec &rv, int blk, const double *x, const double *y, Normal6 *N, double &fsq)
{
    double f[3], res[3];
    rv.rotate(x, f);
    res[0] = y[0] - f[0]; res[1] = y[1] - f[1]; res[2] = y[2] - f[2];
    fsq += res[0] * res[0] + res[1] * res[1] + res[2] * res[2];
    if (N) {
        double D[9];
        rv.jacobian(x, D);
        N->add_block(blk, -1.0, D, res, 1.0);
    }
}
PM_HD void accum_joint(const RotVec &r0, const RotVec &r1, const double *u, double nj, Normal6 *N, double &fsq)
{
    double f0[3], f1[3], res[3];
    r0.rotate(u, f0);
    r1.rotate(u, f1);
    res[0] = f0[0] - f1[0]; res[1] = f0[1] - f1[1]; res[2] = f0[2] - f1[2];
    fsq += nj * (res[0] * res[0] + res[1] * res[1] + res[2] * res[2]);
    if (N) {
        double D0[9], D1[9];
        r0.jacobian(u, D0);
        r1.jacobian(u, D1);
        N->add_block(0, 1.0, D0, res, nj);
        N->add_block(1, -1.0, D1, res, nj);
        N->add_cross(D0, D1, nj);
    }
}

// A serial problem evaluator (host tests; one-thread-per-hypothesis on the device).
// Points are stored as contiguous xyz triples.
struct SerialProb {
    const double *x0, *y0, *x1, *y1;
    int n0, n1;
    double u[3];
    double nj;
    PM_HDN double cost(const double *p) const
    {
        RotVec r0, r1;
        r0.set(p);
        r1.set(p + 3);
        double fsq = 0.0;
        for (int i = 0; i < n0; ++i) accum_part(r0, 0, x0 + 3 * i, y0 + 3 * i, nullptr, fsq);
        for (int i = 0; i < n1; ++i) accum_part(r1, 1, x1 + 3 * i, y1 + 3 * i, nullptr, fsq);
        accum_joint(r0, r1, u, nj, nullptr, fsq);
        return fsq;
    }
    PM_HDN void normal(const double *p, Normal6 &N) const
    {
        RotVec r0, r1;
        r0.set(p);
        r1.set(p + 3);
        N.zero();
        double fsq = 0.0;
        for (int i = 0; i < n0; ++i) accum_part(r0, 0, x0 + 3 * i, y0 + 3 * i, &N, fsq);
        for (int i = 0; i < n1; ++i) accum_part(r1, 1, x1 + 3 * i, y1 + 3 * i, &N, fsq);
        accum_joint(r0, r1, u, nj, &N, fsq);
        N.fsq = fsq;
    }
};

// ---------------------------------------------------------------------------------------------------
// MINPACK qrsolv / lmpar (n = 6, diag = 1).  r: 6x6 row-major, upper triangle = R on entry; the strict lower
// triangle is used as workspace for S^T exactly as in MINPACK.
// ---------------------------------------------------------------------------------------------------
constexpr int LMN = 6;

PM_HDN inline void qrsolv6(double *r, const int *ipvt, const double *diag, const double *qtb, double *x, double *sdiag)
{
    double wa[LMN];
    for (int j = 0; j < LMN; ++j) {
        for (int i = j; i < LMN; ++i) r[i * LMN + j] = r[j * LMN + i];
        x[j] = r[j * LMN + j];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < LMN; ++j) {
        const int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < LMN; ++k) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            for (int k = j; k < LMN; ++k) {
                if (sdiag[k] == 0.0) continue;
                double cs, sn;
                const double rkk = r[k * LMN + k];
                if (fabs(rkk) < fabs(sdiag[k])) {
                    const double cotan = rkk / sdiag[k];
                    sn = 0.5 / sqrt(0.25 + 0.25 * cotan * cotan);
                    cs = sn * cotan;
                } else {
                    const double tn = sdiag[k] / rkk;
                    cs = 0.5 / sqrt(0.25 + 0.25 * tn * tn);
                    sn = cs * tn;
                }
                r[k * LMN + k] = cs * rkk + sn * sdiag[k];
                const double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < LMN; ++i) {
                    const double t2 = cs * r[i * LMN + k] + sn * sdiag[i];
                    sdiag[i] = -sn * r[i * LMN + k] + cs * sdiag[i];
                    r[i * LMN + k] = t2;
                }
            }
        }
        sdiag[j] = r[j * LMN + j];
        r[j * LMN + j] = x[j];
    }
    int nsing = LMN;
    for (int j = 0; j < LMN; ++j) {
        if (sdiag[j] == 0.0 && nsing == LMN) nsing = j;
        if (nsing < LMN) wa[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        const int j = nsing - k;
        double sum = 0.0;
        for (int i = j + 1; i < nsing; ++i) sum += r[i * LMN + j] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < LMN; ++j) x[ipvt[j]] = wa[j];
}

PM_HDN inline void lmpar6(double *r, const int *ipvt, const double *diag, const double *qtb, double delta, double *par,
                          double *x, double *sdiag)
{
    double wa1[LMN], wa2[LMN];
    int nsing = LMN;
    for (int j = 0; j < LMN; ++j) {
        wa1[j] = qtb[j];
        if (r[j * LMN + j] == 0.0 && nsing == LMN) nsing = j;
        if (nsing < LMN) wa1[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        const int j = nsing - k;
        wa1[j] /= r[j * LMN + j];
        const double temp = wa1[j];
        for (int i = 0; i < j; ++i) wa1[i] -= r[i * LMN + j] * temp;
    }
    for (int j = 0; j < LMN; ++j) x[ipvt[j]] = wa1[j];
    int iter = 0;
    double dxnorm = 0.0;
    for (int j = 0; j < LMN; ++j) { wa2[j] = diag[j] * x[j]; dxnorm += wa2[j] * wa2[j]; }
    dxnorm = sqrt(dxnorm);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { *par = 0.0; return; }
    double parl = 0.0;
    if (nsing >= LMN) {
        for (int j = 0; j < LMN; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < LMN; ++j) {
            double sum = 0.0;
            for (int i = 0; i < j; ++i) sum += r[i * LMN + j] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j * LMN + j];
        }
        double temp = 0.0;
        for (int j = 0; j < LMN; ++j) temp += wa1[j] * wa1[j];
        temp = sqrt(temp);
        parl = ((fp / delta) / temp) / temp;
    }
    double gnorm = 0.0;
    for (int j = 0; j < LMN; ++j) {
        double sum = 0.0;
        for (int i = 0; i <= j; ++i) sum += r[i * LMN + j] * qtb[i];
        const int l = ipvt[j];
        wa1[j] = sum / diag[l];
        gnorm += wa1[j] * wa1[j];
    }
    gnorm = sqrt(gnorm);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = DWARF / fmin(delta, 0.1);
    *par = fmax(*par, parl);
    *par = fmin(*par, paru);
    if (*par == 0.0) *par = gnorm / dxnorm;
    for (;;) {
        ++iter;
        if (*par == 0.0) *par = fmax(DWARF, 0.001 * paru);
        double temp = sqrt(*par);
        for (int j = 0; j < LMN; ++j) wa1[j] = temp * diag[j];
        qrsolv6(r, ipvt, wa1, qtb, x, sdiag);
        dxnorm = 0.0;
        for (int j = 0; j < LMN; ++j) { wa2[j] = diag[j] * x[j]; dxnorm += wa2[j] * wa2[j]; }
        dxnorm = sqrt(dxnorm);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        for (int j = 0; j < LMN; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < LMN; ++j) {
            wa1[j] /= sdiag[j];
            const double t2 = wa1[j];
            for (int i = j + 1; i < LMN; ++i) wa1[i] -= r[i * LMN + j] * t2;
        }
        temp = 0.0;
        for (int j = 0; j < LMN; ++j) temp += wa1[j] * wa1[j];
        temp = sqrt(temp);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = fmax(parl, *par);
        if (fp < 0.0) paru = fmin(paru, *par);
        *par = fmax(parl, *par + parc);
    }
}

// pivoted Cholesky of the symmetric 6x6 A = J^T J with qrfac's pivot rule (largest remaining column norm):
// R (upper, row-major, zero below the diagonal) with R^T R = P^T A P, ipvt, acnorm[j] = |column j of J|.
PM_HDN inline void chol_pivot6(const double *Ain, double *R, int *ipvt, double *acnorm)
{
    double A[36];
    for (int i = 0; i < 36; ++i) { A[i] = Ain[i]; R[i] = 0.0; }
    double dmax = 0.0;
    for (int j = 0; j < LMN; ++j) {
        ipvt[j] = j;
        acnorm[j] = sqrt(fmax(Ain[j * LMN + j], 0.0));
        dmax = fmax(dmax, Ain[j * LMN + j]);
    }
    const double tiny = dmax * 1e-28;
    for (int j = 0; j < LMN; ++j) {
        int kmax = j;
        for (int k = j + 1; k < LMN; ++k)
            if (A[k * LMN + k] > A[kmax * LMN + kmax]) kmax = k;
        if (kmax != j) {
            for (int i = 0; i < LMN; ++i) { const double t = A[i * LMN + j]; A[i * LMN + j] = A[i * LMN + kmax]; A[i * LMN + kmax] = t; }
            for (int i = 0; i < LMN; ++i) { const double t = A[j * LMN + i]; A[j * LMN + i] = A[kmax * LMN + i]; A[kmax * LMN + i] = t; }
            for (int i = 0; i < j; ++i) { const double t = R[i * LMN + j]; R[i * LMN + j] = R[i * LMN + kmax]; R[i * LMN + kmax] = t; }
            const int t = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = t;
        }
        const double d = A[j * LMN + j];
        if (!(d > tiny)) {          // rank deficient from here on
            for (int k = j; k < LMN; ++k)
                for (int i = j; i < LMN; ++i) R[k * LMN + i] = 0.0;
            break;
        }
        const double rjj = sqrt(d);
        R[j * LMN + j] = rjj;
        for (int k = j + 1; k < LMN; ++k) R[j * LMN + k] = A[j * LMN + k] / rjj;
        for (int k = j + 1; k < LMN; ++k)
            for (int l = j + 1; l < LMN; ++l) A[k * LMN + l] -= R[j * LMN + k] * R[j * LMN + l];
    }
}

struct LmResult {
    int info, nfev, njev;
    double fnorm;
};

// MINPACK lmder (mode 2, diag = 1) on the normal equations.  Prob must provide
//   double cost(const double *x)            -> sum of squared residuals
//   void normal(const double *x, Normal6&)  -> J^T J, J^T f (and fsq)
template <class Prob>
PM_HDN inline LmResult lm_solve(const Prob &prob, double *x, double ftol, double xtol, double gtol, int maxfev,
                                double factor)
{
    LmResult res;
    res.nfev = 1; res.njev = 0; res.info = 0;
    double diag[LMN];
    for (int j = 0; j < LMN; ++j) diag[j] = 1.0;
    double fnorm = sqrt(prob.cost(x));
    double par = 0.0, delta = 0.0, xnorm = 0.0;
    int iter = 1;
    Normal6 N;
    double Rm[36], acnorm[LMN], qtf[LMN], p[LMN], sdiag[LMN], xnew[LMN];
    int ipvt[LMN];
    for (;;) {
        prob.normal(x, N);
        ++res.njev;
        chol_pivot6(N.JtJ, Rm, ipvt, acnorm);
        // (Q^T f)[:n] = R^-T P^T J^T f
        for (int j = 0; j < LMN; ++j) {
            double sum = N.Jtf[ipvt[j]];
            for (int i = 0; i < j; ++i) sum -= Rm[i * LMN + j] * qtf[i];
            qtf[j] = (Rm[j * LMN + j] != 0.0) ? sum / Rm[j * LMN + j] : 0.0;
        }
        if (iter == 1) {
            xnorm = 0.0;
            for (int j = 0; j < LMN; ++j) xnorm += (diag[j] * x[j]) * (diag[j] * x[j]);
            xnorm = sqrt(xnorm);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        double gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < LMN; ++j) {
                const int l = ipvt[j];
                if (acnorm[l] == 0.0) continue;
                double sum = 0.0;
                for (int i = 0; i <= j; ++i) sum += Rm[i * LMN + j] * (qtf[i] / fnorm);
                gnorm = fmax(gnorm, fabs(sum / acnorm[l]));
            }
        }
        if (gnorm <= gtol) { res.info = 4; break; }
        double ratio = 0.0;
        bool done = false;
        do {
            lmpar6(Rm, ipvt, diag, qtf, delta, &par, p, sdiag);
            double pnorm = 0.0;
            for (int j = 0; j < LMN; ++j) {
                p[j] = -p[j];
                xnew[j] = x[j] + p[j];
                pnorm += (diag[j] * p[j]) * (diag[j] * p[j]);
            }
            pnorm = sqrt(pnorm);
            if (iter == 1) delta = fmin(delta, pnorm);
            const double fnorm1 = sqrt(prob.cost(xnew));
            ++res.nfev;
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            double wa3[LMN];
            for (int j = 0; j < LMN; ++j) wa3[j] = 0.0;
            for (int j = 0; j < LMN; ++j) {
                const double temp = p[ipvt[j]];
                for (int i = 0; i <= j; ++i) wa3[i] += Rm[i * LMN + j] * temp;
            }
            double t1 = 0.0;
            for (int j = 0; j < LMN; ++j) t1 += wa3[j] * wa3[j];
            const double temp1 = sqrt(t1) / fnorm;
            const double temp2 = (sqrt(par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = (prered != 0.0) ? actred / prered : 0.0;
            if (ratio <= 0.25) {
                double temp = (actred >= 0.0) ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * fmin(delta, pnorm / 0.1);
                par /= temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par *= 0.5;
            }
            if (ratio >= 1e-4) {
                xnorm = 0.0;
                for (int j = 0; j < LMN; ++j) { x[j] = xnew[j]; xnorm += (diag[j] * x[j]) * (diag[j] * x[j]); }
                xnorm = sqrt(xnorm);
                fnorm = fnorm1;
                ++iter;
            }
            const bool fconv = fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0;
            if (fconv) res.info = 1;
            if (delta <= xtol * xnorm) res.info = 2;
            if (fconv && res.info == 2) res.info = 3;
            if (res.info != 0) { done = true; break; }
            if (res.nfev >= maxfev) res.info = 5;
            if (fabs(actred) <= EPSMCH && prered <= EPSMCH && 0.5 * ratio <= 1.0) res.info = 6;
            if (delta <= EPSMCH * xnorm) res.info = 7;
            if (gnorm <= EPSMCH) res.info = 8;
            if (res.info != 0) { done = true; break; }
        } while (ratio < 1e-4);
        if (done) break;
    }
    res.fnorm = fnorm;
    return res;
}

#include "lm_fast.cuh"
#include "lm_tick.cuh"

// ---------------------------------------------------------------------------------------------------
// scale_pts (d3_utils.py:237-246) for a handful of points (all ordered pairs; i==j pairs contribute 0)
// ---------------------------------------------------------------------------------------------------
PM_HDN inline double pair_scale_small(const double *src, const double *tgt, int n)
{
    double ab = 0.0, aa = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            if (i == j) continue;
            double ds[3] = {src[3 * i] - src[3 * j], src[3 * i + 1] - src[3 * j + 1], src[3 * i + 2] - src[3 * j + 2]};
            double dt[3] = {tgt[3 * i] - tgt[3 * j], tgt[3 * i + 1] - tgt[3 * j + 1], tgt[3 * i + 2] - tgt[3 * j + 2]};
            const double A = norm3(ds), b = norm3(dt);
            ab += A * b;
            aa += A * A;
        }
    return ab / (aa + 1e-6);
}

// transform_pts (d3_utils.py:223-234) for 3 sampled points: R (row-major), scale, translation
PM_HDN inline void transform3(const double *src, const double *tgt, double *R, double *scale, double *t)
{
    double ms[3] = {0, 0, 0}, mt[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) { ms[c] += src[3 * i + c]; mt[c] += tgt[3 * i + c]; }
    for (int c = 0; c < 3; ++c) { ms[c] /= 3.0; mt[c] /= 3.0; }
    double sc[9], tc[9];
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) { sc[3 * i + c] = src[3 * i + c] - ms[c]; tc[3 * i + c] = tgt[3 * i + c] - mt[c]; }
    double M[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) M[3 * a + b] = tc[a] * sc[b] + tc[3 + a] * sc[3 + b] + tc[6 + a] * sc[6 + b];
    kabsch_rotation(M, R);
    *scale = pair_scale_small(sc, tc, 3);
    // translation = mean(target - scale * R source)
    double acc[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) {
        double rs[3];
        matvec3(R, src + 3 * i, rs);
        for (int c = 0; c < 3; ++c) acc[c] += tgt[3 * i + c] - (*scale) * rs[c];
    }
    for (int c = 0; c < 3; ++c) t[c] = acc[c] / 3.0;
}

// ---------------------------------------------------------------------------------------------------
// joint_transformation_estimator (evaluation/parallel_ancsh_pose.py:106-184) for one RANSAC hypothesis:
// 3 sampled correspondences per part.  model = {R0[9], s0, t0[3], R1[9], s1, t1[3]} (26 doubles).
// LM settings of the reference call: ftol=1e-4, xtol=gtol=1e-8 (scipy defaults), max_nfev=100*n, factor=100.
// ---------------------------------------------------------------------------------------------------
struct JointModel {
    double R0[9], s0, t0[3], R1[9], s1, t1[3];
};

PM_HDN inline void centre_scale3(const double *S, const double *T, double *Sc, double *Tsc, double *scale)
{
    *scale = pair_scale_small(S, T, 3);                        // :121
    const double sinv = pair_scale_small(T, S, 3);             // :123 scale_pts(target, source)
    double ms[3] = {0, 0, 0}, mt[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) { ms[c] += S[3 * i + c]; mt[c] += sinv * T[3 * i + c]; }
    for (int c = 0; c < 3; ++c) { ms[c] /= 3.0; mt[c] /= 3.0; }
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) { Sc[3 * i + c] = S[3 * i + c] - ms[c]; Tsc[3 * i + c] = sinv * T[3 * i + c] - mt[c]; }
}

PM_HDN inline void kabsch_points(const double *Sc, const double *Tc, int n, double *R)
{
    // rotate_pts re-centres its (already centred) inputs, d3_utils.py:211-212
    double ms[3] = {0, 0, 0}, mt[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) { ms[c] += Sc[3 * i + c]; mt[c] += Tc[3 * i + c]; }
    for (int c = 0; c < 3; ++c) { ms[c] /= n; mt[c] /= n; }
    double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) M[3 * a + b] += (Tc[3 * i + a] - mt[a]) * (Sc[3 * i + b] - ms[b]);
    kabsch_rotation(M, R);
}

PM_HDN inline void mean_translation(const double *S, const double *T, int n, const double *R, double scale, double *t)
{
    double acc[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        double rs[3];
        matvec3(R, S + 3 * i, rs);
        for (int c = 0; c < 3; ++c) acc[c] += T[3 * i + c] - scale * rs[c];
    }
    for (int c = 0; c < 3; ++c) t[c] = acc[c] / n;
}

// state != nullptr: resumable (see lm_fast.cuh); the model is only valid when the returned info != LM_SUSPENDED.
PM_HDN inline LmResult joint_estimate3(const double *S0, const double *T0, const double *S1, const double *T1,
                                       const double *u, JointModel &m, LmState *state = nullptr, int budget = 0x7fffffff)
{
    double S0c[9], T0c[9], S1c[9], T1c[9];
    centre_scale3(S0, T0, S0c, T0c, &m.s0);
    centre_scale3(S1, T1, S1c, T1c, &m.s1);
    double x[6];
    if (!(state && state->iter > 0)) {
        kabsch_points(S0c, T0c, 3, m.R0);                      // :138-139
        kabsch_points(S1c, T1c, 3, m.R1);
        matrix_to_rotvec(m.R0, x);                             // :147-148
        matrix_to_rotvec(m.R1, x + 3);
    }
    SerialProb P;
    P.x0 = S0c; P.y0 = T0c; P.n0 = 3; P.x1 = S1c; P.y1 = T1c; P.n1 = 3;
    P.u[0] = u[0]; P.u[1] = u[1]; P.u[2] = u[2];
    P.nj = 3.0;                                                // min(n0,n1) copies of the joint direction, :134
    LmResult r = lm_solve_fast(P, x, 1e-4, 1e-8, 1e-8, 600, 100.0, state, budget); // :154-155
    if (r.info == LM_SUSPENDED) return r;
    rotvec_to_matrix(x, m.R0);                                 // :156-157
    rotvec_to_matrix(x + 3, m.R1);
    mean_translation(S0, T0, 3, m.R0, m.s0, m.t0);             // :174-175 (un-refined scale)
    mean_translation(S1, T1, 3, m.R1, m.s1, m.t1);
    return r;
}

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator for hypothesis sampling (hard part 4).
// ---------------------------------------------------------------------------------------------------
PM_HD void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned *out)
{
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
        const unsigned n1 = (unsigned)p1;
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
        const unsigned n3 = (unsigned)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// three sample positions in [0,n) for hypothesis `hyp` of problem `prob`, stream 0 (single) / 1,2 (joint parts)
PM_HD void sample3(unsigned long long seed, unsigned prob, unsigned hyp, unsigned stream, int n, int *idx)
{
    unsigned r[4];
    philox4x32_10(hyp, prob, stream, 0u, (unsigned)seed, (unsigned)(seed >> 32), r);
    for (int i = 0; i < 3; ++i) idx[i] = (int)(((unsigned long long)r[i] * (unsigned long long)n) >> 32);
}

}  // namespace pm
