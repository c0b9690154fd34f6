// lm_tick.cuh -- MINPACK lmder/lmpar for n = 6, diag = 1 (mode 2) as an explicit state machine, written for SIMT.
//
// Why a third restatement (after the literal port lm_solve and the register-resident lm_solve_fast):
// with one LM solve per thread the ncu profile of lm_solve_fast showed 11.7 of 32 lanes active and 55% of the stall
// samples waiting for instructions (240 KB of SASS): lanes of a warp sit in different places of
//     for (;;) { jacobian; factor; do { lmpar; trial step; } while (step rejected); }
// -- a lane that repeats the inner loop runs alone, a lane that converged idles until the slowest lane of its warp is
// done, and the physically pivoted 6x6 factorisation diverges 5 ways.  Here ONE call of lm_tick() performs
//     [Jacobian phase, only if the previous step was accepted]  ->  lmpar  ->  trial step  ->  update / tests
// and returns; everything that must survive between two calls lives in LmTick.  A warp therefore executes the same
// code for lanes that retry a step and lanes that start a new outer iteration, and a lane whose solve has finished can
// be handed the next solve between two ticks (pose.cu: joint_lm_kernel).
//
// Numerics (all identities exact in real arithmetic, so the iterates equal MINPACK's up to rounding):
//   * no column pivoting.  The LM step p = (J^T J + par I)^-1 J^T f, the Gauss-Newton step, lmpar's Newton correction
//     w^T (J^T J + par I)^-1 w and the predicted reduction |J p|^2 = p^T (J^T J) p do not depend on the column order;
//     qrfac's pivoting only matters when J^T J is singular (the basic solution qrsolv returns then depends on the
//     order).  A zero pivot in the unpivoted Cholesky sends the lane through gn_rank_deficient(), the literal pivoted
//     code path (degenerate samples only);
//   * R^T (Q^T f) = P^T J^T f, so the scaled gradient test uses g = J^T f directly;
//   * Cholesky of (A + par I) instead of qrsolv's Givens rotations (same matrix, see lm_fast.cuh).
// Included from pose_math.cuh (inside namespace pm).
#pragma once

PM_HD double rsqrt_f64(double x)
{
#ifdef __CUDA_ARCH__
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

// packed upper triangle of a symmetric 6x6: (i, j), i <= j
PM_HD constexpr int ut6(int i, int j) { return i * 6 - (i * (i - 1)) / 2 + (j - i); }

// Upper Cholesky S^T S = A + shift I on the packed triangle.  Only what the triangular solves need is produced: the
// off-diagonal entries of S (packed like A; diagonal slots unused) and rinv[j] = 1 / S[j][j].  A pivot that is not
// > tiny ends the factorisation: that row and all later ones are treated as zero; returns the index of that pivot.
PM_HD int chol6p(const double (&A)[21], double shift, double tiny, double (&S)[21], double (&rinv)[6])
{
    int nsing = 6;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = A[ut6(j, j)] + shift;
#pragma unroll
        for (int k = 0; k < j; ++k) d -= S[ut6(k, j)] * S[ut6(k, j)];
        const bool ok = d > tiny && nsing == 6;
        const double inv = ok ? rsqrt_f64(d) : 0.0;
        if (!ok && nsing == 6) nsing = j;
        rinv[j] = inv;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = A[ut6(j, i)];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= S[ut6(k, j)] * S[ut6(k, i)];
            S[ut6(j, i)] = v * inv;
        }
    }
    return nsing;
}
// w = S^-T b
PM_HD void solve_lower6p(const double (&S)[21], const double (&rinv)[6], const double (&b)[6], double (&w)[6])
{
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double v = b[j];
#pragma unroll
        for (int i = 0; i < j; ++i) v -= S[ut6(i, j)] * w[i];
        w[j] = v * rinv[j];
    }
}
// x = S^-1 b
PM_HD void solve_upper6p(const double (&S)[21], const double (&rinv)[6], const double (&b)[6], double (&x)[6])
{
#pragma unroll
    for (int jj = 0; jj < 6; ++jj) {
        const int j = 5 - jj;
        double v = b[j];
#pragma unroll
        for (int i = j + 1; i < 6; ++i) v -= S[ut6(j, i)] * x[i];
        x[j] = v * rinv[j];
    }
}

// Gauss-Newton step of a rank-deficient J^T J exactly as qrfac + qrsolv define it (pivoted basic solution).  Rare
// (repeated sample indices) and therefore out of line: it costs nothing unless a lane takes it.
PM_NOINLINE void gn_rank_deficient(const double *JtJ, const double *Jtf, double *gn)
{
    double R[36], acn[LMN], qtf[LMN], wa[LMN];
    int ipvt[LMN];
    chol_pivot6(JtJ, R, ipvt, acn);
    int nsing = LMN;
#pragma unroll 1
    for (int j = 0; j < LMN; ++j) {
        double sum = Jtf[ipvt[j]];
#pragma unroll 1
        for (int i = 0; i < j; ++i) sum -= R[i * LMN + j] * qtf[i];
        qtf[j] = (R[j * LMN + j] != 0.0) ? sum / R[j * LMN + j] : 0.0;
        if (R[j * LMN + j] == 0.0 && nsing == LMN) nsing = j;
        wa[j] = (nsing < LMN) ? 0.0 : qtf[j];
    }
#pragma unroll 1
    for (int k = 1; k <= nsing; ++k) {
        const int j = nsing - k;
        wa[j] /= R[j * LMN + j];
        const double temp = wa[j];
#pragma unroll 1
        for (int i = 0; i < j; ++i) wa[i] -= R[i * LMN + j] * temp;
    }
#pragma unroll 1
    for (int j = 0; j < LMN; ++j) gn[ipvt[j]] = wa[j];
}

struct LmTick {
    // ---- control state of lmder ----
    double x[6], fnorm, par, delta, xnorm;
    int iter, nfev, njev;
    int nlm;           // lmpar iterations so far (statistics only: feeds the FLOP model of bench.py's roofline)
    int info;          // 0 while the solve is running, MINPACK's info code afterwards
    int need_jac;      // the next tick starts with a Jacobian evaluation (the previous step was accepted)
    // ---- products of the Jacobian phase, valid while need_jac == 0 ----
    double A[21];      // J^T J, packed upper triangle
    double g[6];       // J^T f
    double gn[6];      // Gauss-Newton step A^-1 g
    double dxn;        // |gn|
    double gnt;        // |R^-T (gn / |gn|)|  (lmpar's first lower bound parl = (fp / delta) / gnt^2); 0 if rank deficient
    double gnorm;      // max_j |g_j| / (|J e_j| fnorm)
};

PM_HD void lm_tick_init(LmTick &s, const double *x0, double fnorm0)
{
#pragma unroll
    for (int j = 0; j < 6; ++j) s.x[j] = x0[j];
    s.fnorm = fnorm0; s.par = 0.0; s.delta = 0.0; s.xnorm = 0.0;
    s.iter = 1; s.nfev = 1; s.njev = 0; s.nlm = 0; s.info = 0; s.need_jac = 1;
    s.dxn = 0.0; s.gnt = 0.0; s.gnorm = 0.0;
}

// One trial step of lmder.  Prob: double cost(const double *x) (sum of squares), void normal(const double *x, Normal6 &).
template <class Prob>
PM_HD void lm_tick(const Prob &prob, LmTick &s, double ftol, double xtol, double gtol, int maxfev, double factor)
{
    if (s.need_jac) {
        double tiny = 0.0;
        {
            Normal6 N;
            prob.normal(s.x, N);
            ++s.njev;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int j = i; j < 6; ++j) s.A[ut6(i, j)] = N.JtJ[i * 6 + j];
                s.g[i] = N.Jtf[i];
                tiny = fmax(tiny, N.JtJ[i * 6 + i]);
            }
            tiny *= 1e-28;
            double S[21], rinv[6];
            const int nsing = chol6p(s.A, 0.0, tiny, S, rinv);
            if (nsing == 6) {
                double w[6];
                solve_lower6p(S, rinv, s.g, w);
                solve_upper6p(S, rinv, w, s.gn);
                s.dxn = norm6(s.gn);
                if (s.dxn > 0.0) {
                    const double idx = 1.0 / s.dxn;
                    double w1[6], w2[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) w1[j] = s.gn[j] * idx;
                    solve_lower6p(S, rinv, w1, w2);
                    s.gnt = norm6(w2);
                } else {
                    s.gnt = 0.0;
                }
            } else {
                // copies: taking the address of N itself would force the accumulator of the hot path into local memory
                double JtJc[36], Jtfc[6], gnc[6];
                for (int i = 0; i < 36; ++i) JtJc[i] = N.JtJ[i];
                for (int i = 0; i < 6; ++i) Jtfc[i] = N.Jtf[i];
                gn_rank_deficient(JtJc, Jtfc, gnc);
                for (int i = 0; i < 6; ++i) s.gn[i] = gnc[i];
                s.dxn = norm6(s.gn);
                s.gnt = 0.0;
            }
        }
        if (s.iter == 1) {
            s.xnorm = norm6(s.x);
            s.delta = factor * s.xnorm;
            if (s.delta == 0.0) s.delta = factor;
        }
        double gnorm = 0.0;
        if (s.fnorm != 0.0) {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const double ajj = s.A[ut6(j, j)];
                if (ajj > 0.0) gnorm = fmax(gnorm, fabs(s.g[j]) * rsqrt_f64(ajj));
            }
            gnorm /= s.fnorm;
        }
        s.gnorm = gnorm;
        if (gnorm <= gtol) { s.info = 4; return; }
        s.need_jac = 0;
    }
    // ---------------- lmpar (diag = 1) ----------------
    double p[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) p[j] = s.gn[j];
    const double delta = s.delta;
    double par = s.par;
    {
        double dxnorm = s.dxn;
        double fp = dxnorm - delta;
        if (fp <= 0.1 * delta) {
            par = 0.0;
        } else {
            double parl = (s.gnt > 0.0) ? ((fp / delta) / s.gnt) / s.gnt : 0.0;
            const double gradnorm = norm6(s.g);
            double paru = gradnorm / delta;
            if (paru == 0.0) paru = DWARF / fmin(delta, 0.1);
            par = fmax(par, parl);
            par = fmin(par, paru);
            if (par == 0.0) par = gradnorm / dxnorm;
            for (int it = 1;; ++it) {
                if (par == 0.0) par = fmax(DWARF, 0.001 * paru);
                ++s.nlm;
                double S[21], rinv[6], w[6];
                chol6p(s.A, par, 0.0, S, rinv);                    // S^T S = J^T J + par I
                solve_lower6p(S, rinv, s.g, w);
                solve_upper6p(S, rinv, w, p);
                dxnorm = norm6(p);
                const double temp = fp;
                fp = dxnorm - delta;
                if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || it == 10) break;
                double w1[6], w2[6];
                const double idx = 1.0 / dxnorm;
#pragma unroll
                for (int j = 0; j < 6; ++j) w1[j] = p[j] * idx;
                solve_lower6p(S, rinv, w1, w2);
                double t2 = 0.0;
#pragma unroll
                for (int j = 0; j < 6; ++j) t2 += w2[j] * w2[j];
                const double parc = (fp / delta) / t2;             // ((fp / delta) / t) / t with t = |w2|
                if (fp > 0.0) parl = fmax(parl, par);
                if (fp < 0.0) paru = fmin(paru, par);
                par = fmax(parl, par + parc);
            }
        }
    }
    // ---------------- trial step ----------------
    double xnew[6], pnorm = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) { p[j] = -p[j]; pnorm += p[j] * p[j]; xnew[j] = s.x[j] + p[j]; }
    pnorm = sqrt(pnorm);
    if (s.iter == 1) s.delta = fmin(s.delta, pnorm);
    const double fnorm1 = sqrt(prob.cost(xnew));
    ++s.nfev;
    const double inv_f = 1.0 / s.fnorm;
    double actred = -1.0;
    if (0.1 * fnorm1 < s.fnorm) { const double r = fnorm1 * inv_f; actred = 1.0 - r * r; }
    double pAp = 0.0;                                              // |J p|^2
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double row = 0.5 * s.A[ut6(i, i)] * p[i];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) row += s.A[ut6(i, j)] * p[j];
        pAp += 2.0 * row * p[i];
    }
    const double temp1sq = fmax(pAp, 0.0) * inv_f * inv_f;
    const double temp2sq = par * (pnorm * inv_f) * (pnorm * inv_f);
    const double prered = temp1sq + temp2sq / 0.5;
    const double dirder = -(temp1sq + temp2sq);
    const double ratio = (prered != 0.0) ? actred / prered : 0.0;
    if (ratio <= 0.25) {
        double temp = (actred >= 0.0) ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
        if (0.1 * fnorm1 >= s.fnorm || temp < 0.1) temp = 0.1;
        s.delta = temp * fmin(s.delta, pnorm / 0.1);
        par /= temp;
    } else if (par == 0.0 || ratio >= 0.75) {
        s.delta = pnorm / 0.5;
        par *= 0.5;
    }
    s.par = par;
    if (ratio >= 1e-4) {                                           // successful iteration
#pragma unroll
        for (int j = 0; j < 6; ++j) s.x[j] = xnew[j];
        s.xnorm = norm6(s.x);
        s.fnorm = fnorm1;
        ++s.iter;
        s.need_jac = 1;
    }
    const bool fconv = fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0;
    int info = 0;
    if (fconv) info = 1;
    if (s.delta <= xtol * s.xnorm) info = 2;
    if (fconv && info == 2) info = 3;
    if (info == 0) {
        if (s.nfev >= maxfev) info = 5;
        if (fabs(actred) <= EPSMCH && prered <= EPSMCH && 0.5 * ratio <= 1.0) info = 6;
        if (s.delta <= EPSMCH * s.xnorm) info = 7;
        if (s.gnorm <= EPSMCH) info = 8;
    }
    s.info = info;
}

// Runs a solve to completion with lm_tick (host tests; the kernels drive the ticks themselves).
template <class Prob>
PM_HDN inline LmResult lm_solve_tick(const Prob &prob, double *x, double ftol, double xtol, double gtol, int maxfev,
                                     double factor)
{
    LmTick s;
    lm_tick_init(s, x, sqrt(prob.cost(x)));
    while (s.info == 0) lm_tick(prob, s, ftol, xtol, gtol, maxfev, factor);
    for (int j = 0; j < 6; ++j) x[j] = s.x[j];
    LmResult res;
    res.info = s.info; res.nfev = s.nfev; res.njev = s.njev; res.fnorm = s.fnorm;
    return res;
}
