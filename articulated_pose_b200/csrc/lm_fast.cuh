// lm_fast.cuh -- register-resident restatement of MINPACK lmder/lmpar for n = 6, diag = 1 (mode 2).
//
// Mathematically the same iteration as pm::lm_solve (pose_math.cuh, the literal qrfac/qrsolv port kept as the
// host-side cross-check), reorganised for a GPU thread:
//   * the column pivoting of qrfac is applied PHYSICALLY (rows/columns of J^T J are swapped with predicated moves),
//     so every array index is a compile-time constant after unrolling and the 6x6 factors live in registers
//     instead of local memory;
//   * qrsolv's Givens elimination of sqrt(par)*I is replaced by a Cholesky factorisation of (P^T J^T J P + par I):
//     S^T S = R^T R + par I is exactly the matrix qrsolv triangularises, so the step p = (A + par I)^-1 g and the
//     Newton correction ||S^-T (p/|p|)|| are identical up to rounding; 6 sqrt + 6 div instead of ~21 rotations.
// Included from pose_math.cuh (inside namespace pm).
#pragma once

// (Making the 6x6 primitives __noinline__ was measured: the code shrinks by 3% only -- the f64 sqrt / div / sincos
// expansions dominate the ~230 KB of SASS -- and the by-reference arrays cost 50% in the block-cooperative refit.)
// upper Cholesky of the symmetric 6x6 A (full storage) plus `shift` on the diagonal: S^T S = A + shift I.
// Pivots that are not > tiny are treated as zero (row zeroed); returns the index of the first zero pivot (6 = none).
// rinv[j] receives 1 / S[j][j] (0 for zero pivots): the triangular solves multiply instead of dividing -- f64 division
// and sqrt are long software sequences on the GPU and dominate the latency of one LM iteration otherwise.
PM_HD int chol6(const double (&A)[6][6], double shift, double tiny, double (&S)[6][6], double (&rinv)[6])
{
    int nsing = 6;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = A[j][j] + shift;
#pragma unroll
        for (int k = 0; k < j; ++k) d -= S[k][j] * S[k][j];
        const bool ok = d > tiny && nsing == 6;
        const double inv = ok ? 1.0 / sqrt(d) : 0.0;
        const double sjj = ok ? d * inv : 0.0;
        if (!ok && nsing == 6) nsing = j;
        S[j][j] = sjj;
        rinv[j] = inv;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = A[j][i];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= S[k][j] * S[k][i];
            S[j][i] = v * inv;
        }
#pragma unroll
        for (int i = 0; i < j; ++i) S[j][i] = 0.0;
    }
    return nsing;
}

// w = S^-T b  (forward substitution with the upper factor), rows >= nsing give 0
PM_HD void solve_lower6(const double (&S)[6][6], const double (&rinv)[6], int nsing, const double (&b)[6], double (&w)[6])
{
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double v = b[j];
#pragma unroll
        for (int i = 0; i < j; ++i) v -= S[i][j] * w[i];
        w[j] = (j < nsing) ? v * rinv[j] : 0.0;
    }
}
// x = S^-1 b  (back substitution), components >= nsing are 0
PM_HD void solve_upper6(const double (&S)[6][6], const double (&rinv)[6], int nsing, const double (&b)[6], double (&x)[6])
{
#pragma unroll
    for (int jj = 0; jj < 6; ++jj) {
        const int j = 5 - jj;
        double v = b[j];
#pragma unroll
        for (int i = j + 1; i < 6; ++i) v -= S[j][i] * x[i];
        x[j] = (j < nsing) ? v * rinv[j] : 0.0;
    }
}
PM_HD double norm6(const double (&v)[6])
{
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += v[j] * v[j];
    return sqrt(s);
}

// Resumable state of an LM solve, valid at the top of an outer iteration (before the Jacobian is evaluated).
struct LmState {
    double x[6], fnorm, par, delta, xnorm;
    int iter, nfev, njev;
};
constexpr int LM_SUSPENDED = -1;

// `state` != nullptr makes the solve resumable: a call with state->iter == 0 starts fresh, otherwise it continues
// from the saved state; when res.nfev reaches `budget` at the top of an outer iteration the state is saved and
// info = LM_SUSPENDED is returned.  The sequence of iterates is independent of where the solve was suspended.
template <class Prob>
PM_HDN inline LmResult lm_solve_fast(const Prob &prob, double *x, double ftol, double xtol, double gtol, int maxfev,
                                     double factor, LmState *state = nullptr, int budget = 0x7fffffff)
{
    LmResult res;
    res.nfev = 1; res.njev = 0; res.info = 0;
    double fnorm, par = 0.0, delta = 0.0, xnorm = 0.0;
    int iter = 1;
    if (state && state->iter > 0) {
#pragma unroll
        for (int j = 0; j < 6; ++j) x[j] = state->x[j];
        fnorm = state->fnorm; par = state->par; delta = state->delta; xnorm = state->xnorm;
        iter = state->iter; res.nfev = state->nfev; res.njev = state->njev;
    } else {
        fnorm = sqrt(prob.cost(x));
    }
    for (;;) {
        if (state && res.nfev >= budget) {
#pragma unroll
            for (int j = 0; j < 6; ++j) state->x[j] = x[j];
            state->fnorm = fnorm; state->par = par; state->delta = delta; state->xnorm = xnorm;
            state->iter = iter; state->nfev = res.nfev; state->njev = res.njev;
            res.info = LM_SUSPENDED;
            res.fnorm = fnorm;
            return res;
        }
        double A[6][6], g[6], acn[6];
        int perm[6];
        {
            Normal6 N;
            prob.normal(x, N);
            ++res.njev;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int j = 0; j < 6; ++j) A[i][j] = N.JtJ[i * 6 + j];
                g[i] = N.Jtf[i];
                perm[i] = i;
            }
        }
        // ---- qrfac's column pivoting, applied physically: A <- P^T A P, g <- P^T g -------------------------
        // (selection by the diagonal of the running Schur complement = squared remaining column norms)
        {
            double sd[6];
            double dmax = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) { sd[j] = A[j][j]; dmax = fmax(dmax, sd[j]); acn[j] = sqrt(fmax(A[j][j], 0.0)); }
            // Determine the pivot order with a throw-away elimination on a copy, then permute A once.
            double W[6][6];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) W[i][j] = A[i][j];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                int kmax = j;
                double best = W[j][j];
#pragma unroll
                for (int k = j + 1; k < 6; ++k)
                    if (W[k][k] > best) { best = W[k][k]; kmax = k; }
                // swap rows/cols j <-> kmax of W and A, entries of g, acn, perm (predicated, constant indices)
#pragma unroll
                for (int k = j + 1; k < 6; ++k) {
                    if (k == kmax) {
#pragma unroll
                        for (int i = 0; i < 6; ++i) {
                            double t = W[i][j]; W[i][j] = W[i][k]; W[i][k] = t;
                            t = A[i][j]; A[i][j] = A[i][k]; A[i][k] = t;
                        }
#pragma unroll
                        for (int i = 0; i < 6; ++i) {
                            double t = W[j][i]; W[j][i] = W[k][i]; W[k][i] = t;
                            t = A[j][i]; A[j][i] = A[k][i]; A[k][i] = t;
                        }
                        double t = g[j]; g[j] = g[k]; g[k] = t;
                        t = acn[j]; acn[j] = acn[k]; acn[k] = t;
                        const int ti = perm[j]; perm[j] = perm[k]; perm[k] = ti;
                    }
                }
                // eliminate column j from the trailing block of W (Schur complement)
                const double d = W[j][j];
                const double inv = (d > dmax * 1e-28) ? 1.0 / d : 0.0;
#pragma unroll
                for (int k = j + 1; k < 6; ++k)
#pragma unroll
                    for (int l = j + 1; l < 6; ++l) W[k][l] -= W[j][k] * W[j][l] * inv;
            }
            (void)sd;
        }
        double tiny = 0.0;
#pragma unroll
        for (int j = 0; j < 6; ++j) tiny = fmax(tiny, A[j][j]);
        tiny *= 1e-28;
        double R[6][6], Rinv[6];
        const int nsing = chol6(A, 0.0, tiny, R, Rinv);
        double qtf[6];
        solve_lower6(R, Rinv, nsing, g, qtf);                      // (Q^T f)[:n] = R^-T P^T J^T f
        if (iter == 1) {
            xnorm = 0.0;
            for (int j = 0; j < 6; ++j) xnorm += x[j] * x[j];
            xnorm = sqrt(xnorm);
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        double gnorm = 0.0;
        if (fnorm != 0.0) {
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                // sum_{i<=j} R[i][j] qtf[i] (== (P^T J^T f)[j] for a full-rank factor; kept literal for rank deficiency)
                double sum = 0.0;
#pragma unroll
                for (int i = 0; i <= j; ++i) sum += R[i][j] * (qtf[i] / fnorm);
                if (acn[j] != 0.0) gnorm = fmax(gnorm, fabs(sum / acn[j]));
            }
        }
        if (gnorm <= gtol) { res.info = 4; break; }
        // quantities of lmpar that do not depend on par
        double gn[6];                                        // Gauss-Newton direction (permuted)
        solve_upper6(R, Rinv, nsing, qtf, gn);
        double rtq[6];                                       // R^T qtb
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double sum = 0.0;
#pragma unroll
            for (int i = 0; i <= j; ++i) sum += R[i][j] * qtf[i];
            rtq[j] = sum;
        }
        const double gradnorm = norm6(rtq);
        double ratio = 0.0;
        bool done = false;
        do {
            // ---------------- lmpar (diag = 1) ----------------
            double p[6];
            {
#pragma unroll
                for (int j = 0; j < 6; ++j) p[j] = gn[j];
                double dxnorm = norm6(p);
                double fp = dxnorm - delta;
                if (fp <= 0.1 * delta) {
                    par = 0.0;
                } else {
                    double parl = 0.0;
                    if (nsing >= 6) {
                        double w1[6], w2[6];
                        const double idx = 1.0 / dxnorm;
#pragma unroll
                        for (int j = 0; j < 6; ++j) w1[j] = p[j] * idx;
                        solve_lower6(R, Rinv, 6, w1, w2);
                        const double t = norm6(w2);
                        parl = ((fp / delta) / t) / t;
                    }
                    double paru = gradnorm / delta;
                    if (paru == 0.0) paru = DWARF / fmin(delta, 0.1);
                    par = fmax(par, parl);
                    par = fmin(par, paru);
                    if (par == 0.0) par = gradnorm / dxnorm;
                    for (int it = 1;; ++it) {
                        if (par == 0.0) par = fmax(DWARF, 0.001 * paru);
                        double S[6][6], Sinv[6];
                        const int ns = chol6(A, par, 0.0, S, Sinv);  // S^T S = R^T R + par I
                        double w[6];
                        solve_lower6(S, Sinv, ns, rtq, w);
                        solve_upper6(S, Sinv, ns, w, p);
                        dxnorm = norm6(p);
                        const double temp = fp;
                        fp = dxnorm - delta;
                        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || it == 10) break;
                        double w1[6], w2[6];
                        const double idx = 1.0 / dxnorm;
#pragma unroll
                        for (int j = 0; j < 6; ++j) w1[j] = p[j] * idx;
                        solve_lower6(S, Sinv, ns, w1, w2);
                        const double t = norm6(w2);
                        const double parc = ((fp / delta) / t) / t;
                        if (fp > 0.0) parl = fmax(parl, par);
                        if (fp < 0.0) paru = fmin(paru, par);
                        par = fmax(parl, par + parc);
                    }
                }
            }
            // ---------------- trial step ----------------
            double pnorm = 0.0, xnew[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) { p[j] = -p[j]; pnorm += p[j] * p[j]; }
            pnorm = sqrt(pnorm);
            // un-permute: step for variable perm[j] is p[j]
#pragma unroll
            for (int l = 0; l < 6; ++l) {
                double v = 0.0;
#pragma unroll
                for (int j = 0; j < 6; ++j) v = (perm[j] == l) ? p[j] : v;
                xnew[l] = x[l] + v;
            }
            if (iter == 1) delta = fmin(delta, pnorm);
            const double fnorm1 = sqrt(prob.cost(xnew));
            ++res.nfev;
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) actred = 1.0 - (fnorm1 / fnorm) * (fnorm1 / fnorm);
            double t1 = 0.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                double s = 0.0;
#pragma unroll
                for (int j = i; j < 6; ++j) s += R[i][j] * p[j];
                t1 += s * s;
            }
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
#pragma unroll
                for (int j = 0; j < 6; ++j) { x[j] = xnew[j]; xnorm += x[j] * x[j]; }
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
