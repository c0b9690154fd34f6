// tests/hostsim/pose_math_host.cpp -- TEST INFRASTRUCTURE: compiles the __host__ __device__ pose math of
// articulated_pose_b200/csrc/pose_math.cuh for the CPU so it can be checked against numpy / scipy without a GPU.
// This library is never loaded by the product path.
#include "../../articulated_pose_b200/csrc/pose_math.cuh"

extern "C" {
int hs_kabsch(const double *M, double *R) { return pm::kabsch_rotation(M, R); }
void hs_singular_values(const double *M, double *s) { pm::singular_values3(M, s); }
void hs_matrix_to_rotvec(const double *R, double *rv) { pm::matrix_to_rotvec(R, rv); }
void hs_rotvec_to_matrix(const double *rv, double *R) { pm::rotvec_to_matrix(rv, R); }
void hs_rodrigues(const double *r, const double *p, double *out, double *D)
{
    pm::RotVec rv;
    rv.set(r);
    rv.rotate(p, out);
    rv.jacobian(p, D);
}
void hs_transform3(const double *src, const double *tgt, double *R, double *s, double *t) { pm::transform3(src, tgt, R, s, t); }
double hs_pair_scale(const double *src, const double *tgt, int n) { return pm::pair_scale_small(src, tgt, n); }
// returns info; out[0]=nfev, out[1]=njev, out[2]=fnorm
int hs_lm(const double *x0, const double *y0, int n0, const double *x1, const double *y1, int n1, const double *u, double nj,
          double *x, double ftol, double xtol, double gtol, int maxfev, double factor, double *out)
{
    pm::SerialProb P;
    P.x0 = x0; P.y0 = y0; P.n0 = n0; P.x1 = x1; P.y1 = y1; P.n1 = n1;
    P.u[0] = u[0]; P.u[1] = u[1]; P.u[2] = u[2];
    P.nj = nj;
    pm::LmResult r = pm::lm_solve(P, x, ftol, xtol, gtol, maxfev, factor);
    out[0] = r.nfev; out[1] = r.njev; out[2] = r.fnorm;
    return r.info;
}
// the register-resident variant the kernels run (lm_fast.cuh)
int hs_lm_fast(const double *x0, const double *y0, int n0, const double *x1, const double *y1, int n1, const double *u,
               double nj, double *x, double ftol, double xtol, double gtol, int maxfev, double factor, double *out)
{
    pm::SerialProb P;
    P.x0 = x0; P.y0 = y0; P.n0 = n0; P.x1 = x1; P.y1 = y1; P.n1 = n1;
    P.u[0] = u[0]; P.u[1] = u[1]; P.u[2] = u[2];
    P.nj = nj;
    pm::LmResult r = pm::lm_solve_fast(P, x, ftol, xtol, gtol, maxfev, factor);
    out[0] = r.nfev; out[1] = r.njev; out[2] = r.fnorm;
    return r.info;
}
// the SIMT state-machine variant (lm_tick.cuh) the joint estimation kernels run
int hs_lm_tick(const double *x0, const double *y0, int n0, const double *x1, const double *y1, int n1, const double *u,
               double nj, double *x, double ftol, double xtol, double gtol, int maxfev, double factor, double *out)
{
    pm::SerialProb P;
    P.x0 = x0; P.y0 = y0; P.n0 = n0; P.x1 = x1; P.y1 = y1; P.n1 = n1;
    P.u[0] = u[0]; P.u[1] = u[1]; P.u[2] = u[2];
    P.nj = nj;
    pm::LmResult r = pm::lm_solve_tick(P, x, ftol, xtol, gtol, maxfev, factor);
    out[0] = r.nfev; out[1] = r.njev; out[2] = r.fnorm;
    return r.info;
}
int hs_joint_estimate3(const double *S0, const double *T0, const double *S1, const double *T1, const double *u,
                       double *model26, double *out)
{
    pm::JointModel m;
    pm::LmResult r = pm::joint_estimate3(S0, T0, S1, T1, u, m);
    const double *src = reinterpret_cast<const double *>(&m);
    for (int i = 0; i < 26; ++i) model26[i] = src[i];
    out[0] = r.nfev; out[1] = r.njev; out[2] = r.fnorm;
    return r.info;
}
// the same estimate, suspended every `budget` residual evaluations and resumed (what the two-phase kernels do)
int hs_joint_estimate3_resumed(const double *S0, const double *T0, const double *S1, const double *T1, const double *u,
                               int budget, double *model26, double *out)
{
    pm::JointModel m;
    pm::LmState st;
    st.iter = 0;
    pm::LmResult r;
    int limit = budget, rounds = 0;
    do {
        r = pm::joint_estimate3(S0, T0, S1, T1, u, m, &st, limit);
        limit += budget;
        ++rounds;
    } while (r.info == pm::LM_SUSPENDED);
    const double *src = reinterpret_cast<const double *>(&m);
    for (int i = 0; i < 26; ++i) model26[i] = src[i];
    out[0] = r.nfev; out[1] = r.njev; out[2] = rounds;
    return r.info;
}
void hs_sample3(unsigned long long seed, unsigned prob, unsigned hyp, unsigned stream, int n, int *idx)
{
    pm::sample3(seed, prob, hyp, stream, n, idx);
}
void hs_normal(const double *x0, const double *y0, int n0, const double *x1, const double *y1, int n1, const double *u,
               double nj, const double *x, double *JtJ, double *Jtf, double *fsq)
{
    pm::SerialProb P;
    P.x0 = x0; P.y0 = y0; P.n0 = n0; P.x1 = x1; P.y1 = y1; P.n1 = n1;
    P.u[0] = u[0]; P.u[1] = u[1]; P.u[2] = u[2];
    P.nj = nj;
    pm::Normal6 N;
    P.normal(x, N);
    for (int i = 0; i < 36; ++i) JtJ[i] = N.JtJ[i];
    for (int i = 0; i < 6; ++i) Jtf[i] = N.Jtf[i];
    *fsq = N.fsq;
}
}
