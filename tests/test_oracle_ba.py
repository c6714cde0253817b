"""CPU suite: pins the C++ oracle of the BA path (rows a1-a8) against the independent
numpy restatement, finite differences (convention of ProjectionFactor::check,
projection_factor.cpp:176-223: right-multiplicative deltaQ perturbation, eps=1e-6) and
a second solver (dense damped Gauss-Newton) for the fixed point."""
import ctypes as C

import numpy as np
import pytest

import np_ref


@pytest.fixture(scope="module")
def mods(pkg, oracle):
    return pkg.abi, pkg.synth, oracle


def _rand_pose(rng, scale=1.0):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return np.concatenate([rng.normal(size=3) * scale, q])


def test_projection_factor_vs_numpy_and_fd(mods):
    abi, synth, orc = mods
    rng = np.random.default_rng(0)
    for _ in range(20):
        pose_i, pose_j = _rand_pose(rng, 0.3), _rand_pose(rng, 0.3)
        pose_j[3:] = np_ref.pose_plus(pose_i, np.concatenate([np.zeros(3), rng.normal(size=3) * 0.2]))[3:]
        ex = np.concatenate([synth.EUROC_TIC, synth.rot_to_quat(synth.EUROC_RIC)])
        ex[3:] /= np.linalg.norm(ex[3:])
        pts_i = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4), 1.0])
        pts_j = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4), 1.0])
        inv_dep = rng.uniform(0.1, 0.5)
        res = np.zeros(2)
        Ji, Jj, Jex, Jf = np.zeros(14), np.zeros(14), np.zeros(14), np.zeros(2)
        orc.oracle_projection_factor(abi.dptr(pts_i), abi.dptr(pts_j), abi.dptr(pose_i), abi.dptr(pose_j),
                                     abi.dptr(ex), inv_dep, 460 / 1.5, abi.dptr(res), abi.dptr(Ji), abi.dptr(Jj),
                                     abi.dptr(Jex), abi.dptr(Jf))
        r, nJi, nJj, nJex, nJf = np_ref.projection(pts_i, pts_j, pose_i, pose_j, ex, inv_dep)
        sc = max(1.0, np.abs(r).max())
        assert np.allclose(res, r, rtol=1e-11, atol=1e-10 * sc)
        for a, b in ((Ji, nJi), (Jj, nJj), (Jex, nJex)):
            a = a.reshape(2, 7)
            assert np.all(a[:, 6] == 0)
            assert np.allclose(a[:, :6], b, rtol=1e-10, atol=1e-9 * np.abs(b).max())
        assert np.allclose(Jf, nJf, rtol=1e-10)
        # finite differences through the oracle's own residual
        eps = 1e-6

        def rr(pi, pj, e, lam):
            out = np.zeros(2)
            orc.oracle_projection_factor(abi.dptr(pts_i), abi.dptr(pts_j), abi.dptr(pi), abi.dptr(pj), abi.dptr(e),
                                         lam, 460 / 1.5, abi.dptr(out), None, None, None, None)
            return out

        num = np.zeros((2, 19))
        for k in range(6):
            d = np.zeros(6)
            d[k] = eps
            num[:, k] = (rr(np_ref.pose_plus(pose_i, d), pose_j, ex, inv_dep) - res) / eps
            num[:, 6 + k] = (rr(pose_i, np_ref.pose_plus(pose_j, d), ex, inv_dep) - res) / eps
            num[:, 12 + k] = (rr(pose_i, pose_j, np_ref.pose_plus(ex, d), inv_dep) - res) / eps
        num[:, 18] = (rr(pose_i, pose_j, ex, inv_dep + eps) - res) / eps
        ana = np.hstack([Ji.reshape(2, 7)[:, :6], Jj.reshape(2, 7)[:, :6], Jex.reshape(2, 7)[:, :6], Jf.reshape(2, 1)])
        assert np.allclose(num, ana, rtol=2e-3, atol=2e-3 * np.abs(ana).max())


def test_preintegration_oracle_vs_numpy(mods):
    abi, synth, orc = mods
    rng = np.random.default_rng(1)
    ba, bg = rng.normal(0, 0.02, 3), rng.normal(0, 0.002, 3)
    a0, g0 = rng.normal(0, 1, 3) + [0, 0, 9.8], rng.normal(0, 0.3, 3)
    pre = synth.Preintegration(a0, g0, ba, bg)
    c = abi.Preint()
    c.delta_q[3] = 1.0
    for i in range(3):
        c.lin_ba[i], c.lin_bg[i] = ba[i], bg[i]
    for i in range(15):
        c.jacobian[i * 15 + i] = 1.0
    pa, pg = a0.copy(), g0.copy()
    for _ in range(20):
        a1, g1 = rng.normal(0, 1, 3) + [0, 0, 9.8], rng.normal(0, 0.3, 3)
        pre.push_back(0.005, a1, g1)
        orc.oracle_preint_propagate(C.byref(c), 0.005, abi.dptr(pa), abi.dptr(pg), abi.dptr(a1), abi.dptr(g1),
                                    synth.ACC_N, synth.GYR_N, synth.ACC_W, synth.GYR_W)
        pa, pg = a1.copy(), g1.copy()
    packed = synth.pack_preint(pre)
    got = np.frombuffer(bytes(c), dtype=np.float64)
    assert np.allclose(got[:17], packed[:17], rtol=1e-12, atol=1e-15)
    assert np.allclose(got[17:242], packed[17:242], rtol=1e-10, atol=1e-14)
    assert np.allclose(got[242:], packed[242:], rtol=1e-9, atol=1e-20)


def test_imu_factor_vs_numpy_and_fd(mods):
    abi, synth, orc = mods
    w = synth.make_window(seed=2, K=4, L=10)
    G = np.array([0, 0, synth.G_NORM])
    for j in range(1, 4):
        pre_c = C.cast(abi.WindowHandle(w).pre.ctypes.data + j * 467 * 8, C.POINTER(abi.Preint))
        h = abi.WindowHandle(w)
        pre_c = C.cast(h.pre.ctypes.data + j * 467 * 8, C.POINTER(abi.Preint))
        pi, si, pj, sj = w.para_pose[j - 1].copy(), w.para_speed_bias[j - 1].copy(), w.para_pose[j].copy(), w.para_speed_bias[j].copy()
        res = np.zeros(15)
        Jpi, Jsi, Jpj, Jsj = np.zeros(105), np.zeros(135), np.zeros(105), np.zeros(135)
        orc.oracle_imu_factor(pre_c, abi.dptr(G), abi.dptr(pi), abi.dptr(si), abi.dptr(pj), abi.dptr(sj), abi.dptr(res),
                              abi.dptr(Jpi), abi.dptr(Jsi), abi.dptr(Jpj), abi.dptr(Jsj))
        pre = np_ref.unpack_preint(w.preint[j])
        r, nJpi, nJsi, nJpj, nJsj = np_ref.imu(pre, G, pi, si, pj, sj)
        # sqrt_info is built from an inverse of a matrix with cond ~1e9: agreement ~1e-7 relative is the
        # best two different inverse algorithms (LU here, LAPACK there) can do.
        assert np.allclose(res, r, rtol=1e-6, atol=1e-6 * np.abs(r).max())
        assert np.allclose(Jpi.reshape(15, 7)[:, :6], nJpi, rtol=1e-6, atol=1e-6 * np.abs(nJpi).max())
        assert np.allclose(Jsi.reshape(15, 9), nJsi, rtol=1e-6, atol=1e-6 * np.abs(nJsi).max())
        assert np.allclose(Jpj.reshape(15, 7)[:, :6], nJpj, rtol=1e-6, atol=1e-6 * np.abs(nJpj).max())
        assert np.allclose(Jsj.reshape(15, 9), nJsj, rtol=1e-6, atol=1e-6 * np.abs(nJsj).max())
        assert np.all(Jpi.reshape(15, 7)[:, 6] == 0) and np.all(Jpj.reshape(15, 7)[:, 6] == 0)
        # sqrt_info^T sqrt_info == cov^-1
        sq = np.zeros(225)
        cov = np.ascontiguousarray(pre["covariance"]).reshape(-1).copy()
        orc.oracle_imu_sqrt_info(abi.dptr(cov), abi.dptr(sq))
        sq = sq.reshape(15, 15)
        assert np.allclose(sq, np.triu(sq))
        lhs = sq.T @ sq @ pre["covariance"]
        assert np.allclose(lhs, np.eye(15), atol=1e-6)
        # finite differences (unwhitened effect removed by comparing whitened residuals)
        eps = 1e-6

        def rr(a, b, c, d):
            out = np.zeros(15)
            orc.oracle_imu_factor(pre_c, abi.dptr(G), abi.dptr(a), abi.dptr(b), abi.dptr(c), abi.dptr(d), abi.dptr(out),
                                  None, None, None, None)
            return out

        ana = np.hstack([Jpi.reshape(15, 7)[:, :6], Jsi.reshape(15, 9), Jpj.reshape(15, 7)[:, :6], Jsj.reshape(15, 9)])
        num = np.zeros((15, 30))
        for k in range(6):
            d = np.zeros(6)
            d[k] = eps
            num[:, k] = (rr(np_ref.pose_plus(pi, d), si, pj, sj) - res) / eps
            num[:, 15 + k] = (rr(pi, si, np_ref.pose_plus(pj, d), sj) - res) / eps
        for k in range(9):
            d = np.zeros(9)
            d[k] = eps
            num[:, 6 + k] = (rr(pi, si + d, pj, sj) - res) / eps
            num[:, 21 + k] = (rr(pi, si, pj, sj + d) - res) / eps
        # the reference's O_R/O_BG jacobian uses delta_q instead of corrected_delta_q (imu_factor.h:124) and a
        # first-order Qleft/Qright form: agreement with FD is ~1e-3 relative, not exact
        assert np.allclose(num, ana, rtol=5e-3, atol=5e-3 * np.abs(ana).max())


def test_prior_residual_vs_numpy(mods):
    abi, synth, orc = mods
    rng = np.random.default_rng(5)
    w = synth.make_window(seed=5, K=5, L=12)
    n = 6 + 9 + 6 + 6
    J = rng.normal(size=(n, n))
    w.prior = dict(n=n, block_kind=np.array([0, 1, 2, 0], np.int32), block_frame=np.array([0, 0, 0, 2], np.int32),
                   block_idx=np.array([0, 6, 15, 21], np.int32),
                   x0=np.concatenate([w.gt_pose[0], w.gt_speed_bias[0], w.para_ex_pose, -w.gt_pose[2]]),
                   lin_jac=J.reshape(-1, order="F").copy(), lin_res=rng.normal(size=n))
    # x0 of the last block has a negated quaternion -> exercises the w<0 sign flip
    h = abi.WindowHandle(w)
    res, dx = np.zeros(n), np.zeros(n)
    orc.oracle_prior_residual(C.byref(h.prior_s), C.byref(h.s), abi.dptr(res), abi.dptr(dx))
    r2, dx2, _ = np_ref.prior_residual(w.prior, w)
    # the negated-position x0 makes dx huge for block 3 positions; compare exactly anyway
    assert np.allclose(dx, dx2, rtol=1e-12, atol=1e-13)
    assert np.allclose(res, r2, rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("seed,K,L", [(0, 2, 20), (1, 5, 30), (2, 11, 40)])
def test_linearize_vs_numpy(mods, seed, K, L):
    abi, synth, orc = mods
    w = synth.make_window(seed=seed, K=K, L=L)
    h = abi.WindowHandle(w)
    o = abi.default_opts()
    npar = 15 * K
    S, g, hh, bb = np.zeros(npar * npar), np.zeros(npar), np.zeros(L), np.zeros(L)
    cost = C.c_double()
    assert orc.oracle_linearize(C.byref(h.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(hh), abi.dptr(bb),
                                C.cast(C.byref(cost), abi.c_double_p)) == 0
    S2, g2, h2, b2, cost2 = np_ref.reduced_system(w)
    S = S.reshape(npar, npar)
    assert abs(cost.value - cost2) <= 1e-7 * cost2          # limited by sqrt_info (see imu test)
    assert np.allclose(hh, h2, rtol=1e-10)
    assert np.allclose(bb, b2, rtol=1e-9, atol=1e-9 * np.abs(b2).max())
    assert np.allclose(S, S.T, rtol=1e-12, atol=1e-6)
    assert np.allclose(S, S2, rtol=1e-6, atol=1e-6 * np.abs(S2).max())
    assert np.allclose(g, g2, rtol=1e-6, atol=1e-6 * np.abs(g2).max())
    assert abs(orc.oracle_cost(C.byref(h.s), C.byref(o)) - cost.value) <= 1e-12 * cost.value


@pytest.mark.parametrize("seed,K,L,ex,td", [(0, 5, 30, 1, 0), (1, 5, 30, 0, 1), (2, 11, 40, 1, 1), (3, 2, 20, 1, 1)])
def test_linearize_with_free_extrinsics_and_td_vs_numpy(mods, seed, K, L, ex, td):
    """Reduced system with para_Ex_Pose and / or para_Td free (estimator.cpp:672-683, 732-740): the oracle's block
    assembly + Schur elimination against the dense numpy system with the extra columns."""
    abi, synth, orc = mods
    kw = dict(track_min=2, track_max=2) if K == 2 else {}
    w = synth.make_window(seed=seed, K=K, L=L, td_true=0.004, **kw)
    w.para_td[0] = 0.001
    rng = np.random.default_rng(seed)
    w.para_ex_pose[:3] += rng.normal(0, 0.02, 3)
    h = abi.WindowHandle(w)
    o = abi.default_opts(estimate_extrinsic=ex, estimate_td=td, TR=0.02)
    npar = 15 * K + 6 * ex + td
    S, g, hh, bb = np.zeros(npar * npar), np.zeros(npar), np.zeros(L), np.zeros(L)
    cost = C.c_double()
    assert orc.oracle_linearize(C.byref(h.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(hh), abi.dptr(bb),
                                C.cast(C.byref(cost), abi.c_double_p)) == 0
    S2, g2, h2, b2, cost2 = np_ref.reduced_system_ext(w, est_ex=bool(ex), est_td=bool(td), TR=0.02)
    S = S.reshape(npar, npar)
    assert abs(cost.value - cost2) <= 1e-7 * cost2
    assert np.allclose(hh, h2, rtol=1e-10) and np.allclose(bb, b2, rtol=1e-9, atol=1e-9 * np.abs(b2).max())
    assert np.allclose(S, S2, rtol=1e-6, atol=1e-6 * np.abs(S2).max())
    assert np.allclose(g, g2, rtol=1e-6, atol=1e-6 * np.abs(g2).max())
    assert np.abs(S2[15 * K:, :]).max() > 0


def _solve(orc, abi, w, **kw):
    h = abi.WindowHandle(w)
    o = abi.default_opts(**kw)
    s = abi.Summary()
    assert orc.oracle_optimize(C.byref(h.s), C.byref(o), C.byref(s)) == 0
    return h, s


# The reference's IMU bias random walk is so stiff (information ~5e14, integration_base.h:21-27 with
# GYR_W = 2e-6) that the absolute gradient floors at ~1e-5 from roundoff: convergence is declared on the
# function / parameter tolerances.
TIGHT = dict(max_iters=60, function_tolerance=1e-13, gradient_tolerance=1e-10, parameter_tolerance=1e-13)


@pytest.mark.parametrize("seed,K,L", [(0, 2, 20), (3, 6, 30), (4, 11, 40)])
def test_lm_dogleg_and_dense_gn_reach_same_fixed_point(mods, seed, K, L):
    abi, synth, orc = mods
    w = synth.make_window(seed=seed, K=K, L=L)
    h_lm, s_lm = _solve(orc, abi, w, strategy=0, **TIGHT)
    h_dl, s_dl = _solve(orc, abi, w, strategy=1, **dict(TIGHT, max_iters=200))
    assert s_lm.final_cost < s_lm.initial_cost and s_dl.final_cost < s_dl.initial_cost
    assert s_lm.termination in (1, 2, 3) and s_dl.termination in (1, 2, 3)
    x_lm, x_dl = h_lm.state_vector(), h_dl.state_vector()
    assert np.linalg.norm(x_lm - x_dl) / np.linalg.norm(x_dl) < 1e-6
    assert abs(s_lm.final_cost - s_dl.final_cost) <= 1e-9 * s_dl.final_cost
    # independent solver (dense damped GN in numpy)
    w_gn, cost_gn, gmax = np_ref.solve_gn(w)
    x_gn = np.concatenate([w_gn.para_pose.ravel(), w_gn.para_speed_bias.ravel(), w_gn.para_ex_pose, w_gn.inv_depth])
    assert gmax < 1e-2          # roundoff floor of the absolute gradient (information up to 5e14)
    assert np.linalg.norm(x_lm - x_gn) / np.linalg.norm(x_gn) < 1e-6
    assert abs(s_lm.final_cost - cost_gn) <= 1e-6 * cost_gn


@pytest.mark.parametrize("seed,K,L,strategy,radius,iters", [(7, 6, 30, 0, 1e4, 8), (8, 6, 30, 1, 1e4, 8), (9, 5, 25, 1, 1.0, 20),
                                                            (10, 5, 25, 1, 0.05, 25), (11, 11, 40, 0, 1e4, 8), (12, 5, 25, 0, 1e-2, 15)])
def test_trust_region_trajectory_matches_dense_numpy(mods, seed, K, L, strategy, radius, iters):
    """The oracle's Schur-eliminated LM / dogleg loop against the same Ceres logic run on the DENSE unreduced system in
    numpy (tests/np_ref.solve_trust_region): same number of iterations, same accept / reject path, same final radius,
    cost and state.  Small radii force the Cauchy-point and interpolated dogleg branches."""
    abi, synth, orc = mods
    w = synth.make_window(seed=seed, K=K, L=L)
    h, s = _solve(orc, abi, w, strategy=strategy, initial_radius=radius, max_iters=iters)
    wn, trace, term = np_ref.solve_trust_region(w, strategy=strategy, initial_radius=radius, max_iters=iters)
    assert len(trace) == s.iterations or (term in (1, 3) and len(trace) + 1 == s.iterations), (len(trace), s.as_dict())
    assert sum(t[2] for t in trace) == s.num_accepted and term == s.termination
    assert abs(trace[-1][3] - s.final_radius) <= 1e-6 * s.final_radius
    final_cost = trace[-1][1] if trace[-1][2] else trace[-1][0]
    assert abs(final_cost - s.final_cost) <= 1e-8 * s.final_cost
    xo = h.state_vector()
    xn = np.concatenate([wn.para_pose.ravel(), wn.para_speed_bias.ravel(), wn.para_ex_pose, wn.inv_depth])
    assert np.linalg.norm(xo - xn) <= 1e-7 * np.linalg.norm(xn)


def test_reference_defaults_terminate_like_ceres(mods):
    """8 iterations / default tolerances (config/euroc/euroc_config.yaml:54-55)."""
    abi, synth, orc = mods
    w = synth.make_window(seed=7, K=11, L=60)
    for strat in (0, 1):
        h, s = _solve(orc, abi, w, strategy=strat)
        assert s.iterations <= 8 and s.num_accepted >= 1
        assert s.final_cost < 0.5 * s.initial_cost
        assert s.termination in (0, 1, 2, 3)


def test_double2vector_restores_gauge(mods):
    abi, synth, orc = mods
    w = synth.make_window(seed=9, K=6, L=20)
    rng = np.random.default_rng(0)
    yaw, t = 0.3, rng.normal(size=3)
    Rz = synth.euler_zyx(yaw, 0, 0)
    pose, sb = w.para_pose.copy(), w.para_speed_bias.copy()
    for i in range(w.K):       # move the whole window by a yaw + translation (the BA gauge)
        pose[i, :3] = Rz @ w.para_pose[i, :3] + t
        pose[i, 3:] = synth.rot_to_quat(Rz @ synth.quat_to_rot(w.para_pose[i, 3:]))
        sb[i, :3] = Rz @ w.para_speed_bias[i, :3]
    pre0 = w.para_pose[0].copy()
    orc.oracle_double2vector(abi.dptr(pre0), w.K, abi.dptr(pose), abi.dptr(sb))
    for i in range(w.K):
        assert np.allclose(pose[i, :3], w.para_pose[i, :3], atol=1e-12)
        q = pose[i, 3:] * np.sign(pose[i, 6]) * np.sign(w.para_pose[i, 6])
        assert np.allclose(q, w.para_pose[i, 3:], atol=1e-12)
        assert np.allclose(sb[i, :3], w.para_speed_bias[i, :3], atol=1e-12)


# ---------------------------------------------------------------------------------------------
# a2': ProjectionTdFactor (projection_td_factor.cpp:34-141)
# ---------------------------------------------------------------------------------------------
def _td_factor(orc, abi, x, pts_i, pts_j, vel_i, vel_j, tdi, tdj, rowi, rowj, TR, ROW, jac=True):
    pose_i, pose_j, ex, lam, td = x[:7].copy(), x[7:14].copy(), x[14:21].copy(), float(x[21]), float(x[22])
    res = np.zeros(2)
    Ji, Jj, Je, Jf, Jt = np.zeros(14), np.zeros(14), np.zeros(14), np.zeros(2), np.zeros(2)
    orc.oracle_projection_td_factor(abi.dptr(pts_i), abi.dptr(pts_j), abi.dptr(vel_i), abi.dptr(vel_j), tdi, tdj, rowi,
                                    rowj, TR, ROW, abi.dptr(pose_i), abi.dptr(pose_j), abi.dptr(ex), lam, td, 460 / 1.5,
                                    abi.dptr(res), *(abi.dptr(a) if jac else None for a in (Ji, Jj, Je, Jf, Jt)))
    return res, Ji.reshape(2, 7), Jj.reshape(2, 7), Je.reshape(2, 7), Jf, Jt


def test_td_factor_jacobians_match_finite_differences(pkg, oracle):
    """The check() convention of the reference (projection_td_factor.cpp:143-233): perturb each local coordinate
    (pose blocks through PoseLocalParameterization::Plus) and compare with the analytic Jacobians."""
    abi, S = pkg.abi, pkg.synth
    rng = np.random.default_rng(0)
    w = S.make_window(seed=4, K=6, L=8, td_true=0.004)
    k_i, k_j = int(w.lm_obs_offset[2]), int(w.lm_obs_offset[2]) + 2
    fi, fj = int(w.obs_frame[k_i]), int(w.obs_frame[k_j])
    pts_i, pts_j = np.array([*w.obs_xy[k_i], 1.0]), np.array([*w.obs_xy[k_j], 1.0])
    args = (pts_i, pts_j, w.obs_vel[k_i].copy(), w.obs_vel[k_j].copy(), 0.001, -0.002, float(w.obs_row[k_i]),
            float(w.obs_row[k_j]), 0.03, 480.0)
    x = np.concatenate([w.gt_pose[fi], w.gt_pose[fj], w.para_ex_pose, [w.gt_inv_depth[2]], [0.003]])
    r0, Ji, Jj, Je, Jf, Jt = _td_factor(oracle, abi, x, *args)
    assert np.abs(Ji[:, 6]).max() == 0 and np.abs(Jj[:, 6]).max() == 0 and np.abs(Je[:, 6]).max() == 0
    ana = np.hstack([Ji[:, :6], Jj[:, :6], Je[:, :6], Jf[:, None], Jt[:, None]])
    num = np.zeros_like(ana)
    eps = 1e-6

    def plus(pose, d):
        q = S.quat_mul(pose[3:], np.array([d[3] / 2, d[4] / 2, d[5] / 2, 1.0]))
        return np.concatenate([pose[:3] + d[:3], q / np.linalg.norm(q)])

    for c in range(20):
        xp = x.copy()
        d = np.zeros(6)
        if c < 18:
            b = c // 6
            d[c % 6] = eps
            xp[7 * b:7 * b + 7] = plus(x[7 * b:7 * b + 7], d)
        else:
            xp[21 + (c - 18)] += eps
        num[:, c] = (_td_factor(oracle, abi, xp, *args, jac=False)[0] - r0) / eps
    assert np.abs(ana - num).max() <= 2e-5 * max(np.abs(ana).max(), 1.0), np.abs(ana - num).max()
    assert np.abs(Jt).max() > 1.0        # the td column is really exercised


def test_td_factor_reduces_to_projection_factor(pkg, oracle):
    """td == td_i == td_j and TR == 0 (global shutter) leave the points unshifted: same residual and pose / depth
    Jacobians as ProjectionFactor."""
    abi, S = pkg.abi, pkg.synth
    w = S.make_window(seed=5, K=6, L=8, td_true=0.0)
    k_i, k_j = int(w.lm_obs_offset[1]), int(w.lm_obs_offset[1]) + 1
    fi, fj = int(w.obs_frame[k_i]), int(w.obs_frame[k_j])
    pts_i, pts_j = np.array([*w.obs_xy[k_i], 1.0]), np.array([*w.obs_xy[k_j], 1.0])
    x = np.concatenate([w.para_pose[fi], w.para_pose[fj], w.para_ex_pose, [w.inv_depth[1]], [0.01]])
    r, Ji, Jj, Je, Jf, Jt = _td_factor(oracle, abi, x, pts_i, pts_j, w.obs_vel[k_i].copy(), w.obs_vel[k_j].copy(),
                                        0.01, 0.01, 100.0, 300.0, 0.0, 480.0)
    res = np.zeros(2)
    ji, jj, je, jf = np.zeros(14), np.zeros(14), np.zeros(14), np.zeros(2)
    oracle.oracle_projection_factor(abi.dptr(pts_i), abi.dptr(pts_j), abi.dptr(x[:7].copy()), abi.dptr(x[7:14].copy()),
                                    abi.dptr(x[14:21].copy()), float(x[21]), 460 / 1.5, abi.dptr(res), abi.dptr(ji),
                                    abi.dptr(jj), abi.dptr(je), abi.dptr(jf))
    assert np.array_equal(r, res) and np.array_equal(Ji.ravel(), ji) and np.array_equal(Jj.ravel(), jj)
    assert np.array_equal(Je.ravel(), je) and np.array_equal(Jf, jf)


@pytest.mark.parametrize("strategy", [0, 1])
def test_estimate_td_recovers_time_offset(pkg, oracle, strategy):
    """Noise-free observations generated 5 ms late, state started at the truth with td = 0: with ESTIMATE_TD the
    solver pulls para_Td to the true offset (from a perturbed start td is only weakly observable over 1 s)."""
    abi, S = pkg.abi, pkg.synth
    w = S.make_window(seed=6, K=11, L=120, td_true=0.005, noise=False, perturb=False)
    o = abi.default_opts(estimate_td=1, strategy=strategy, max_iters=40, function_tolerance=1e-14,
                         gradient_tolerance=1e-12, parameter_tolerance=1e-14)
    h, s = abi.WindowHandle(w), abi.Summary()
    assert oracle.oracle_optimize(C.byref(h.s), C.byref(o), C.byref(s)) == 0
    assert s.final_cost < 1e-5 * s.initial_cost
    assert abs(h.td[0] - 0.005) < 1e-5, h.td
    # without the td parameter the same data cannot be explained as well
    o0 = abi.default_opts(max_iters=40, function_tolerance=1e-14, gradient_tolerance=1e-12, parameter_tolerance=1e-14)
    h0, s0 = abi.WindowHandle(w), abi.Summary()
    assert oracle.oracle_optimize(C.byref(h0.s), C.byref(o0), C.byref(s0)) == 0
    assert s0.final_cost > 5 * s.final_cost


def test_window_omega_prior_is_the_schur_complement(pkg, oracle):
    """oracle_window_omega_prior (Omega_PRIOR for the selector from the back end's window) against numpy on the oracle's
    own reduced matrix: S_aa - S_ab S_bb^-1 S_ba for a = (position, velocity, accelerometer bias) of the newest frame;
    symmetric positive definite, and far from the reference's I9."""
    import ctypes as C
    abi, synth = pkg.abi, pkg.synth
    for seed, K, L in ((0, 11, 150), (1, 6, 40)):
        w = synth.make_window(seed=seed, K=K, L=L)
        hw, o = abi.WindowHandle(w), abi.default_opts()
        n = 15 * K
        S, g, h, b, c = np.zeros((n, n)), np.zeros(n), np.zeros(L), np.zeros(L), np.zeros(1)
        assert oracle.oracle_linearize(C.byref(hw.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(h), abi.dptr(b), abi.dptr(c)) == 0
        base = 15 * (K - 1)
        a = np.array([base, base + 1, base + 2, base + 6, base + 7, base + 8, base + 9, base + 10, base + 11])
        rest = np.array([i for i in range(n) if i not in set(a.tolist())])
        want = S[np.ix_(a, a)] - S[np.ix_(a, rest)] @ np.linalg.solve(S[np.ix_(rest, rest)], S[np.ix_(rest, a)])
        om = np.zeros(81)
        assert oracle.oracle_window_omega_prior(C.byref(hw.s), C.byref(o), abi.dptr(om)) == 0
        om = om.reshape(9, 9)
        assert np.abs(om - want).max() <= 1e-8 * np.abs(want).max()
        assert np.abs(om - om.T).max() <= 1e-9 * np.abs(om).max() and np.linalg.eigvalsh(0.5 * (om + om.T)).min() > 0
        assert np.abs(om - np.eye(9)).max() > 10.0


def test_feature_manager_order_is_a_relabelling(pkg, oracle):
    """synth.feature_manager_order sorts the landmarks by their first frame (FeatureManager's list order, stable): the
    window is the same problem -- same cost and, landmark by landmark, the same solution."""
    import ctypes as C
    abi, synth = pkg.abi, pkg.synth
    w = synth.make_window(seed=11, K=11, L=80)
    ws = synth.feature_manager_order(w)
    start, start_s = w.obs_frame[w.lm_obs_offset[:-1]], ws.obs_frame[ws.lm_obs_offset[:-1]]
    assert (np.diff(start_s) >= 0).all() and not (np.diff(start) >= 0).all()
    order = np.argsort(start, kind="stable")
    assert np.array_equal(ws.inv_depth, w.inv_depth[order]) and ws.n_factors == w.n_factors
    for l_new, l_old in enumerate(order[:10]):
        a, b = slice(ws.lm_obs_offset[l_new], ws.lm_obs_offset[l_new + 1]), slice(w.lm_obs_offset[l_old], w.lm_obs_offset[l_old + 1])
        assert np.array_equal(ws.obs_frame[a], w.obs_frame[b]) and np.array_equal(ws.obs_xy[a], w.obs_xy[b])
    o = abi.default_opts(max_iters=6)
    h1, h2, s1, s2 = abi.WindowHandle(w), abi.WindowHandle(ws), abi.Summary(), abi.Summary()
    assert oracle.oracle_optimize(C.byref(h1.s), C.byref(o), C.byref(s1)) == 0
    assert oracle.oracle_optimize(C.byref(h2.s), C.byref(o), C.byref(s2)) == 0
    assert abs(s1.final_cost - s2.final_cost) <= 1e-10 * s1.final_cost
    assert np.abs(h2.inv - h1.inv[order]).max() <= 1e-9 * np.abs(h1.inv).max()
