"""CPU suite: pins the C++ oracle of the selector path (rows a10-a15) against the numpy
restatement, the MATLAB known-answer case shipped with the reference
(support_files/scripts/createMatricesLinearImuFactor.m:18-101) and the algebraic
identities of SURVEY.md section 8c."""
import ctypes as C

import numpy as np
import pytest
import scipy.linalg

import np_ref


@pytest.fixture(scope="module")
def mods(pkg, oracle):
    return pkg.abi, pkg.synth, oracle


def test_linear_imu_matrices_matlab_known_answer(mods):
    """createMatricesLinearImuFactor.m: imuDeltaT=0.005, accVar=0.01, biasVar=1e-4, Ri=I,
    Rj=expm(skew([1 0 1])*0.01), 2 IMU samples."""
    abi, synth, orc = mods
    d, av, bv, n = 0.005, 0.01, 1e-4, 2
    Ri = np.eye(3)
    Rj = scipy.linalg.expm(np_ref.skew([1, 0, 1]) * (n * d))
    Rimu = scipy.linalg.expm(np_ref.skew([1, 0, 1]) * d)
    qi, qj = synth.rot_to_quat(Ri), synth.rot_to_quat(Rj)
    om, A, cov = np.zeros(81), np.zeros(81), np.zeros(81)
    orc.oracle_linear_imu_matrices(abi.dptr(qi), abi.dptr(qj), n, d, av, bv, abi.dptr(om), abi.dptr(A), abi.dptr(cov))
    cov, A, om = cov.reshape(9, 9), A.reshape(9, 9), om.reshape(9, 9)
    # the script's loop, h = 0..1: Rh = Ri, Ri*Rimu
    Nij = (1.5 * Ri + 0.5 * Ri @ Rimu) * d ** 2
    Mij = (Ri + Ri @ Rimu) * d
    exp_cov = np.zeros((9, 9))
    exp_cov[0:3, 0:3] = np.eye(3) * (n * 1 * 2.5 * d ** 4 * av)       # CCt_11 = 1.5^2 + 0.5^2
    exp_cov[0:3, 3:6] = np.eye(3) * 2.0 * d ** 3 * av                 # CCt_12 = 1.5 + 0.5
    exp_cov[3:6, 0:3] = exp_cov[0:3, 3:6].T
    exp_cov[3:6, 3:6] = np.eye(3) * n * d ** 2 * av
    exp_cov[6:9, 6:9] = np.eye(3) * n * bv
    exp_A = -np.eye(9)
    exp_A[0:3, 3:6] = -np.eye(3) * n * d
    exp_A[0:3, 6:9] = Nij
    exp_A[3:6, 6:9] = Mij
    assert np.allclose(cov, exp_cov, rtol=1e-13, atol=0)
    assert abs(cov[0, 0] - 3.125e-11) < 1e-24 and abs(cov[0, 3] - 2.5e-9) < 1e-22
    assert np.allclose(A, exp_A, rtol=1e-12, atol=1e-15)
    assert np.allclose(om @ exp_cov, np.eye(9), atol=1e-9)
    # the script's sanity check: cov(1:6,1:6) is positive definite
    assert np.linalg.eigvalsh(cov[:6, :6]).min() > 0


@pytest.mark.parametrize("H", [10, 13])
def test_omega_imu_vs_numpy(mods, H):
    abi, synth, orc = mods
    p = synth.make_select_problem(seed=0, N=10, H=H)
    h = abi.SelectHandle(p)
    D = 9 * (H + 1)
    om = np.zeros(D * D)
    orc.oracle_omega_imu(C.byref(h.s), abi.dptr(om))
    om = om.reshape(D, D)
    ref = np_ref.omega_imu(p)
    assert np.allclose(om, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    assert np.allclose(om, om.T, rtol=1e-12, atol=1e-3)
    # block tridiagonal
    for a in range(H + 1):
        for b in range(H + 1):
            if abs(a - b) > 1:
                assert np.all(om[9 * a:9 * a + 9, 9 * b:9 * b + 9] == 0)
    assert np.linalg.eigvalsh(ref).min() > 0


@pytest.mark.parametrize("H,C_", [(10, 150), (13, 40), (10, 0)])
def test_build_delta_vs_numpy_and_identities(mods, H, C_):
    abi, synth, orc = mods
    p = synth.make_select_problem(seed=2, N=60, H=H, C=C_)
    h = abi.SelectHandle(p)
    D, T, N = 9 * (H + 1), 3 * H, p.N
    delta, Cc = np.zeros(N * D * D), np.zeros(N * T * T)
    valid, depth = np.zeros(N, np.int32), np.zeros(N)
    orc.oracle_build_delta(C.byref(h.s), 0, abi.dptr(delta), abi.dptr(Cc), abi.iptr(valid), abi.dptr(depth))
    delta, Cc = delta.reshape(N, D, D), Cc.reshape(N, T, T)
    P = np_ref.pos_index(H)
    nvalid = 0
    for f in range(N):
        ref = np_ref.feature_C(p, p.cand_xy[f])
        assert (ref is not None) == bool(valid[f])
        if C_ == 0:
            assert depth[f] == 1.0          # findNNDepth with an empty cloud, feature_selector.cpp:444
        if ref is None:
            assert not delta[f].any()
            continue
        nvalid += 1
        assert np.allclose(Cc[f], ref, rtol=1e-9, atol=1e-11)
        # dense Delta is zero outside the position rows/cols of frames 1..H, and equals C there
        assert np.array_equal(delta[f][np.ix_(P, P)], Cc[f])
        mask = np.ones((D, D), bool)
        mask[np.ix_(P, P)] = False
        assert not delta[f][mask].any()
        assert np.allclose(Cc[f], Cc[f].T, atol=1e-13)
        ev = np.linalg.eigvalsh(Cc[f])
        assert ev.min() > -1e-10                                        # PSD
        nvis = sum(1 for k in range(H) if Cc[f][3 * k:3 * k + 3, 3 * k:3 * k + 3].any())
        assert np.sum(ev > 1e-9) == 2 * nvis - 3                        # rank 2 n_vis - 3
    assert nvalid > N // 2


def test_logdet_and_compact_identity(mods):
    abi, synth, orc = mods
    H = 10
    p = synth.make_select_problem(seed=3, N=20, H=H)
    D, T = 9 * (H + 1), 3 * H
    M = np_ref.omega_imu(p)
    Mc = np.ascontiguousarray(M).reshape(-1).copy()
    ld = orc.oracle_logdet(abi.dptr(Mc), D)
    assert abs(ld - np.linalg.slogdet(M)[1]) < 1e-9 * abs(ld)
    P = np_ref.pos_index(H)
    Sigma = np.linalg.inv(M)[np.ix_(P, P)]
    for f in range(p.N):
        Cm = np_ref.feature_C(p, p.cand_xy[f])
        if Cm is None:
            continue
        pr = p.cand_prob[f]
        A = M.copy()
        A[np.ix_(P, P)] += pr * Cm
        Ac = np.ascontiguousarray(A).reshape(-1).copy()
        full = orc.oracle_logdet(abi.dptr(Ac), D)
        compact = ld + np.linalg.slogdet(np.eye(T) + pr * Cm @ Sigma)[1]
        assert abs(full - compact) < 1e-10 * abs(full)   # 1/cov entries ~3e10 next to the unit prior
        ub = np.sum(np.log(np.diag(A)))
        assert ub >= full                                                # Hadamard bound (eq 29)


@pytest.mark.parametrize("seed,N,H,U,kappa", [(0, 60, 10, 0, 12), (1, 50, 13, 5, 10), (2, 8, 10, 3, 20)])
def test_select_vs_numpy_greedy(mods, seed, N, H, U, kappa):
    abi, synth, orc = mods
    p = synth.make_select_problem(seed=seed, N=N, H=H, U=U, kappa=kappa)
    h = abi.SelectHandle(p)
    ids, vals = np.zeros(kappa, np.int32), np.zeros(kappa)
    s = abi.SelectSummary()
    assert orc.oracle_select(C.byref(h.s), abi.iptr(ids), abi.dptr(vals), C.byref(s)) == 0
    rids, rvals, margins = np_ref.greedy_select(p)
    assert s.n_selected == len(rids)
    assert ids[:s.n_selected].tolist() == rids                       # the lazy UB break never changes the arg-max
    assert np.allclose(vals[:s.n_selected], rvals, rtol=1e-11)
    assert min(margins) > 1e-8                                       # decisions are not roundoff coin flips
    assert s.n_selected <= min(kappa, s.n_candidates_valid)
    assert len(set(ids[:s.n_selected].tolist())) == s.n_selected
    # lazy evaluation must have skipped work when there is a choice
    if N > 20:
        assert s.candidates_scored < kappa * N


def test_select_degenerate_inputs(mods):
    abi, synth, orc = mods
    # kappa = 0 and N = 0 select nothing
    for N, kappa in ((0, 5), (10, 0)):
        p = synth.make_select_problem(seed=5, N=max(N, 1), H=10, kappa=kappa)
        if N == 0:
            p.cand_id, p.cand_xy, p.cand_prob = p.cand_id[:0], p.cand_xy[:0], p.cand_prob[:0]
        h = abi.SelectHandle(p)
        ids = np.zeros(max(kappa, 1), np.int32)
        s = abi.SelectSummary()
        assert orc.oracle_select(C.byref(h.s), abi.iptr(ids), None, C.byref(s)) == 0
        assert s.n_selected == 0
