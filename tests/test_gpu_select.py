"""GPU parity suite for the selector path: libbvio.so (through the C-ABI) against the CPU oracle's
literal lazy-greedy restatement.  Bar (BASELINE.json north_star): bit-exact selected index set, in
selection order; log-det values to 1e-9 relative (different factorization: T x T compact Cholesky on
the device, dense D x D LLT in the oracle/reference)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env(pkg, oracle):
    ctx = pkg.lib.Context(0)
    yield pkg.abi, pkg.synth, oracle, ctx
    ctx.close()


@pytest.mark.parametrize("seed,N,H,U", [(0, 64, 10, 0), (1, 40, 13, 6), (2, 33, 4, 2), (3, 17, 16, 0)])
def test_build_delta_and_omega_match_oracle(env, seed, N, H, U):
    abi, synth, orc, ctx = env
    p = synth.make_select_problem(seed=seed, N=N, H=H, U=U, kappa=5)
    h = abi.SelectHandle(p)
    T, D = 3 * H, 9 * (H + 1)
    Cg, vg, Og = np.zeros((N, T, T)), np.zeros(N, np.int32), np.zeros((D, D))
    ctx.check(ctx.L.bvio_debug_build_delta(ctx.h, C.byref(h.s), abi.dptr(Cg), abi.iptr(vg), abi.dptr(Og)), "build_delta")
    Co, vo, Oo = np.zeros((N, T, T)), np.zeros(N, np.int32), np.zeros((D, D))
    orc.oracle_build_delta(C.byref(h.s), 0, None, abi.dptr(Co), abi.iptr(vo), None)
    orc.oracle_omega_imu(C.byref(h.s), abi.dptr(Oo))
    assert (vg == vo).all()
    assert vo.sum() > 0
    assert np.abs(Cg - Co).max() <= 1e-12 * max(np.abs(Co).max(), 1.0)
    assert np.abs(Og - Oo).max() <= 1e-11 * np.abs(Oo).max()


def _select_both(env, p):
    abi, synth, orc, ctx = env
    h = abi.SelectHandle(p)
    k = max(p.kappa, 1)
    ig, io = np.full(k, -1, np.int32), np.full(k, -1, np.int32)
    vg, vo = np.zeros(k), np.zeros(k)
    sg, so = abi.SelectSummary(), abi.SelectSummary()
    ctx.check(ctx.L.bvio_select(ctx.h, C.byref(h.s), abi.iptr(ig), abi.dptr(vg), C.byref(sg)), "bvio_select")
    assert orc.oracle_select(C.byref(h.s), abi.iptr(io), abi.dptr(vo), C.byref(so)) == 0
    return ig, vg, sg, io, vo, so


@pytest.mark.parametrize("seed,N,H,U,kappa", [(0, 60, 10, 0, 12), (1, 50, 13, 5, 10), (2, 8, 10, 3, 20),
                                              (3, 300, 10, 0, 40), (4, 120, 16, 4, 25), (5, 200, 5, 0, 30)])
def test_select_bit_exact_index_set(env, seed, N, H, U, kappa):
    ig, vg, sg, io, vo, so = _select_both(env, env[1].make_select_problem(seed=seed, N=N, H=H, U=U, kappa=kappa))
    assert sg.n_selected == so.n_selected
    n = so.n_selected
    assert ig[:n].tolist() == io[:n].tolist(), (sg.as_dict(), so.as_dict())
    assert np.allclose(vg[:n], vo[:n], rtol=1e-9, atol=0)
    assert sg.n_candidates_valid == so.n_candidates_valid
    assert abs(sg.final_logdet - so.final_logdet) <= 1e-9 * abs(so.final_logdet)
    # the device scores every remaining candidate each round; the reference's lazy loop scores fewer
    assert sg.candidates_scored >= so.candidates_scored
    assert so.min_margin > 1e-9, "oracle decision margin too small for a meaningful bit-exact comparison"


def test_config4_2000_candidates_kappa150(env):
    """BASELINE config 4 on one GPU: 2000 candidates, H = 10, kappa = 150."""
    ig, vg, sg, io, vo, so = _select_both(env, env[1].make_select_problem(seed=0, N=2000, H=10, kappa=150))
    n = so.n_selected
    assert sg.n_selected == n == 150
    assert ig[:n].tolist() == io[:n].tolist()
    assert np.allclose(vg[:n], vo[:n], rtol=1e-9, atol=0)
    nv = sg.n_candidates_valid
    assert sg.candidates_scored == sum(nv - i for i in range(150))


def test_select_degenerate_inputs(env):
    abi, synth, orc, ctx = env
    for N, kappa in ((0, 5), (10, 0)):
        p = synth.make_select_problem(seed=5, N=max(N, 1), H=10, kappa=kappa)
        if N == 0:
            p.cand_id, p.cand_xy, p.cand_prob = p.cand_id[:0], p.cand_xy[:0], p.cand_prob[:0]
        ig, vg, sg, io, vo, so = _select_both(env, p)
        assert sg.n_selected == so.n_selected == 0
    # empty depth cloud: findNNDepth returns 1.0 (feature_selector.cpp:444)
    p = synth.make_select_problem(seed=6, N=40, H=10, C=0, kappa=8)
    ig, vg, sg, io, vo, so = _select_both(env, p)
    assert ig[:so.n_selected].tolist() == io[:so.n_selected].tolist()
    # kappa larger than the number of valid candidates: every valid candidate ends up selected
    p = synth.make_select_problem(seed=7, N=12, H=10, kappa=30)
    ig, vg, sg, io, vo, so = _select_both(env, p)
    assert sg.n_selected == so.n_selected <= 12
    assert ig[:so.n_selected].tolist() == io[:so.n_selected].tolist()


def test_resident_selector_is_deterministic(env):
    abi, synth, orc, ctx = env
    p = synth.make_select_problem(seed=9, N=500, H=10, kappa=50)
    h = abi.SelectHandle(p)
    ph = C.c_void_p()
    ctx.check(ctx.L.bvio_select_upload(ctx.h, C.byref(h.s), C.byref(ph)), "upload")
    runs = []
    for _ in range(3):
        ctx.check(ctx.L.bvio_select_run(ctx.h, ph), "run")
        ids, vals, s = np.zeros(50, np.int32), np.zeros(50), abi.SelectSummary()
        ctx.check(ctx.L.bvio_select_fetch(ctx.h, ph, abi.iptr(ids), abi.dptr(vals), C.byref(s)), "fetch")
        runs.append((ids.copy(), vals.copy(), s.candidates_scored))
    ctx.L.bvio_select_free(ctx.h, ph)
    for r in runs[1:]:
        assert np.array_equal(r[0], runs[0][0]) and np.array_equal(r[1], runs[0][1]) and r[2] == runs[0][2]


@pytest.mark.parametrize("seed,N,H,U,kappa", [(0, 80, 10, 0, 15), (1, 120, 13, 6, 25)])
def test_select_with_separate_state_k1_and_omega_prior(env, seed, N, H, U, kappa):
    """ABI v2 inputs: candidates back-projected from an explicit state_k1 (feature_selector.cpp:247-266) and a caller
    supplied Omega_PRIOR instead of I9 (:602-609) -- device vs oracle: information blocks, Omega, ids."""
    abi, synth, orc, ctx = env
    rng = np.random.default_rng(50 + seed)
    p = synth.make_select_problem(seed=seed, N=N, H=H, U=U, kappa=kappa)
    dq = np.concatenate([0.5 * rng.normal(0, 0.02, 3), [1.0]])
    p.state_k1_pos = p.horizon_pos[1] + rng.normal(0, 0.05, 3)
    q = synth.quat_mul(p.horizon_quat[1], dq)
    p.state_k1_quat = q / np.linalg.norm(q)
    A = rng.normal(0, 1.0, (9, 9))
    p.omega_prior = A @ A.T + np.diag(rng.uniform(0.5, 50.0, 9))
    h = abi.SelectHandle(p)
    T, D = 3 * H, 9 * (H + 1)
    Cg, vg, Og = np.zeros((N, T, T)), np.zeros(N, np.int32), np.zeros((D, D))
    ctx.check(ctx.L.bvio_debug_build_delta(ctx.h, C.byref(h.s), abi.dptr(Cg), abi.iptr(vg), abi.dptr(Og)), "build_delta")
    Co, vo, Oo = np.zeros((N, T, T)), np.zeros(N, np.int32), np.zeros((D, D))
    orc.oracle_build_delta(C.byref(h.s), 0, None, abi.dptr(Co), abi.iptr(vo), None)
    orc.oracle_omega_imu(C.byref(h.s), abi.dptr(Oo))
    assert (vg == vo).all() and vo.sum() > 0
    assert np.abs(Cg - Co).max() <= 1e-12 * max(np.abs(Co).max(), 1.0)
    assert np.abs(Og - Oo).max() <= 1e-11 * np.abs(Oo).max()
    assert np.abs(Oo[:9, :9] - p.omega_prior).max() > 0         # the prior block really went in (plus the IMU term)
    ig, vg2, sg, io, vo2, so = _select_both(env, p)
    n = so.n_selected
    assert sg.n_selected == n > 0 and ig[:n].tolist() == io[:n].tolist()
    assert np.allclose(vg2[:n], vo2[:n], rtol=1e-9, atol=0)
    # and the optional inputs matter: without them the information blocks differ
    p2 = synth.make_select_problem(seed=seed, N=N, H=H, U=U, kappa=kappa)
    h2 = abi.SelectHandle(p2)
    C2 = np.zeros((N, T, T))
    orc.oracle_build_delta(C.byref(h2.s), 0, None, abi.dptr(C2), abi.iptr(vo), None)
    assert np.abs(C2 - Co).max() > 1e-6


def test_select_exact_duplicates_name_the_larger_id_first(env):
    """Tie rule of the device = the reference's UB-map collision (feature_selector.cpp:697,724): device vs oracle on
    inputs with bit-identical twins, also when the twins sit in different warps / CTAs of the persistent kernel."""
    abi, synth, orc, ctx = env
    for seed, N, kappa, pairs in ((0, 64, 20, [(0, 1), (10, 11), (30, 63)]), (1, 600, 40, [(5, 599), (100, 101), (7, 300), (8, 301)])):
        p = synth.make_select_problem(seed=seed, N=N, H=10, kappa=kappa)
        for a, b in pairs:
            p.cand_xy[b], p.cand_prob[b] = p.cand_xy[a], p.cand_prob[a]
        ig, vg, sg, io, vo, so = _select_both(env, p)
        n = so.n_selected
        assert sg.n_selected == n and ig[:n].tolist() == io[:n].tolist(), (ig[:n], io[:n])
        order = {int(i): k for k, i in enumerate(ig[:n])}
        both = [(int(p.cand_id[a]), int(p.cand_id[b])) for a, b in pairs if int(p.cand_id[a]) in order and int(p.cand_id[b]) in order]
        assert all(order[b] < order[a] for a, b in both)


@pytest.mark.parametrize("seed,K,L,relo", [(0, 11, 150, False), (1, 11, 1500, False), (2, 6, 40, False), (3, 11, 80, True)])
def test_window_omega_prior_matches_oracle(env, seed, K, L, relo):
    """bvio_window_omega_prior (the window's information on position / velocity / accelerometer bias of its newest frame,
    the opt-in Omega_PRIOR of the selector) against the oracle's dense Schur complement."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=seed, K=K, L=L)
    if relo:
        w = synth.add_relocalization(w, seed, local_index=4)
    hw, o = abi.WindowHandle(w), abi.default_opts()
    og, oo = np.zeros(81), np.zeros(81)
    ctx.check(ctx.L.bvio_window_omega_prior(ctx.h, C.byref(hw.s), C.byref(o), abi.dptr(og)), "bvio_window_omega_prior")
    assert orc.oracle_window_omega_prior(C.byref(hw.s), C.byref(o), abi.dptr(oo)) == 0
    assert np.abs(og - oo).max() <= 1e-8 * np.abs(oo).max()
    # and it is usable as bvio_select_in.omega_prior: the selection runs and differs from the I9 one in its log-dets
    p = synth.make_select_problem(seed=seed, N=150, H=10, kappa=15)
    ig, vg, sg, io, vo, so = _select_both(env, p)
    p.omega_prior = og.reshape(9, 9)
    ig2, vg2, sg2, io2, vo2, so2 = _select_both(env, p)
    assert sg2.n_selected == so2.n_selected and ig2[:so2.n_selected].tolist() == io2[:so2.n_selected].tolist()
    assert abs(sg2.final_logdet - sg.final_logdet) > 1.0
