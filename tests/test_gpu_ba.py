"""GPU parity suite for the BA path: libbvio.so (through the C-ABI) against the CPU oracle on the
same seeded windows.  Tolerances: 1e-6 relative on the final state vector (BASELINE.json
north_star); linearization products to 1e-9 of their max-norm (different summation order)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TIGHT = dict(max_iters=50, function_tolerance=1e-14, gradient_tolerance=1e-12, parameter_tolerance=1e-14)


@pytest.fixture(scope="module")
def env(pkg, oracle):
    ctx = pkg.lib.Context(0)
    yield pkg.abi, pkg.synth, oracle, ctx
    ctx.close()


def _perturb_extrinsic(synth, w, seed, sigma_t=0.02, sigma_r=0.01):
    """Window copy whose camera-IMU extrinsics are off by (2 cm, 0.6 deg): something for ESTIMATE_EXTRINSIC to fix."""
    rng = np.random.default_rng(1000 + seed)
    w = w.copy()
    w.para_ex_pose[:3] += rng.normal(0, sigma_t, 3)
    q = synth.quat_mul(w.para_ex_pose[3:], np.concatenate([0.5 * rng.normal(0, sigma_r, 3), [1.0]]))
    w.para_ex_pose[3:] = q / np.linalg.norm(q)
    return w


def _linearize_both(env, w, **opts_kw):
    abi, synth, orc, ctx = env
    o = abi.default_opts(**opts_kw)
    h1, h2 = abi.WindowHandle(w), abi.WindowHandle(w)
    np_ = 15 * w.K + (6 if opts_kw.get("estimate_extrinsic") else 0) + (1 if opts_kw.get("estimate_td") else 0)
    L = w.L
    out = {}
    for name, fn, hh in (("gpu", None, h1), ("cpu", orc.oracle_linearize, h2)):
        S, g, h, b, c = np.zeros((np_, np_)), np.zeros(np_), np.zeros(L), np.zeros(L), np.zeros(1)
        if name == "gpu":
            ctx.check(ctx.L.bvio_debug_linearize(ctx.h, C.byref(hh.s), C.byref(o), abi.dptr(S), abi.dptr(g),
                                                 abi.dptr(h), abi.dptr(b), abi.dptr(c)), "debug_linearize")
        else:
            assert fn(C.byref(hh.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(h), abi.dptr(b), abi.dptr(c)) == 0
        out[name] = (S, g, h, b, c[0])
    return out


@pytest.mark.parametrize("seed,K,L,prior", [(0, 11, 150, "frame0"), (1, 11, 150, "none"), (2, 2, 20, "none"),
                                            (3, 5, 37, "frame0"), (4, 11, 1500, "frame0"), (5, 15, 64, "frame0")])
def test_linearize_matches_oracle(env, seed, K, L, prior):
    abi, synth, orc, ctx = env
    kw = dict(track_min=2, track_max=2) if K == 2 else {}
    w = synth.make_window(seed=seed, K=K, L=L, prior=prior, **kw)
    r = _linearize_both(env, w)
    (S1, g1, h1, b1, c1), (S2, g2, h2, b2, c2) = r["gpu"], r["cpu"]
    assert np.isfinite(S1).all()
    assert abs(c1 - c2) <= 1e-11 * abs(c2)
    assert np.abs(h1 - h2).max() <= 1e-11 * np.abs(h2).max()
    assert np.abs(b1 - b2).max() <= 1e-10 * max(np.abs(b2).max(), 1.0)
    # S entries are differences of nearly equal sums (gauge directions): scale by the unreduced diagonal
    assert np.abs(S1 - S2).max() <= 1e-9 * np.abs(S2).max()
    assert np.abs(g1 - g2).max() <= 1e-9 * max(np.abs(g2).max(), 1.0)
    assert np.abs(S1 - S1.T).max() == 0.0


def _solve_both(env, w, opts_kw):
    abi, synth, orc, ctx = env
    o = abi.default_opts(**opts_kw)
    hg, ho = abi.WindowHandle(w), abi.WindowHandle(w)
    sg, so = abi.Summary(), abi.Summary()
    ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(hg.s), C.byref(o), C.byref(sg)), "bvio_optimize")
    assert orc.oracle_optimize(C.byref(ho.s), C.byref(o), C.byref(so)) == 0
    return hg, ho, sg, so


@pytest.mark.parametrize("seed", range(6))
def test_config2_converged_state_matches_oracle(env, seed):
    """BASELINE config 2: 11-kf / 150-feature window, LM to convergence."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=seed, K=11, L=150)
    hg, ho, sg, so = _solve_both(env, w, TIGHT)
    xg, xo = hg.state_vector(), ho.state_vector()
    rel = np.linalg.norm(xg - xo) / np.linalg.norm(xo)
    assert rel < 1e-6, (rel, sg.as_dict(), so.as_dict())
    # per-block check as well: poses, speed/bias, inverse depths each to 1e-6 of their own norm
    for a, b in ((hg.pose, ho.pose), (hg.sb, ho.sb), (hg.inv, ho.inv)):
        assert np.linalg.norm(a - b) <= 1e-6 * np.linalg.norm(b)
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    assert abs(sg.initial_cost - so.initial_cost) <= 1e-10 * so.initial_cost
    assert sg.final_cost < sg.initial_cost


def test_config1_one_iteration_plumbing(env):
    """BASELINE config 1: 2-keyframe, 20-feature window, exactly one GN/LM iteration."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=1, K=2, L=20, track_min=2, track_max=2, prior="frame0")
    hg, ho, sg, so = _solve_both(env, w, dict(max_iters=1))
    assert sg.iterations == so.iterations == 1
    assert sg.num_accepted == so.num_accepted
    xg, xo = hg.state_vector(), ho.state_vector()
    assert np.linalg.norm(xg - xo) <= 1e-9 * np.linalg.norm(xo)
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * abs(so.final_cost)


def test_default_opts_trajectory_matches_oracle(env):
    """Reference budget (8 iterations, Ceres default tolerances): same iteration count / termination."""
    abi, synth, orc, ctx = env
    for seed in range(3):
        w = synth.make_window(seed=10 + seed, K=11, L=150)
        hg, ho, sg, so = _solve_both(env, w, {})
        assert (sg.iterations, sg.num_accepted, sg.num_rejected, sg.termination) == \
               (so.iterations, so.num_accepted, so.num_rejected, so.termination), (sg.as_dict(), so.as_dict())
        xg, xo = hg.state_vector(), ho.state_vector()
        assert np.linalg.norm(xg - xo) <= 1e-8 * np.linalg.norm(xo)


@pytest.mark.parametrize("seed,radius,iters", [(10, 1e4, 8), (11, 1e4, 8), (12, 1e4, 8), (13, 1.0, 8), (14, 0.05, 8),
                                               (15, 1e-3, 8), (16, 1.0, 25), (17, 0.05, 30)])
def test_dogleg_reference_budget_matches_oracle(env, seed, radius, iters):
    """The strategy the reference configures (estimator.cpp:800 DOGLEG, 8 iterations, Ceres default
    tolerances): identical iteration / accept / reject / termination trajectory and the same state.  Small
    initial radii force the Cauchy-point and interpolated branches of the traditional dogleg."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=seed, K=11, L=150)
    hg, ho, sg, so = _solve_both(env, w, dict(strategy=1, initial_radius=radius, max_iters=iters))
    assert (sg.iterations, sg.num_accepted, sg.num_rejected, sg.termination) == \
           (so.iterations, so.num_accepted, so.num_rejected, so.termination), (sg.as_dict(), so.as_dict())
    xg, xo = hg.state_vector(), ho.state_vector()
    assert np.linalg.norm(xg - xo) <= 1e-8 * np.linalg.norm(xo), (sg.as_dict(), so.as_dict())
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    assert abs(sg.final_radius - so.final_radius) <= 1e-6 * so.final_radius


@pytest.mark.parametrize("seed,K,L", [(0, 11, 150), (1, 11, 150), (2, 2, 20), (3, 5, 37), (7, 11, 1500)])
def test_dogleg_converged_state_matches_oracle(env, seed, K, L):
    abi, synth, orc, ctx = env
    kw = dict(track_min=2, track_max=2) if K == 2 else (dict(track_min=6) if L == 1500 else {})
    w = synth.make_window(seed=seed, K=K, L=L, **kw)
    hg, ho, sg, so = _solve_both(env, w, dict(strategy=1, **TIGHT))
    xg, xo = hg.state_vector(), ho.state_vector()
    assert np.linalg.norm(xg - xo) <= 1e-6 * np.linalg.norm(xo), (sg.as_dict(), so.as_dict())
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    # dogleg and LM descend to the same minimum (the state along weakly observable directions is only
    # determined to ~sqrt of the cost tolerance, so compare costs)
    hl, _, sl, _ = _solve_both(env, w, TIGHT)
    assert abs(sg.final_cost - sl.final_cost) <= 1e-6 * sl.final_cost


@pytest.mark.parametrize("seed,K,L,prior", [(0, 11, 150, "frame0"), (1, 11, 150, "none"), (2, 2, 20, "none"),
                                            (3, 5, 37, "frame0"), (4, 11, 1500, "frame0"), (5, 14, 64, "frame0")])
def test_extrinsic_linearize_matches_oracle(env, seed, K, L, prior):
    """ESTIMATE_EXTRINSIC (estimator.cpp:672-683): para_Ex_Pose is a free block; the reduced system carries the
    extrinsic Jacobian of every ProjectionFactor (projection_factor.cpp:97-106) in its last 6 rows / columns."""
    abi, synth, orc, ctx = env
    kw = dict(track_min=2, track_max=2) if K == 2 else {}
    w = _perturb_extrinsic(synth, synth.make_window(seed=seed, K=K, L=L, prior=prior, **kw), seed)
    r = _linearize_both(env, w, estimate_extrinsic=1)
    (S1, g1, h1, b1, c1), (S2, g2, h2, b2, c2) = r["gpu"], r["cpu"]
    assert np.isfinite(S1).all() and S1.shape == (15 * K + 6, 15 * K + 6)
    assert np.abs(S2[15 * K:, :]).max() > 0
    assert abs(c1 - c2) <= 1e-11 * abs(c2)
    assert np.abs(h1 - h2).max() <= 1e-11 * np.abs(h2).max()
    assert np.abs(S1 - S2).max() <= 1e-9 * np.abs(S2).max()
    assert np.abs(g1 - g2).max() <= 1e-9 * max(np.abs(g2).max(), 1.0)
    assert np.abs(S1 - S1.T).max() == 0.0


@pytest.mark.parametrize("seed,strategy", [(0, 0), (1, 0), (2, 0), (0, 1), (1, 1), (2, 1)])
def test_extrinsic_converged_state_matches_oracle(env, seed, strategy):
    abi, synth, orc, ctx = env
    w0 = synth.make_window(seed=seed, K=11, L=150)
    w = _perturb_extrinsic(synth, w0, seed)
    hg, ho, sg, so = _solve_both(env, w, dict(estimate_extrinsic=1, strategy=strategy, **TIGHT))
    xg, xo = hg.state_vector(), ho.state_vector()
    assert np.linalg.norm(xg - xo) <= 1e-6 * np.linalg.norm(xo), (sg.as_dict(), so.as_dict())
    assert np.linalg.norm(hg.ex - ho.ex) <= 1e-6 * np.linalg.norm(ho.ex)
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    # the extrinsic block is really free (t_ic is only weakly observable over a 1 s window, so no claim on where it goes)
    assert np.linalg.norm(hg.ex - w.para_ex_pose) > 1e-4


@pytest.mark.parametrize("seed,strategy", [(10, 0), (11, 0), (10, 1), (11, 1)])
def test_extrinsic_reference_budget_trajectory(env, seed, strategy):
    abi, synth, orc, ctx = env
    w = _perturb_extrinsic(synth, synth.make_window(seed=seed, K=11, L=150), seed)
    hg, ho, sg, so = _solve_both(env, w, dict(estimate_extrinsic=1, strategy=strategy))
    assert (sg.iterations, sg.num_accepted, sg.num_rejected, sg.termination) == \
           (so.iterations, so.num_accepted, so.num_rejected, so.termination), (sg.as_dict(), so.as_dict())
    xg, xo = hg.state_vector(), ho.state_vector()
    assert np.linalg.norm(xg - xo) <= 1e-8 * np.linalg.norm(xo)


@pytest.mark.parametrize("seed,K,L,ex,TR", [(0, 11, 150, 0, 0.0), (1, 11, 150, 1, 0.0), (2, 2, 20, 0, 0.0),
                                            (3, 5, 37, 1, 0.02), (4, 11, 1500, 0, 0.03), (5, 14, 64, 1, 0.0)])
def test_td_linearize_matches_oracle(env, seed, K, L, ex, TR):
    """ESTIMATE_TD (estimator.cpp:732-740): ProjectionTdFactor on time-shifted points, para_Td free: the reduced
    system carries the td row / column after the frames (and after the extrinsic block when that is free too)."""
    abi, synth, orc, ctx = env
    kw = dict(track_min=2, track_max=2) if K == 2 else {}
    w = synth.make_window(seed=seed, K=K, L=L, td_true=0.004, **kw)
    w.para_td[0] = 0.001
    if ex:
        w = _perturb_extrinsic(synth, w, seed)
    n = 15 * K + 6 * ex + 1
    r = _linearize_both(env, w, estimate_td=1, estimate_extrinsic=ex, TR=TR)
    (S1, g1, h1, b1, c1), (S2, g2, h2, b2, c2) = r["gpu"], r["cpu"]
    assert np.isfinite(S1).all() and S1.shape == (n, n)
    assert np.abs(S2[n - 1, :]).max() > 0
    assert abs(c1 - c2) <= 1e-11 * abs(c2)
    assert np.abs(h1 - h2).max() <= 1e-11 * np.abs(h2).max()
    assert np.abs(S1 - S2).max() <= 1e-9 * np.abs(S2).max()
    assert np.abs(g1 - g2).max() <= 1e-9 * max(np.abs(g2).max(), 1.0)
    assert np.abs(S1 - S1.T).max() == 0.0


@pytest.mark.parametrize("seed,strategy,ex", [(0, 0, 0), (1, 0, 1), (2, 1, 0), (3, 1, 1)])
def test_td_converged_state_matches_oracle(env, seed, strategy, ex):
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=seed, K=11, L=150, td_true=0.006)
    if ex:
        w = _perturb_extrinsic(synth, w, seed)
    hg, ho, sg, so = _solve_both(env, w, dict(estimate_td=1, estimate_extrinsic=ex, strategy=strategy, TR=0.01, **TIGHT))
    xg, xo = hg.state_vector(), ho.state_vector()
    assert np.linalg.norm(xg - xo) <= 1e-6 * np.linalg.norm(xo), (sg.as_dict(), so.as_dict())
    assert abs(hg.td[0] - ho.td[0]) <= 1e-6 * max(abs(ho.td[0]), 1e-3), (hg.td, ho.td)
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost
    assert abs(hg.td[0]) > 1e-5                      # the offset really moved away from 0


@pytest.mark.parametrize("seed,strategy", [(10, 0), (11, 1)])
def test_td_reference_budget_trajectory(env, seed, strategy):
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=seed, K=11, L=150, td_true=0.006)
    hg, ho, sg, so = _solve_both(env, w, dict(estimate_td=1, strategy=strategy))
    assert (sg.iterations, sg.num_accepted, sg.num_rejected, sg.termination) == \
           (so.iterations, so.num_accepted, so.num_rejected, so.termination), (sg.as_dict(), so.as_dict())
    assert np.linalg.norm(hg.state_vector() - ho.state_vector()) <= 1e-8 * np.linalg.norm(ho.state_vector())
    assert abs(hg.td[0] - ho.td[0]) <= 1e-8


def test_td_recovers_time_offset_on_device(env):
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=6, K=11, L=120, td_true=0.005, noise=False, perturb=False)
    o = abi.default_opts(estimate_td=1, **TIGHT)
    h, s = abi.WindowHandle(w), abi.Summary()
    ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(o), C.byref(s)), "bvio_optimize")
    assert abs(h.td[0] - 0.005) < 1e-5, h.td


def test_stress_window_1500_features(env):
    """BASELINE config 3: 11-kf / 1500-feature window."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=7, K=11, L=1500, track_min=6)
    hg, ho, sg, so = _solve_both(env, w, TIGHT)
    xg, xo = hg.state_vector(), ho.state_vector()
    assert np.linalg.norm(xg - xo) <= 1e-6 * np.linalg.norm(xo)


def test_batch_equals_single(env):
    """B windows in one launch sequence give bit-identical results to B single calls (fixed-order reductions)."""
    abi, synth, orc, ctx = env
    ws = [synth.make_window(seed=20 + i, K=11, L=100 + 17 * i) for i in range(5)]
    o = abi.default_opts(max_iters=6)
    singles = []
    for w in ws:
        h = abi.WindowHandle(w)
        s = abi.Summary()
        ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(o), C.byref(s)), "optimize")
        singles.append((h.state_vector().copy(), s.final_cost))
    hs = [abi.WindowHandle(w) for w in ws]
    arr = (abi.WindowS * len(ws))(*[h.s for h in hs])
    sums = (abi.Summary * len(ws))()
    ctx.check(ctx.L.bvio_optimize_batch(ctx.h, arr, len(ws), C.byref(o), sums), "optimize_batch")
    for h, (x, c), s in zip(hs, singles, sums):
        # tile partition depends on the batch size, so reductions are reordered: not bit-equal, but tight
        assert np.linalg.norm(h.state_vector() - x) <= 1e-9 * np.linalg.norm(x)
        assert abs(s.final_cost - c) <= 1e-9 * c


def test_window_without_landmarks(env):
    """L = 0: only IMU factors and the prior (the reference's problem right after a feature drought)."""
    abi, synth, orc, ctx = env
    for strategy in (0, 1):
        w = synth.make_window(seed=1, K=6, L=0)
        hg, ho, sg, so = _solve_both(env, w, dict(strategy=strategy))
        assert (sg.iterations, sg.num_accepted, sg.termination) == (so.iterations, so.num_accepted, so.termination)
        assert np.linalg.norm(hg.state_vector() - ho.state_vector()) <= 1e-8 * np.linalg.norm(ho.state_vector())


def test_maximum_sizes(env):
    """K = 15 keyframes with every landmark tracked through all of them (15 observations; the ABI allows 16)."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=2, K=15, L=40, track_min=15, track_max=15)
    assert np.diff(w.lm_obs_offset).max() == 15
    r = _linearize_both(env, w)
    (S1, g1, h1, b1, c1), (S2, g2, h2, b2, c2) = r["gpu"], r["cpu"]
    assert np.abs(S1 - S2).max() <= 1e-9 * np.abs(S2).max() and np.abs(g1 - g2).max() <= 1e-9 * np.abs(g2).max()
    hg, ho, sg, so = _solve_both(env, w, {})
    assert (sg.iterations, sg.num_accepted, sg.termination) == (so.iterations, so.num_accepted, so.termination)
    assert np.linalg.norm(hg.state_vector() - ho.state_vector()) <= 1e-8 * np.linalg.norm(ho.state_vector())


def test_ragged_batch(env):
    """windows of one batch with very different landmark counts, including none"""
    abi, synth, orc, ctx = env
    ws = [synth.make_window(seed=60 + i, K=7, L=L) for i, L in enumerate((0, 3, 250, 1, 40, 0, 97))]
    o = abi.default_opts(max_iters=5)
    hs = [abi.WindowHandle(w) for w in ws]
    arr = (abi.WindowS * len(ws))(*[h.s for h in hs])
    sums = (abi.Summary * len(ws))()
    ctx.check(ctx.L.bvio_optimize_batch(ctx.h, arr, len(ws), C.byref(o), sums), "optimize_batch")
    for w, h, s in zip(ws, hs, sums):
        ho, so = abi.WindowHandle(w), abi.Summary()
        assert orc.oracle_optimize(C.byref(ho.s), C.byref(o), C.byref(so)) == 0
        assert (s.iterations, s.num_accepted, s.termination) == (so.iterations, so.num_accepted, so.termination)
        assert np.linalg.norm(h.state_vector() - ho.state_vector()) <= 1e-8 * np.linalg.norm(ho.state_vector())


def test_pipelined_batch_equals_single(env):
    """bvio_optimize_batch cuts batches of >= 4 windows per SM ... into pipelined sub-batches (own pinned slab, H2D on the
    copy stream, trailing D2H): every window must come back in its own slot with the result of a single call."""
    import torch
    abi, synth, orc, ctx = env
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    B = 4 * n_sm + 7                                   # two sub-batches of unequal size
    pool = [synth.make_window(seed=40 + i, K=5, L=10 + (i % 5)) for i in range(9)]
    o = abi.default_opts(max_iters=4)
    singles = []
    for w in pool:
        h, s = abi.WindowHandle(w), abi.Summary()
        ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(o), C.byref(s)), "optimize")
        singles.append((h.state_vector().copy(), s.final_cost, s.iterations))
    hs = [abi.WindowHandle(pool[(7 * i) % 9]) for i in range(B)]
    arr = (abi.WindowS * B)(*[h.s for h in hs])
    sums = (abi.Summary * B)()
    ctx.check(ctx.L.bvio_optimize_batch(ctx.h, arr, B, C.byref(o), sums), "optimize_batch")
    for i, (h, s) in enumerate(zip(hs, sums)):
        x, c, it = singles[(7 * i) % 9]
        assert s.iterations == it
        assert np.linalg.norm(h.state_vector() - x) <= 1e-9 * np.linalg.norm(x), i
        assert abs(s.final_cost - c) <= 1e-9 * c


def test_resident_batch_is_deterministic(env):
    abi, synth, orc, ctx = env
    ws = [synth.make_window(seed=30 + i, K=11, L=150) for i in range(4)]
    o = abi.default_opts()
    hs = [abi.WindowHandle(w) for w in ws]
    arr = (abi.WindowS * len(ws))(*[h.s for h in hs])
    sums = (abi.Summary * len(ws))()
    bh = C.c_void_p()
    ctx.check(ctx.L.bvio_batch_upload(ctx.h, arr, len(ws), C.byref(o), C.byref(bh)), "upload")
    res = []
    for rep in range(3):
        ctx.check(ctx.L.bvio_batch_solve(ctx.h, bh), "solve")
        ctx.check(ctx.L.bvio_batch_download(ctx.h, bh, arr, sums), "download")
        res.append(np.concatenate([h.state_vector() for h in hs]).copy())
    ctx.L.bvio_batch_free(ctx.h, bh)
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[1], res[2])
    assert ctx.L.bvio_launch_count(ctx.h) > 0


def test_rejects_unsupported_and_invalid(env):
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=0, K=4, L=10)
    h = abi.WindowHandle(w)
    s = abi.Summary()
    assert ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(abi.default_opts(estimate_td=1)), C.byref(s)) == -1   # no obs_vel
    assert ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(abi.default_opts(strategy=7)), C.byref(s)) == -1
    w15 = synth.make_window(seed=0, K=15, L=10)
    h15 = abi.WindowHandle(w15)
    assert ctx.L.bvio_optimize(ctx.h, C.byref(h15.s), C.byref(abi.default_opts(estimate_extrinsic=1)), C.byref(s)) == -1
    h.frame[1] = h.frame[0]   # not strictly ascending
    assert ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(abi.default_opts()), C.byref(s)) == -1


def test_max_solver_time_cut(env):
    """options.max_solver_time_in_seconds (estimator.cpp:799-806): a budget that is already spent after the first
    iteration ends the solve with BVIO_TERM_TIME; a generous one changes nothing (device clock, bvio.h)."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=3, K=11, L=150)
    runs = {}
    for name, kw in (("off", {}), ("generous", dict(max_time_s=10.0)), ("spent", dict(max_time_s=1e-9))):
        h, s = abi.WindowHandle(w), abi.Summary()
        ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(abi.default_opts(**kw)), C.byref(s)), "bvio_optimize")
        runs[name] = (h.state_vector(), s.as_dict())
    assert runs["off"][1]["iterations"] > 1 and runs["off"][1]["termination"] != 5
    assert np.array_equal(runs["off"][0], runs["generous"][0]) and runs["generous"][1]["termination"] == runs["off"][1]["termination"]
    assert runs["spent"][1]["iterations"] == 1 and runs["spent"][1]["termination"] == 5   # BVIO_TERM_TIME


def test_malformed_prior_is_rejected_not_dereferenced(env):
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=0, K=4, L=10, prior="frame0")
    h, s = abi.WindowHandle(w), abi.Summary()
    assert h.prior_s is not None
    h.prior_s.lin_jac = None
    h.s.prior = C.pointer(h.prior_s)
    assert ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(abi.default_opts()), C.byref(s)) == -1
    h2 = abi.WindowHandle(w)
    h2.prior_s.block_kind = None
    h2.s.prior = C.pointer(h2.prior_s)
    assert ctx.L.bvio_optimize(ctx.h, C.byref(h2.s), C.byref(abi.default_opts()), C.byref(s)) == -1


def test_two_contexts_on_two_devices(pkg, oracle):
    """The __constant__ tables and shared-memory opt-ins are per device: a second context on another GPU of the same
    process must solve and select like the first (skipped on a 1-GPU box)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    abi, synth = pkg.abi, pkg.synth
    w = synth.make_window(seed=1, K=11, L=150)
    p = synth.make_select_problem(seed=2, N=300, H=13, kappa=30)
    res = []
    ctxs = [pkg.lib.Context(0), pkg.lib.Context(1)]
    for ctx in ctxs:
        h, s = abi.WindowHandle(w), abi.Summary()
        ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(abi.default_opts()), C.byref(s)), "bvio_optimize")
        hs, ss, ids = abi.SelectHandle(p), abi.SelectSummary(), np.zeros(30, np.int32)
        ctx.check(ctx.L.bvio_select(ctx.h, C.byref(hs.s), abi.iptr(ids), None, C.byref(ss)), "bvio_select")
        res.append((h.state_vector(), s.iterations, ids.copy()))
    for ctx in ctxs:
        ctx.close()
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1] and np.array_equal(res[0][2], res[1][2])
    ho, so = abi.WindowHandle(w), abi.Summary()
    assert oracle.oracle_optimize(C.byref(ho.s), C.byref(abi.default_opts()), C.byref(so)) == 0
    assert np.linalg.norm(res[1][0] - ho.state_vector()) <= 1e-6 * np.linalg.norm(ho.state_vector())


@pytest.mark.parametrize("seed,L,strategy,ex", [(0, 80, 1, 0), (1, 150, 0, 0), (2, 60, 1, 1), (3, 1500, 1, 0)])
def test_relocalization_factors_match_oracle(env, seed, L, strategy, ex):
    """estimator.cpp:760-792 on the device (bvio_window.relo_*): reduced system at entry and the solve at the
    reference's budget against the oracle, which is pinned to the reference's own optimization() with
    relocalization_info set (tests/test_reference_pin.py::test_relocalization_factors_against_reference)."""
    abi, synth, orc, ctx = env
    K = 11
    w = synth.add_relocalization(synth.make_window(seed=seed, K=K, L=L), seed, local_index=4, max_matches=60)
    if ex:
        w = dataclasses_replace_relo(_perturb_extrinsic(synth, w, seed), w)
    assert len(w.relo_lm) >= 5
    kw = dict(strategy=strategy, estimate_extrinsic=ex)
    npar = 15 * (K + 1) + 6 * ex
    o = abi.default_opts(**kw)
    out = []
    for dev in (True, False):
        hh = abi.WindowHandle(w)
        S, g, h, b, c = np.zeros((npar, npar)), np.zeros(npar), np.zeros(L), np.zeros(L), np.zeros(1)
        if dev:
            ctx.check(ctx.L.bvio_debug_linearize(ctx.h, C.byref(hh.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(h), abi.dptr(b),
                                                 abi.dptr(c)), "debug_linearize")
        else:
            assert orc.oracle_linearize(C.byref(hh.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(h), abi.dptr(b), abi.dptr(c)) == 0
        out.append((S, g, h, b, c[0]))
    (S1, g1, h1, b1, c1), (S2, g2, h2, b2, c2) = out
    assert abs(c1 - c2) <= 1e-11 * c2
    assert np.abs(S1 - S2).max() <= 1e-9 * np.abs(S2).max() and np.abs(g1 - g2).max() <= 1e-9 * max(np.abs(g2).max(), 1.0)
    assert np.abs(h1 - h2).max() <= 1e-9 * np.abs(h2).max() and np.abs(b1 - b2).max() <= 1e-9 * np.abs(b2).max()
    assert np.abs(S2[15 * K:15 * K + 6, 15 * K:15 * K + 6]).max() > 0 and np.abs(S1[15 * K + 6:15 * K + 15]).max() == 0
    hg, ho, sg, so = abi.WindowHandle(w), abi.WindowHandle(w), abi.Summary(), abi.Summary()
    ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(hg.s), C.byref(o), C.byref(sg)), "bvio_optimize")
    assert orc.oracle_optimize(C.byref(ho.s), C.byref(o), C.byref(so)) == 0
    assert (sg.iterations, sg.num_accepted, sg.num_rejected, sg.termination) == (so.iterations, so.num_accepted, so.num_rejected, so.termination)
    assert np.linalg.norm(hg.state_vector() - ho.state_vector()) <= 1e-6 * np.linalg.norm(ho.state_vector())
    assert np.abs(hg.relo_pose - ho.relo_pose).max() <= 1e-6 and np.abs(ho.relo_pose - w.relo_pose).max() > 1e-5
    assert abs(sg.final_cost - so.final_cost) <= 1e-8 * so.final_cost


def dataclasses_replace_relo(w_new, w_relo):
    import dataclasses
    return dataclasses.replace(w_new, relo_pose=w_relo.relo_pose, relo_lm=w_relo.relo_lm, relo_xy=w_relo.relo_xy)


def test_relocalization_rejections(env):
    abi, synth, orc, ctx = env
    w = synth.add_relocalization(synth.make_window(seed=0, K=11, L=40, td_true=0.002), 0)
    h, s = abi.WindowHandle(w), abi.Summary()
    assert ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(abi.default_opts(estimate_td=1, TR=0.01)), C.byref(s)) == -4   # UNSUPPORTED
    w2 = synth.add_relocalization(synth.make_window(seed=0, K=11, L=40), 0)
    w2.relo_lm = w2.relo_lm[::-1].copy()                                  # not ascending
    h2 = abi.WindowHandle(w2)
    assert ctx.L.bvio_optimize(ctx.h, C.byref(h2.s), C.byref(abi.default_opts()), C.byref(s)) == -1
    # batches mix windows with and without matches
    wa = synth.add_relocalization(synth.make_window(seed=5, K=11, L=60), 5)
    wb = synth.make_window(seed=6, K=11, L=60)
    hs = [abi.WindowHandle(wa), abi.WindowHandle(wb)]
    arr = (abi.WindowS * 2)(hs[0].s, hs[1].s)
    sums = (abi.Summary * 2)()
    ctx.check(ctx.L.bvio_optimize_batch(ctx.h, arr, 2, C.byref(abi.default_opts()), sums), "batch")
    for wsrc, hdev in zip((wa, wb), hs):
        ho, so = abi.WindowHandle(wsrc), abi.Summary()
        assert orc.oracle_optimize(C.byref(ho.s), C.byref(abi.default_opts()), C.byref(so)) == 0
        assert np.linalg.norm(hdev.state_vector() - ho.state_vector()) <= 1e-6 * np.linalg.norm(ho.state_vector())


# ---- warp-specialised linearize kernel (throughput mode: batches of at least one window per SM) ---------------------------
@pytest.fixture
def force_ws(monkeypatch):
    """BVIO_LIN_WS=1: take ba_linearize_ws_kernel for any batch size (it is otherwise chosen when B >= the SM count)."""
    monkeypatch.setenv("BVIO_LIN_WS", "1")


@pytest.mark.parametrize("seed,K,L,prior,kw", [(0, 11, 150, "frame0", {}), (1, 11, 150, "none", {}), (2, 2, 20, "none", dict(track_min=2, track_max=2)),
                                               (3, 5, 37, "frame0", {}), (4, 11, 1500, "frame0", {}), (5, 12, 64, "frame0", {}),
                                               (6, 11, 400, "frame0", dict(track_min=2, track_max=4)),      # 40 landmarks per chunk
                                               (7, 11, 300, "none", dict(track_min=10))])                     # ~25 per chunk
def test_ws_linearize_matches_oracle(env, force_ws, seed, K, L, prior, kw):
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=seed, K=K, L=L, prior=prior, **kw)
    r = _linearize_both(env, w)
    (S1, g1, h1, b1, c1), (S2, g2, h2, b2, c2) = r["gpu"], r["cpu"]
    assert np.isfinite(S1).all()
    assert abs(c1 - c2) <= 1e-11 * abs(c2)
    assert np.abs(h1 - h2).max() <= 1e-11 * np.abs(h2).max()
    assert np.abs(b1 - b2).max() <= 1e-10 * max(np.abs(b2).max(), 1.0)
    assert np.abs(S1 - S2).max() <= 1e-9 * np.abs(S2).max()
    assert np.abs(g1 - g2).max() <= 1e-9 * max(np.abs(g2).max(), 1.0)
    assert np.abs(S1 - S1.T).max() == 0.0


@pytest.mark.parametrize("seed,L,strategy", [(0, 150, 0), (1, 150, 1), (7, 1500, 1)])
def test_ws_converged_state_matches_oracle(env, force_ws, seed, L, strategy):
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=seed, K=11, L=L, **(dict(track_min=6) if L == 1500 else {}))
    hg, ho, sg, so = _solve_both(env, w, dict(strategy=strategy, **TIGHT))
    xg, xo = hg.state_vector(), ho.state_vector()
    assert np.linalg.norm(xg - xo) <= 1e-6 * np.linalg.norm(xo), (sg.as_dict(), so.as_dict())
    assert abs(sg.final_cost - so.final_cost) <= 1e-9 * so.final_cost


def test_ws_batch_equals_single_solves(env):
    """A batch of one window per SM and more takes the warp-specialised kernel on its own (no env override): every window
    must come back with the result of its single-window solve (latency mode = ba_linearize_mma_kernel)."""
    import torch
    abi, synth, orc, ctx = env
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    B = n_sm + 3
    pool = [synth.make_window(seed=60 + i, K=11, L=120 + 40 * i, **(dict(track_min=2, track_max=5) if i == 2 else {})) for i in range(5)]
    o = abi.default_opts(max_iters=6, strategy=1)
    singles = []
    for w in pool:
        h, s = abi.WindowHandle(w), abi.Summary()
        ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(o), C.byref(s)), "optimize")
        singles.append((h.state_vector().copy(), s.final_cost, s.iterations))
    hs = [abi.WindowHandle(pool[(3 * i) % 5]) for i in range(B)]
    arr = (abi.WindowS * B)(*[h.s for h in hs])
    sums = (abi.Summary * B)()
    ctx.check(ctx.L.bvio_optimize_batch(ctx.h, arr, B, C.byref(o), sums), "optimize_batch")
    for i, (h, s) in enumerate(zip(hs, sums)):
        x, c, it = singles[(3 * i) % 5]
        assert s.iterations == it
        assert np.linalg.norm(h.state_vector() - x) <= 1e-9 * np.linalg.norm(x), i
        assert abs(s.final_cost - c) <= 1e-9 * c


def test_landmark_order_does_not_matter(env):
    """Windows in FeatureManager's list order are packed with bulk copies, any other order landmark by landmark after a
    host-side sort by anchor: same problem, same answer (per landmark), alone and in a batch that mixes both."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=12, K=11, L=200)
    ws = synth.feature_manager_order(w)
    order = np.argsort(w.obs_frame[w.lm_obs_offset[:-1]], kind="stable")
    o = abi.default_opts(max_iters=8, strategy=1)
    h1, h2, s1, s2 = abi.WindowHandle(w), abi.WindowHandle(ws), abi.Summary(), abi.Summary()
    ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h1.s), C.byref(o), C.byref(s1)), "optimize")
    ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h2.s), C.byref(o), C.byref(s2)), "optimize")
    assert s1.iterations == s2.iterations and abs(s1.final_cost - s2.final_cost) <= 1e-12 * s1.final_cost
    assert np.abs(h1.pose - h2.pose).max() <= 1e-12 and np.abs(h2.inv - h1.inv[order]).max() <= 1e-12 * np.abs(h1.inv).max()
    hs = [abi.WindowHandle(w if i % 2 else ws) for i in range(20)]
    arr = (abi.WindowS * 20)(*[h.s for h in hs])
    sums = (abi.Summary * 20)()
    ctx.check(ctx.L.bvio_optimize_batch(ctx.h, arr, 20, C.byref(o), sums), "optimize_batch")
    for i, h in enumerate(hs):
        ref = h1 if i % 2 else h2
        assert np.abs(h.pose - ref.pose).max() <= 1e-12 and np.abs(h.inv - ref.inv).max() <= 1e-12 * np.abs(ref.inv).max()
