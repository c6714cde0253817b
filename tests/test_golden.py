"""Golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle on
stored inputs): the oracle must keep reproducing them (CPU suite) and the CUDA path must match them
(GPU suite), so drift on either side is caught without re-deriving one from the other."""
import ctypes as C
import glob
import os

import numpy as np
import pytest

import golden_io

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BA = sorted(glob.glob(os.path.join(GOLD, "ba_*.npz")))
SEL = sorted(glob.glob(os.path.join(GOLD, "sel_*.npz")))
TIGHT = dict(max_iters=50, function_tolerance=1e-14, gradient_tolerance=1e-12, parameter_tolerance=1e-14)


def test_fixtures_present():
    assert len(BA) == 3 and len(SEL) == 3


@pytest.mark.parametrize("path", BA, ids=os.path.basename)
def test_oracle_reproduces_ba_golden(pkg, oracle, path):
    abi = pkg.abi
    d = np.load(path)
    w = golden_io.window_from_dict(d)
    np_ = 15 * w.K
    h = abi.WindowHandle(w)
    S, g, hh, bb, c = np.zeros((np_, np_)), np.zeros(np_), np.zeros(w.L), np.zeros(w.L), np.zeros(1)
    o = abi.default_opts()
    assert oracle.oracle_linearize(C.byref(h.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(hh), abi.dptr(bb), abi.dptr(c)) == 0
    assert np.allclose(S, d["out_S"], rtol=0, atol=1e-12 * np.abs(d["out_S"]).max())
    assert np.allclose(g, d["out_g"], rtol=0, atol=1e-12 * np.abs(d["out_g"]).max())
    assert abs(c[0] - d["out_cost"][0]) <= 1e-13 * d["out_cost"][0]
    h8, s8 = abi.WindowHandle(w), abi.Summary()
    assert oracle.oracle_optimize(C.byref(h8.s), C.byref(o), C.byref(s8)) == 0
    assert [s8.iterations, s8.num_accepted, s8.num_rejected, s8.termination] == d["out8_summary"].tolist()
    assert np.linalg.norm(h8.state_vector() - d["out8_state"]) <= 1e-10 * np.linalg.norm(d["out8_state"])


VARIANTS = [(s, e, t) for s in (0, 1) for e in (0, 1) for t in (0, 1)]


def _variant(pkg, solve, strategy, ex, td):
    abi = pkg.abi
    d = np.load(os.path.join(GOLD, "opt_variants_k11_l80.npz"))
    w = golden_io.window_from_dict(d)
    o = abi.default_opts(strategy=strategy, estimate_extrinsic=ex, estimate_td=td, TR=0.01)
    h, s = abi.WindowHandle(w), abi.Summary()
    solve(h, o, s)
    key = f"out_s{strategy}_e{ex}_t{td}"
    return d, key, np.concatenate([h.state_vector(), h.td]), s


@pytest.mark.parametrize("strategy,ex,td", VARIANTS)
def test_oracle_reproduces_variant_golden(pkg, oracle, strategy, ex, td):
    """{LM, DOGLEG} x estimate_extrinsic x estimate_td at the reference budget (8 iterations)."""
    def solve(h, o, s):
        assert oracle.oracle_optimize(C.byref(h.s), C.byref(o), C.byref(s)) == 0
    d, key, x, s = _variant(pkg, solve, strategy, ex, td)
    assert [s.iterations, s.num_accepted, s.num_rejected, s.termination] == d[key + "_summary"].tolist()
    assert np.linalg.norm(x - d[key + "_state"]) <= 1e-10 * np.linalg.norm(d[key + "_state"])


@pytest.mark.gpu
@pytest.mark.parametrize("strategy,ex,td", VARIANTS)
def test_cuda_matches_variant_golden(pkg, strategy, ex, td):
    ctx = pkg.lib.Context(0)

    def solve(h, o, s):
        ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(o), C.byref(s)), "optimize")
    d, key, x, s = _variant(pkg, solve, strategy, ex, td)
    assert [s.iterations, s.num_accepted, s.num_rejected, s.termination] == d[key + "_summary"].tolist()
    assert np.linalg.norm(x - d[key + "_state"]) <= 1e-8 * np.linalg.norm(d[key + "_state"])
    assert abs(s.final_cost - d[key + "_cost"][1]) <= 1e-9 * d[key + "_cost"][1]
    assert abs(s.final_radius - d[key + "_cost"][2]) <= 1e-6 * d[key + "_cost"][2]
    ctx.close()


@pytest.mark.parametrize("path", SEL, ids=os.path.basename)
def test_oracle_reproduces_select_golden(pkg, oracle, path):
    abi = pkg.abi
    d = np.load(path)
    p = golden_io.select_from_dict(d)
    h = abi.SelectHandle(p)
    ids, vals, s = np.full(p.kappa, -1, np.int32), np.zeros(p.kappa), abi.SelectSummary()
    assert oracle.oracle_select(C.byref(h.s), abi.iptr(ids), abi.dptr(vals), C.byref(s)) == 0
    assert ids.tolist() == d["out_ids"].tolist()
    assert np.allclose(vals, d["out_vals"], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("path", BA, ids=os.path.basename)
def test_cuda_matches_ba_golden(pkg, path):
    abi = pkg.abi
    ctx = pkg.lib.Context(0)
    d = np.load(path)
    w = golden_io.window_from_dict(d)
    np_ = 15 * w.K
    h = abi.WindowHandle(w)
    S, g, hh, bb, c = np.zeros((np_, np_)), np.zeros(np_), np.zeros(w.L), np.zeros(w.L), np.zeros(1)
    o = abi.default_opts()
    ctx.check(ctx.L.bvio_debug_linearize(ctx.h, C.byref(h.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(hh),
                                         abi.dptr(bb), abi.dptr(c)), "debug_linearize")
    assert np.abs(S - d["out_S"]).max() <= 1e-9 * np.abs(d["out_S"]).max()
    assert np.abs(g - d["out_g"]).max() <= 1e-9 * np.abs(d["out_g"]).max()
    assert np.abs(hh - d["out_h"]).max() <= 1e-11 * np.abs(d["out_h"]).max()
    assert abs(c[0] - d["out_cost"][0]) <= 1e-11 * d["out_cost"][0]
    # reference budget: same trajectory, state to 1e-8
    h8, s8 = abi.WindowHandle(w), abi.Summary()
    ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h8.s), C.byref(o), C.byref(s8)), "optimize")
    assert [s8.iterations, s8.num_accepted, s8.num_rejected, s8.termination] == d["out8_summary"].tolist()
    assert np.linalg.norm(h8.state_vector() - d["out8_state"]) <= 1e-8 * np.linalg.norm(d["out8_state"])
    # converged: the north-star bar, 1e-6 relative on the final state vector
    hc, sc = abi.WindowHandle(w), abi.Summary()
    ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(hc.s), C.byref(abi.default_opts(**TIGHT)), C.byref(sc)), "optimize")
    xg, xo = hc.state_vector(), d["outc_state"].copy()
    if not int(d["in_has_prior"][0]):
        # without a prior the problem has a 4-dof gauge (x, y, z, yaw) that only the LM damping pins:
        # compare where the reference does, after double2vector()'s re-anchoring (estimator.cpp:521-555)
        import oracle_lib
        orc = oracle_lib.load()
        K, pre0 = w.K, np.ascontiguousarray(d["in_para_pose"][0]).copy()
        for x in (xg, xo):
            pose, sb = x[:7 * K].copy(), x[7 * K:16 * K].copy()
            orc.oracle_double2vector(abi.dptr(pre0), K, abi.dptr(pose), abi.dptr(sb))
            x[:7 * K], x[7 * K:16 * K] = pose, sb
    # With a prior the fixed point is unique: north-star bar 1e-6.  Without one, metric scale is only
    # weakly observable over a 1 s window (a nearly flat valley that 50 LM iterations do not finish
    # descending), so the end points agree in cost but only to ~1e-4 along that direction.
    tol = 1e-6 if int(d["in_has_prior"][0]) else 1e-3
    assert np.linalg.norm(xg - xo) <= tol * np.linalg.norm(xo)
    assert abs(sc.final_cost - d["outc_cost"][0]) <= (1e-8 if tol == 1e-6 else 1e-6) * d["outc_cost"][0]
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", SEL, ids=os.path.basename)
def test_cuda_matches_select_golden(pkg, path):
    abi = pkg.abi
    ctx = pkg.lib.Context(0)
    d = np.load(path)
    p = golden_io.select_from_dict(d)
    h = abi.SelectHandle(p)
    T, D = 3 * p.H, 9 * (p.H + 1)
    Cg, vg, Og = np.zeros((p.N, T, T)), np.zeros(p.N, np.int32), np.zeros((D, D))
    ctx.check(ctx.L.bvio_debug_build_delta(ctx.h, C.byref(h.s), abi.dptr(Cg), abi.iptr(vg), abi.dptr(Og)), "build_delta")
    assert (vg == d["out_valid"]).all()
    assert np.abs(Cg - d["out_C"]).max() <= 1e-12 * max(np.abs(d["out_C"]).max(), 1.0)
    assert np.abs(Og - d["out_omega"]).max() <= 1e-11 * np.abs(d["out_omega"]).max()
    ids, vals, s = np.full(p.kappa, -1, np.int32), np.zeros(p.kappa), abi.SelectSummary()
    ctx.check(ctx.L.bvio_select(ctx.h, C.byref(h.s), abi.iptr(ids), abi.dptr(vals), C.byref(s)), "select")
    assert ids.tolist() == d["out_ids"].tolist()           # bit-exact index set, in selection order
    assert np.allclose(vals, d["out_vals"], rtol=1e-9)
    assert [s.n_selected, s.n_candidates_valid] == d["out_summary"].tolist()
    assert abs(s.final_logdet - d["out_final_logdet"][0]) <= 1e-9 * abs(d["out_final_logdet"][0])
    ctx.close()
