"""GPU parity for the marginalization path (row a9): bvio_marginalize against the oracle's literal
restatement.  The rows of linearized_jacobians are only defined up to the eigenvector basis, so
parity is on the reference's own invariants J^T J (= A) and J^T r (= b)
(marginalization_factor.cpp:295-296), in state coordinates, plus block bookkeeping and x0."""
import ctypes as C
import dataclasses
import os

import numpy as np
import pytest

from test_oracle_marg import info_in_state_coords, run_marg

pytestmark = pytest.mark.gpu
KEYS = ("n", "block_kind", "block_frame", "block_idx", "x0", "lin_jac", "lin_res")


@pytest.fixture(scope="module", params=["eigen", "cholesky"])
def env(pkg, oracle, request):
    """Both factorizations of the kept information into (J, r): the reference's eigen-decomposition (default) and the
    opt-in pivoted Cholesky (BVIO_MARG_CHOLESKY=1, 2x faster); the prior they define is the same quadratic form."""
    if request.param == "cholesky":
        os.environ["BVIO_MARG_CHOLESKY"] = "1"
    ctx = pkg.lib.Context(0)
    yield pkg.abi, pkg.synth, oracle, ctx
    ctx.close()
    os.environ.pop("BVIO_MARG_CHOLESKY", None)


def _both(env, w, flag, **opts_kw):
    abi, synth, orc, ctx = env
    pg = run_marg(abi, ctx.L.bvio_marginalize, w, flag, ctx=ctx.h, opts=abi.default_opts(**opts_kw))
    po = run_marg(abi, orc.oracle_marginalize, w, flag, opts=abi.default_opts(**opts_kw))
    return pg, po


def _compare(pg, po, K, unshift, htol=1e-7):
    assert pg["n"] == po["n"]
    for k in ("block_kind", "block_frame", "block_idx"):
        assert np.array_equal(pg[k], po[k]), k
    assert np.array_equal(pg["x0"], po["x0"])
    Hg, gg = info_in_state_coords(pg, K, unshift)
    Ho, go = info_in_state_coords(po, K, unshift)
    # cond(Amm) ~ 1e8 and different elimination orders (analytic depth elimination vs one big pseudo-inverse)
    assert np.abs(Hg - Ho).max() <= htol * np.abs(Ho).max()
    assert np.abs(gg - go).max() <= 5e-5 * max(np.abs(go).max(), 1.0)
    # the factorisation itself reproduces a PSD matrix
    ev = np.linalg.eigvalsh(pg["J"].T @ pg["J"])
    assert ev.min() >= -1e-9 * ev.max()


@pytest.mark.parametrize("seed,K,L,prior", [(0, 11, 150, "frame0"), (1, 6, 40, "frame0"), (2, 11, 30, "none"), (3, 11, 400, "frame0")])
def test_margin_old_matches_oracle(env, seed, K, L, prior):
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=seed, K=K, L=L, prior=prior)
    pg, po = _both(env, w, 0)
    # without any prior on frame 0 the dropped block is conditioned ~1e11 (bias random walk vs depth curvature)
    _compare(pg, po, K, lambda f: f + 1, htol=1e-7 if prior != "none" else 2e-5)


def test_chain_and_second_new(env):
    abi, synth, orc, ctx = env
    K = 11
    w = synth.make_window(seed=5, K=K, L=80)
    pg, po = _both(env, w, 0)
    _compare(pg, po, K, lambda f: f + 1)
    # feed the DEVICE prior forward: MARGIN_SECOND_NEW and a second MARGIN_OLD on top of it
    w2 = dataclasses.replace(w, prior={k: pg[k] for k in KEYS})
    pg2, po2 = _both(env, w2, 1)
    assert pg2["n"] == pg["n"] - 6
    _compare(pg2, po2, K, lambda f: f)
    pg3, po3 = _both(env, w2, 0)
    _compare(pg3, po3, K, lambda f: f + 1)
    # prior that does not touch Pose[K-2]: unchanged (n = -1)
    assert _both(env, w, 1) == (None, None)


def test_marginalized_prior_drives_the_next_solve(env):
    """optimize() with the device prior and with the oracle prior converge to the same state."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=8, K=11, L=120)
    pg, po = _both(env, w, 0)
    tight = abi.default_opts(max_iters=40, function_tolerance=1e-14, gradient_tolerance=1e-12, parameter_tolerance=1e-14)
    xs = []
    for p in (pg, po):
        w2 = dataclasses.replace(synth.make_window(seed=9, K=11, L=120, prior="none"), prior={k: p[k] for k in KEYS})
        h, s = abi.WindowHandle(w2), abi.Summary()
        ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(tight), C.byref(s)), "optimize")
        xs.append(h.state_vector())
    assert np.linalg.norm(xs[0] - xs[1]) <= 1e-6 * np.linalg.norm(xs[1])


def _info_td(p, K, unshift):
    """like info_in_state_coords, with the td column after the extrinsic block"""
    M = 15 * K + 7
    cols = np.full(p["n"], -1)
    for kind, frame, idx in zip(p["block_kind"], p["block_frame"], p["block_idx"]):
        f = unshift(int(frame)) if kind in (0, 1) else 0
        if kind == 0:
            cols[idx:idx + 6] = 15 * f + np.arange(6)
        elif kind == 1:
            cols[idx:idx + 9] = 15 * f + 6 + np.arange(9)
        elif kind == 2:
            cols[idx:idx + 6] = 15 * K + np.arange(6)
        else:
            cols[idx] = 15 * K + 6
    assert (cols >= 0).all()
    H, g = np.zeros((M, M)), np.zeros(M)
    H[np.ix_(cols, cols)] = p["J"].T @ p["J"]
    g[cols] = p["J"].T @ p["lin_res"]
    return H, g


@pytest.mark.parametrize("seed,K,L,TR", [(0, 11, 150, 0.0), (1, 6, 40, 0.02), (2, 11, 300, 0.01)])
def test_margin_old_with_td_matches_oracle(env, seed, K, L, TR):
    """ESTIMATE_TD: ProjectionTdFactor in the marginalization (estimator.cpp:863-871), para_Td kept."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=seed, K=K, L=L, td_true=0.004)
    w.para_td[0] = 0.002
    pg, po = _both(env, w, 0, estimate_td=1, TR=TR)
    assert pg["n"] == po["n"] and pg["block_kind"][-1] == 3
    for k in ("block_kind", "block_frame", "block_idx"):
        assert np.array_equal(pg[k], po[k]), k
    assert np.array_equal(pg["x0"], po["x0"])
    Hg, gg = _info_td(pg, K, lambda f: f + 1)
    Ho, go = _info_td(po, K, lambda f: f + 1)
    assert Ho[-1, -1] > 0
    assert np.abs(Hg - Ho).max() <= 1e-7 * np.abs(Ho).max()
    assert np.abs(gg - go).max() <= 5e-5 * max(np.abs(go).max(), 1.0)


def test_td_prior_drives_the_next_solve(env):
    """A prior that carries the td block feeds an ESTIMATE_TD solve: device and oracle reach the same state."""
    abi, synth, orc, ctx = env
    w = synth.make_window(seed=8, K=11, L=120, td_true=0.004)
    pg, po = _both(env, w, 0, estimate_td=1)
    w2 = dataclasses.replace(synth.make_window(seed=9, K=11, L=120, prior="none", td_true=0.004), prior={k: pg[k] for k in KEYS})
    o = abi.default_opts(estimate_td=1)
    hg, ho, sg, so = abi.WindowHandle(w2), abi.WindowHandle(w2), abi.Summary(), abi.Summary()
    ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(hg.s), C.byref(o), C.byref(sg)), "optimize")
    assert orc.oracle_optimize(C.byref(ho.s), C.byref(o), C.byref(so)) == 0
    assert (sg.iterations, sg.num_accepted, sg.termination) == (so.iterations, so.num_accepted, so.termination)
    assert np.linalg.norm(hg.state_vector() - ho.state_vector()) <= 1e-8 * np.linalg.norm(ho.state_vector())
    assert abs(hg.td[0] - ho.td[0]) <= 1e-9


def test_marginalize_begin_end_equals_synchronous_call(pkg):
    """bvio_marginalize_begin / _end (second stream, pinned staging) around a bvio_select on the main stream: the prior
    is bit-identical to the synchronous call's, the selection is undisturbed, and a second begin is refused."""
    abi, synth = pkg.abi, pkg.synth
    ctx = pkg.lib.Context(0)
    for seed, flag in ((3, 0), (4, 0), (5, 1)):
        w0 = synth.make_window(seed=seed, K=11, L=150)
        p0 = run_marg(abi, ctx.L.bvio_marginalize, w0, 0, ctx=ctx.h)
        w = dataclasses.replace(synth.make_window(seed=seed + 50, K=11, L=150),
                                prior={k: p0[k] for k in ("n", "block_kind", "block_frame", "block_idx", "x0", "lin_jac", "lin_res")})
        ps = run_marg(abi, ctx.L.bvio_marginalize, w, flag, ctx=ctx.h)
        prob = synth.make_select_problem(seed=seed, N=300, H=10, kappa=20)
        hs, ss, ids0, ids1 = abi.SelectHandle(prob), abi.SelectSummary(), np.zeros(20, np.int32), np.zeros(20, np.int32)
        ctx.check(ctx.L.bvio_select(ctx.h, C.byref(hs.s), abi.iptr(ids0), None, C.byref(ss)), "select")
        job = abi.MarginalizeJob(ctx.L, ctx.h, w, flag)
        with pytest.raises(RuntimeError):
            abi.MarginalizeJob(ctx.L, ctx.h, w, flag)                       # one in flight per context
        ctx.check(ctx.L.bvio_select(ctx.h, C.byref(hs.s), abi.iptr(ids1), None, C.byref(ss)), "select")
        pa = job.end()
        assert (ids0 == ids1).all()
        assert pa["n"] == ps["n"] and np.array_equal(pa["lin_jac"], ps["lin_jac"]) and np.array_equal(pa["lin_res"], ps["lin_res"])
        assert np.array_equal(pa["x0"], ps["x0"]) and np.array_equal(pa["block_idx"], ps["block_idx"])
    ctx.close()
