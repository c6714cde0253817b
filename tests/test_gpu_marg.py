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


def _both(env, w, flag):
    abi, synth, orc, ctx = env
    pg = run_marg(abi, ctx.L.bvio_marginalize, w, flag, ctx=ctx.h)
    po = run_marg(abi, orc.oracle_marginalize, w, flag)
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
