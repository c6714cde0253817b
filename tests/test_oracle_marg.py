"""CPU suite: pins the oracle's marginalization restatement (row a9) against an independent numpy
Schur complement of the dense Jacobian built by tests/np_ref.py, and against the reference's own
(commented) reconstruction invariant J^T J == A, J^T r == b (marginalization_factor.cpp:295-296)."""
import ctypes as C
import dataclasses

import numpy as np
import pytest

import np_ref


def run_marg(abi, lib_fn, w, flag, ctx=None, opts=None):
    h = abi.WindowHandle(w)
    cap_n, cap_b = 15 * w.K + 16, 2 * w.K + 2
    bk, bf, bi = (np.zeros(cap_b, np.int32) for _ in range(3))
    x0, jac, res = np.zeros(9 * cap_b), np.zeros(cap_n * cap_n), np.zeros(cap_n)
    out = abi.PriorOut()
    out.block_kind, out.block_frame, out.block_idx = abi.iptr(bk), abi.iptr(bf), abi.iptr(bi)
    out.x0, out.lin_jac, out.lin_res = abi.dptr(x0), abi.dptr(jac), abi.dptr(res)
    out.cap_n, out.cap_blocks = cap_n, cap_b
    o = opts if opts is not None else abi.default_opts()
    rc = lib_fn(C.byref(h.s), C.byref(o), flag, C.byref(out)) if ctx is None else \
        lib_fn(ctx, C.byref(h.s), C.byref(o), flag, C.byref(out))
    assert rc == 0, rc
    n, nb = out.n, out.nblocks
    if n < 0:
        return None
    J = jac[:n * n].reshape(n, n, order="F").copy()
    return dict(n=n, block_kind=bk[:nb].copy(), block_frame=bf[:nb].copy(), block_idx=bi[:nb].copy(),
                x0=x0.copy(), lin_jac=jac[:n * n].copy(), lin_res=res[:n].copy(), J=J)


def info_in_state_coords(p, K, unshift):
    """J^T J and J^T r of a prior scattered to the 15K (+6 extrinsic) frame-major layout."""
    M = 15 * K + 6
    cols = np.full(p["n"], -1)
    for kind, frame, idx in zip(p["block_kind"], p["block_frame"], p["block_idx"]):
        f = unshift(int(frame)) if kind in (0, 1) else 0
        if kind == 0:
            cols[idx:idx + 6] = 15 * f + np.arange(6)
        elif kind == 1:
            cols[idx:idx + 9] = 15 * f + 6 + np.arange(9)
        elif kind == 2:
            cols[idx:idx + 6] = 15 * K + np.arange(6)
    assert (cols >= 0).all()
    H, g = np.zeros((M, M)), np.zeros(M)
    H[np.ix_(cols, cols)] = p["J"].T @ p["J"]
    g[cols] = p["J"].T @ p["lin_res"]
    return H, g


def numpy_margin_old(w):
    """Independent: dense Jacobian of the MARGIN_OLD factor subset, Schur complement with pinv."""
    K = w.K
    keep = [l for l in range(w.L) if w.obs_frame[w.lm_obs_offset[l]] == 0]
    offs, fr, xy = [0], [], []
    for l in keep:
        o0, o1 = w.lm_obs_offset[l], w.lm_obs_offset[l + 1]
        fr += list(w.obs_frame[o0:o1])
        xy += list(w.obs_xy[o0:o1])
        offs.append(len(fr))
    pre = w.preint.copy()
    pre[2:, 16] = 11.0                                    # sum_dt > 10: only the 0 -> 1 IMU factor stays
    sub = dataclasses.replace(w, inv_depth=w.inv_depth[keep], lm_obs_offset=np.array(offs, np.int32),
                              obs_frame=np.array(fr, np.int32), obs_xy=np.array(xy).reshape(-1, 2), preint=pre)
    J, r, _ = np_ref.full_system(sub)
    H, g = J.T @ J, J.T @ r
    npar = 15 * K
    dropped = np.r_[np.arange(15), npar + np.arange(len(keep))]
    rest = np.setdiff1d(np.arange(npar), np.arange(15))
    Hmm = 0.5 * (H[np.ix_(dropped, dropped)] + H[np.ix_(dropped, dropped)].T)
    wv, V = np.linalg.eigh(Hmm)
    inv = V @ np.diag(np.where(wv > 1e-8, 1.0 / np.where(wv > 1e-8, wv, 1.0), 0.0)) @ V.T
    Hs = H[np.ix_(rest, rest)] - H[np.ix_(rest, dropped)] @ inv @ H[np.ix_(dropped, rest)]
    gs = g[rest] - H[np.ix_(rest, dropped)] @ inv @ g[dropped]
    return rest, Hs, gs


@pytest.mark.parametrize("seed,K,L", [(0, 11, 150), (1, 6, 40), (2, 11, 30)])
def test_margin_old_matches_numpy_schur(pkg, oracle, seed, K, L):
    abi, synth = pkg.abi, pkg.synth
    w = synth.make_window(seed=seed, K=K, L=L)
    p = run_marg(abi, oracle.oracle_marginalize, w, 0)
    assert p["n"] == p["J"].shape[0] and (p["block_frame"][p["block_kind"] < 2] >= 0).all()
    H, g = info_in_state_coords(p, K, lambda f: f + 1)     # undo addr_shift
    rest, Hs, gs = numpy_margin_old(w)
    sc = np.abs(Hs).max()
    # cond(Amm) ~ 1e8 (bias prior 1e8 next to depth curvatures 1e3): two eigen-solvers agree to ~1e-8
    assert np.abs(H[np.ix_(rest, rest)] - Hs).max() <= 1e-7 * sc
    assert np.abs(g[rest] - gs).max() <= 5e-5 * max(np.abs(gs).max(), 1.0)   # |Arm| |Amm^+| |bmm| * 1e-16 ~ 1e-2
    # frame 0 is gone, nothing refers to the (shifted) last frame that no factor touched
    assert np.abs(H[:15, :15]).max() == 0.0
    # x0 = the window state of the kept blocks
    off = 0
    for kind, frame in zip(p["block_kind"], p["block_frame"]):
        src = w.para_pose[frame + 1] if kind == 0 else (w.para_speed_bias[frame + 1] if kind == 1 else w.para_ex_pose)
        assert np.array_equal(p["x0"][off:off + len(src)], src)
        off += len(src)


def test_prior_chain_and_second_new(pkg, oracle):
    """MARGIN_OLD output is a valid prior for the next call; MARGIN_SECOND_NEW drops Pose[K-2] only."""
    abi, synth = pkg.abi, pkg.synth
    K = 11
    w = synth.make_window(seed=5, K=K, L=80)
    p1 = run_marg(abi, oracle.oracle_marginalize, w, 0)
    w2 = dataclasses.replace(w, prior={k: p1[k] for k in ("n", "block_kind", "block_frame", "block_idx", "x0", "lin_jac", "lin_res")})
    # the new prior touches Pose[0..K-2] (shifted) => SECOND_NEW applies and removes exactly Pose[K-2]
    assert K - 2 in p1["block_frame"][p1["block_kind"] == 0]
    p2 = run_marg(abi, oracle.oracle_marginalize, w2, 1)
    assert p2 is not None and p2["n"] == p1["n"] - 6
    assert K - 2 not in p2["block_frame"][p2["block_kind"] == 0]
    H1, g1 = info_in_state_coords(p1, K, lambda f: f)
    H2, g2 = info_in_state_coords(p2, K, lambda f: f)
    # marginalising a block of a Gaussian prior evaluated AT its linearisation point: Schur complement
    d = np.arange(15 * (K - 2), 15 * (K - 2) + 6)
    keep = np.setdiff1d(np.flatnonzero(np.abs(H1).sum(0) > 0), d)
    S = H1[np.ix_(keep, keep)] - H1[np.ix_(keep, d)] @ np.linalg.pinv(H1[np.ix_(d, d)]) @ H1[np.ix_(d, keep)]
    assert np.abs(H2[np.ix_(keep, keep)] - S).max() <= 1e-8 * np.abs(S).max()
    # chained MARGIN_OLD with the marginalised prior still runs and stays PSD
    p3 = run_marg(abi, oracle.oracle_marginalize, w2, 0)
    ev = np.linalg.eigvalsh(p3["J"].T @ p3["J"])
    assert ev.min() >= -1e-8 * ev.max()
    # no prior on Pose[K-2] => unchanged
    assert run_marg(abi, oracle.oracle_marginalize, w, 1) is None


@pytest.mark.parametrize("seed,K,L", [(0, 8, 60), (4, 11, 120)])
def test_margin_old_with_td_matches_numpy_schur(pkg, oracle, seed, K, L):
    """MARGIN_OLD with ESTIMATE_TD (estimator.cpp:863-871): ProjectionTdFactor blocks in the marginalized set; the kept
    prior spans poses, speed/bias, the extrinsic AND para_Td.  Checked against the Schur complement of the dense numpy
    system with the extrinsic and td columns."""
    abi, synth = pkg.abi, pkg.synth
    w = synth.make_window(seed=seed, K=K, L=L, td_true=0.003)
    w.para_td[0] = 0.001
    TR = 0.015
    p = run_marg(abi, oracle.oracle_marginalize, w, 0, opts=abi.default_opts(estimate_td=1, TR=TR))
    assert p["block_kind"][-1] == 3
    itd = int(p["block_idx"][-1])
    # scatter to [15K | 6 | 1]
    M = 15 * K + 7
    cols = np.full(p["n"], -1)
    for kind, frame, idx in zip(p["block_kind"], p["block_frame"], p["block_idx"]):
        if kind == 0:
            cols[idx:idx + 6] = 15 * (frame + 1) + np.arange(6)
        elif kind == 1:
            cols[idx:idx + 9] = 15 * (frame + 1) + 6 + np.arange(9)
        elif kind == 2:
            cols[idx:idx + 6] = 15 * K + np.arange(6)
        else:
            cols[idx] = 15 * K + 6
    H, g = np.zeros((M, M)), np.zeros(M)
    H[np.ix_(cols, cols)] = p["J"].T @ p["J"]
    g[cols] = p["J"].T @ p["lin_res"]
    # numpy: the MARGIN_OLD factor subset with free extrinsic and td
    keep = [l for l in range(w.L) if w.obs_frame[w.lm_obs_offset[l]] == 0]
    sel = np.concatenate([np.arange(w.lm_obs_offset[l], w.lm_obs_offset[l + 1]) for l in keep])
    offs = np.concatenate([[0], np.cumsum([w.lm_obs_offset[l + 1] - w.lm_obs_offset[l] for l in keep])]).astype(np.int32)
    pre = w.preint.copy()
    pre[2:, 16] = 11.0
    sub = dataclasses.replace(w, inv_depth=w.inv_depth[keep], lm_obs_offset=offs, obs_frame=w.obs_frame[sel],
                              obs_xy=w.obs_xy[sel], obs_vel=w.obs_vel[sel], obs_td=w.obs_td[sel], obs_row=w.obs_row[sel],
                              preint=pre)
    Jd, rd, _ = np_ref.full_system_ext(sub, est_ex=True, est_td=True, TR=TR)
    Hd, gd = Jd.T @ Jd, Jd.T @ rd
    dropped = np.r_[np.arange(15), M + np.arange(len(keep))]
    rest = np.setdiff1d(np.arange(M), np.arange(15))
    Hmm = 0.5 * (Hd[np.ix_(dropped, dropped)] + Hd[np.ix_(dropped, dropped)].T)
    wv, V = np.linalg.eigh(Hmm)
    inv = V @ np.diag(np.where(wv > 1e-8, 1.0 / np.where(wv > 1e-8, wv, 1.0), 0.0)) @ V.T
    Hs = Hd[np.ix_(rest, rest)] - Hd[np.ix_(rest, dropped)] @ inv @ Hd[np.ix_(dropped, rest)]
    gs = gd[rest] - Hd[np.ix_(rest, dropped)] @ inv @ gd[dropped]
    sc = np.abs(Hs).max()
    assert np.abs(H[np.ix_(rest, rest)] - Hs).max() <= 1e-7 * sc
    assert np.abs(g[rest] - gs).max() <= 5e-5 * max(np.abs(gs).max(), 1.0)
    assert Hs[-1, -1] > 0 and abs(H[M - 1, M - 1] - Hs[-1, -1]) <= 1e-7 * sc       # information on td itself
    assert p["x0"][sum(7 if k == 0 else 9 if k == 1 else 7 for k in p["block_kind"][:-1])] == w.para_td[0]
    assert itd == p["n"] - 1


def test_td_marginalization_plumbing(pkg, oracle):
    """ESTIMATE_TD (estimator.cpp:863-871): para_Td is a kept block of the new prior.  With zero feature velocities the
    td factor is the plain projection factor and carries no information on td: same prior on every other block, one
    extra all-zero td column."""
    import dataclasses
    abi, synth = pkg.abi, pkg.synth
    w = synth.make_window(seed=3, K=8, L=40, td_true=0.0)
    w0 = dataclasses.replace(w, obs_vel=np.zeros_like(w.obs_vel))
    p_td = run_marg(abi, oracle.oracle_marginalize, w0, 0, opts=abi.default_opts(estimate_td=1))
    p = run_marg(abi, oracle.oracle_marginalize, w0, 0)
    assert p_td["n"] == p["n"] + 1 and p_td["block_kind"][-1] == 3 and p_td["block_idx"][-1] == p["n"]
    H, Htd = p["J"].T @ p["J"], p_td["J"].T @ p_td["J"]
    assert np.abs(Htd[-1]).max() == 0 and np.abs(Htd[:, -1]).max() == 0
    assert np.abs(Htd[:-1, :-1] - H).max() <= 1e-9 * np.abs(H).max()
    # real velocities: td couples with the poses
    p_real = run_marg(abi, oracle.oracle_marginalize, w, 0, opts=abi.default_opts(estimate_td=1, TR=0.01))
    Hr = p_real["J"].T @ p_real["J"]
    assert Hr[-1, -1] > 0 and np.abs(Hr[-1, :-1]).max() > 0
    ev = np.linalg.eigvalsh(Hr)
    assert ev.min() >= -1e-9 * ev.max()


def test_ql_eigensolver_mode_gives_the_same_prior(pkg, oracle, monkeypatch):
    """ORACLE_EIGH=ql (Householder + implicit QL, used for the timed CPU baseline) against the default cyclic Jacobi:
    same quadratic form on a well-conditioned kept system."""
    abi, synth, orc = pkg.abi, pkg.synth, oracle
    w = synth.make_window(seed=21, K=11, L=120)
    pj = run_marg(abi, orc.oracle_marginalize, w, 0)
    monkeypatch.setenv("ORACLE_EIGH", "ql")
    pq = run_marg(abi, orc.oracle_marginalize, w, 0)
    assert pj["n"] == pq["n"]
    Hj, gj = info_in_state_coords(pj, w.K, lambda f: f + 1)
    Hq, gq = info_in_state_coords(pq, w.K, lambda f: f + 1)
    assert np.abs(Hj - Hq).max() <= 1e-7 * np.abs(Hj).max()          # (measured 2e-8: the eps = 1e-8 pseudo-inverse sees different roundoff)
    assert np.abs(gj - gq).max() <= 1e-4 * max(np.abs(gj).max(), 1.0)
