"""GPU parity against the REFERENCE's own code, directly: libbvio.so next to oracle/_ref/libvins_ref.so (the reference's
sources compiled here from /root/reference, see tests/test_reference_pin.py; the prebuilt library travels to the GPU
box), and against the golden vectors that code produced (tests/golden/ref_*.npz).  The library-based tests are skipped
when the library is absent.  (File name: sorts last, after the suites whose device side was exercised first.)"""
import ctypes as C
import os

import numpy as np
import pytest

import golden_io
import ref_lib
from test_oracle_marg import info_in_state_coords, run_marg
from test_ref_golden import GOLD, SEL, _marg_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    L = ref_lib.load()
    if L is None:
        pytest.skip("oracle/_ref/libvins_ref.so not built and /root/reference not present")
    return L


@pytest.mark.parametrize("seed,N,U,n_lm,kappa", [(0, 120, 0, 60, 25), (1, 150, 12, 80, 30), (3, 200, 20, 120, 40)])
def test_cuda_select_matches_reference_select(pkg, ref, seed, N, U, n_lm, kappa):
    """bvio_select vs FeatureSelector::select (the reference's code, nanoflann included): same ids, same order."""
    from test_reference_pin import reference_select_case
    abi = pkg.abi
    ref_ids, prob = reference_select_case(pkg, ref, seed, N, U, n_lm, kappa)
    ctx = pkg.lib.Context(0)
    hs, ss = abi.SelectHandle(prob), abi.SelectSummary()
    out = np.zeros(kappa, np.int32)
    ctx.check(ctx.L.bvio_select(ctx.h, C.byref(hs.s), abi.iptr(out), None, C.byref(ss)), "bvio_select")
    ctx.close()
    assert ss.n_selected == len(ref_ids) and (out[:ss.n_selected] == ref_ids).all(), (out[:ss.n_selected], ref_ids)


@pytest.mark.parametrize("seed,N,U,n_lm,kappa", [(0, 120, 0, 60, 25), (1, 150, 12, 80, 30), (3, 200, 20, 120, 40), (4, 180, 8, 100, 35)])
def test_cuda_select_matches_reference_in_ground_truth_horizon_mode(pkg, ref, tmp_path, seed, N, U, n_lm, kappa):
    """USE_GT, the reference's shipped EuRoC default (config/euroc/euroc_config.yaml:88): horizon from the ground-truth
    csv, candidates back-projected with the IMU-propagated state_k1_ (feature_selector.cpp:247-266,321-326) handed over
    in bvio_select_in.state_k1_pos / state_k1_quat (ABI v2).  Same ids as FeatureSelector::select, same order."""
    from test_reference_pin import gt_mode_case
    abi = pkg.abi
    ref_ids, prob = gt_mode_case(pkg, ref, seed, N, U, n_lm, kappa, str(tmp_path / "gt.csv"))
    ctx = pkg.lib.Context(0)
    hs, ss = abi.SelectHandle(prob), abi.SelectSummary()
    out = np.zeros(kappa, np.int32)
    ctx.check(ctx.L.bvio_select(ctx.h, C.byref(hs.s), abi.iptr(out), None, C.byref(ss)), "bvio_select")
    ctx.close()
    assert len(ref_ids) > 0 and out[:ss.n_selected].tolist() == ref_ids.tolist(), (out[:ss.n_selected], ref_ids)


@pytest.mark.parametrize("seed,twins", [(0, 10), (2, 6)])
def test_cuda_select_duplicate_candidates_like_the_reference(pkg, ref, seed, twins):
    """The `UBs[ub] = feature_id` collision (feature_selector.cpp:697,724): of two bit-identical candidates the
    reference names the larger id first.  The device's exact-tie rule does the same: identical ids, in order."""
    from test_reference_pin import reference_select_case
    abi = pkg.abi
    ref_ids, prob = reference_select_case(pkg, ref, seed, 60, 0, 60, 25, twins=twins)
    ctx = pkg.lib.Context(0)
    hs, ss = abi.SelectHandle(prob), abi.SelectSummary()
    out = np.zeros(25, np.int32)
    ctx.check(ctx.L.bvio_select(ctx.h, C.byref(hs.s), abi.iptr(out), None, C.byref(ss)), "bvio_select")
    ctx.close()
    assert out[:ss.n_selected].tolist() == ref_ids.tolist(), (out[:ss.n_selected], ref_ids)
    order = {int(i): k for k, i in enumerate(out[:ss.n_selected])}
    pairs = [(int(prob.cand_id[k]), int(prob.cand_id[k + 1])) for k in range(0, 2 * twins, 2)]
    both = [(a, b) for a, b in pairs if a in order and b in order]
    assert both and all(order[b] < order[a] for a, b in both)


@pytest.mark.parametrize("seed,K,L", [(0, 11, 150), (1, 6, 40)])
def test_cuda_marginalize_matches_reference_marginalize(pkg, ref, seed, K, L):
    """bvio_marginalize vs MarginalizationInfo::marginalize on the reference's residual blocks: the new prior's
    quadratic form in state coordinates (tolerances of tests/test_gpu_marg.py)."""
    abi, synth = pkg.abi, pkg.synth
    w = synth.make_window(seed=seed, K=K, L=L)
    ctx = pkg.lib.Context(0)
    pg = run_marg(abi, ctx.L.bvio_marginalize, w, 0, ctx=ctx.h)
    ctx.close()
    pr = run_marg(abi, ref.ref_marginalize, w, 0)
    assert pg["n"] == pr["n"]
    Hg, gg = info_in_state_coords(pg, K, lambda f: f + 1)
    Hr, gr = info_in_state_coords(pr, K, lambda f: f + 1)
    # device vs oracle is held to 1e-7 / 5e-5 (tests/test_gpu_marg.py), oracle vs reference measures 4e-9 / 5e-11
    assert np.abs(Hg - Hr).max() <= 2e-7 * np.abs(Hr).max()
    assert np.abs(gg - gr).max() <= 1e-4 * max(np.abs(gr).max(), 1.0)


@pytest.mark.parametrize("seed,L,ex,td", [(0, 150, 0, 0), (1, 80, 1, 1)])
def test_cuda_linearization_matches_reference_cost_functions(pkg, ref, seed, L, ex, td):
    """bvio_debug_linearize (the device's reduced system S, g, h, b and cost) vs the Schur-reduced normal equations that
    the reference's own Estimator::optimization() problem -- its cost functions, its loss corrector -- yields."""
    from test_reference_pin import reference_reduced_system
    abi, synth = pkg.abi, pkg.synth
    K = 11
    w = synth.make_window(seed=seed, K=K, L=L, **(dict(td_true=0.003) if td else {}))
    o_kw = dict(estimate_extrinsic=ex, estimate_td=td, TR=0.01 if td else 0.0)
    S_r, g_r, h_r, b_r, cost_r = reference_reduced_system(pkg, ref, w, **o_kw)
    npar = 15 * K + 6 * ex + td
    S, g, h, b, c = np.zeros((npar, npar)), np.zeros(npar), np.zeros(w.L), np.zeros(w.L), np.zeros(1)
    ctx = pkg.lib.Context(0)
    hw, o = abi.WindowHandle(w), abi.default_opts(**o_kw)
    ctx.check(ctx.L.bvio_debug_linearize(ctx.h, C.byref(hw.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(h), abi.dptr(b), abi.dptr(c)),
              "debug_linearize")
    ctx.close()
    assert abs(c[0] - cost_r) <= 1e-9 * cost_r
    assert np.abs(h - h_r).max() <= 1e-9 * np.abs(h_r).max() and np.abs(b - b_r).max() <= 1e-8 * np.abs(b_r).max()
    assert np.abs(S - S_r).max() <= 1e-6 * np.abs(S_r).max() and np.abs(g - g_r).max() <= 1e-6 * np.abs(g_r).max()


@pytest.mark.parametrize("name", SEL)
def test_cuda_matches_reference_selection_vector(pkg, name):
    abi = pkg.abi
    d = np.load(os.path.join(GOLD, name))
    prob = golden_io.select_from_dict(d)
    ctx = pkg.lib.Context(0)
    hs, ss = abi.SelectHandle(prob), abi.SelectSummary()
    out = np.zeros(prob.kappa, np.int32)
    ctx.check(ctx.L.bvio_select(ctx.h, C.byref(hs.s), abi.iptr(out), None, C.byref(ss)), "bvio_select")
    ctx.close()
    assert out[:ss.n_selected].tolist() == d["out_ids"].tolist()


def test_cuda_matches_reference_marginalization_vector(pkg):
    d, w = _marg_case()
    ctx = pkg.lib.Context(0)
    p = run_marg(pkg.abi, ctx.L.bvio_marginalize, w, 0, ctx=ctx.h)
    ctx.close()
    H, g = info_in_state_coords(p, w.K, lambda f: f + 1)
    assert p["n"] == int(d["out_n"][0])
    assert np.abs(H - d["out_H"]).max() <= 2e-7 * np.abs(d["out_H"]).max()
    assert np.abs(g - d["out_g"]).max() <= 1e-4 * max(np.abs(d["out_g"]).max(), 1.0)


@pytest.mark.parametrize("skip", [3, 1])
def test_cuda_leaves_out_imu_intervals_longer_than_10s(pkg, oracle, skip):
    """`if (pre_integrations[j]->sum_dt > 10.0) continue;` (estimator.cpp:705) and the `< 10.0` test of the MARGIN_OLD
    branch (:846): the device's linearization, solve and marginalization against the oracle (which the reference's
    compiled Estimator::optimization() confirms on the same case, tests/test_reference_pin.py)."""
    import dataclasses
    abi, synth = pkg.abi, pkg.synth
    K = 11
    w = synth.make_window(seed=5 + skip, K=K, L=60)
    w.preint[skip, 16] = 11.0
    ctx = pkg.lib.Context(0)
    o = abi.default_opts()
    npar = 15 * K
    out = []
    for dev in (True, False):
        hw = abi.WindowHandle(w)
        S, g, h, b, c = np.zeros((npar, npar)), np.zeros(npar), np.zeros(w.L), np.zeros(w.L), np.zeros(1)
        if dev:
            ctx.check(ctx.L.bvio_debug_linearize(ctx.h, C.byref(hw.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(h), abi.dptr(b),
                                                 abi.dptr(c)), "debug_linearize")
        else:
            assert oracle.oracle_linearize(C.byref(hw.s), C.byref(o), abi.dptr(S), abi.dptr(g), abi.dptr(h), abi.dptr(b), abi.dptr(c)) == 0
        out.append((S, g, c[0]))
    (S1, g1, c1), (S2, g2, c2) = out
    assert abs(c1 - c2) <= 1e-11 * c2 and np.abs(S1 - S2).max() <= 1e-9 * np.abs(S2).max() and np.abs(g1 - g2).max() <= 1e-9 * max(np.abs(g2).max(), 1.0)
    hg, ho, sg, so = abi.WindowHandle(w), abi.WindowHandle(w), abi.Summary(), abi.Summary()
    ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(hg.s), C.byref(o), C.byref(sg)), "bvio_optimize")
    assert oracle.oracle_optimize(C.byref(ho.s), C.byref(o), C.byref(so)) == 0
    assert (sg.iterations, sg.num_accepted, sg.termination) == (so.iterations, so.num_accepted, so.termination)
    assert np.linalg.norm(hg.state_vector() - ho.state_vector()) <= 1e-6 * np.linalg.norm(ho.state_vector())
    wpost = dataclasses.replace(w, para_pose=ho.pose.copy(), para_speed_bias=ho.sb.copy(), inv_depth=ho.inv.copy())
    pg = run_marg(abi, ctx.L.bvio_marginalize, wpost, 0, ctx=ctx.h)
    ctx.close()
    po = run_marg(abi, oracle.oracle_marginalize, wpost, 0)
    assert pg["n"] == po["n"] == (66 if skip == 1 else 75)
    Hg, gg = info_in_state_coords(pg, K, lambda f: f + 1)
    Ho, go = info_in_state_coords(po, K, lambda f: f + 1)
    assert np.abs(Hg - Ho).max() <= 1e-7 * np.abs(Ho).max() and np.abs(gg - go).max() <= 5e-5 * max(np.abs(go).max(), 1.0)
