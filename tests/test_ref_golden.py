"""Golden vectors produced by the REFERENCE's own compiled code (tests/golden/ref_*.npz, made by
tests/golden/make_ref_golden.py from oracle/_ref/libvins_ref.so).  They travel where the library cannot be rebuilt: the
oracle must reproduce them (here) and the CUDA path must match them (tests/test_zz_gpu_vs_reference.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import golden_io
from test_oracle_marg import info_in_state_coords, run_marg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEL = ["ref_sel_n120_u0.npz", "ref_sel_n150_u12.npz"]


def test_oracle_matches_reference_factor_vectors(pkg, oracle):
    abi, synth = pkg.abi, pkg.synth
    d = np.load(os.path.join(GOLD, "ref_factors.npz"))
    ex = d["ex"].copy()
    for k in range(len(d["f_lam"])):
        a, b, pi, pj = (d["f_" + n][k].copy() for n in ("pts_i", "pts_j", "pose_i", "pose_j"))
        lam, vi, vj, tds, rows, TR = float(d["f_lam"][k]), d["f_vel_i"][k].copy(), d["f_vel_j"][k].copy(), d["f_tds"][k], d["f_rows"][k], float(d["f_TR"][k])
        res, Ji, Jj, Jex, Jf, Jtd = np.zeros(2), np.zeros(14), np.zeros(14), np.zeros(14), np.zeros(2), np.zeros(2)
        oracle.oracle_projection_factor(abi.dptr(a), abi.dptr(b), abi.dptr(pi), abi.dptr(pj), abi.dptr(ex), lam, 460 / 1.5, abi.dptr(res),
                                        abi.dptr(Ji), abi.dptr(Jj), abi.dptr(Jex), abi.dptr(Jf))
        got = np.concatenate([res, Ji, Jj, Jex, Jf])
        assert np.abs(got - d["proj_out"][k]).max() <= 1e-12 * np.abs(got).max()
        oracle.oracle_projection_td_factor(abi.dptr(a), abi.dptr(b), abi.dptr(vi), abi.dptr(vj), float(tds[0]), float(tds[1]), float(rows[0]),
                                           float(rows[1]), TR, 480.0, abi.dptr(pi), abi.dptr(pj), abi.dptr(ex), lam, float(tds[2]), 460 / 1.5,
                                           abi.dptr(res), abi.dptr(Ji), abi.dptr(Jj), abi.dptr(Jex), abi.dptr(Jf), abi.dptr(Jtd))
        got = np.concatenate([res, Ji, Jj, Jex, Jf, Jtd])
        assert np.abs(got - d["td_out"][k]).max() <= 1e-12 * np.abs(got).max()
    # preintegration: 20 midpoint steps
    c = abi.Preint()
    c.delta_q[3] = 1.0
    for i in range(3):
        c.lin_ba[i], c.lin_bg[i] = d["pre_ba"][i], d["pre_bg"][i]
    for i in range(15):
        c.jacobian[i * 15 + i] = 1.0
    acc, gyr = d["pre_acc"], d["pre_gyr"]
    for k in range(len(d["pre_dt"])):
        oracle.oracle_preint_propagate(C.byref(c), float(d["pre_dt"][k]), abi.dptr(acc[k].copy()), abi.dptr(gyr[k].copy()),
                                       abi.dptr(acc[k + 1].copy()), abi.dptr(gyr[k + 1].copy()), synth.ACC_N, synth.GYR_N, synth.ACC_W, synth.GYR_W)
    got, want = np.frombuffer(bytes(c), dtype=np.float64), d["pre_out"]
    assert np.abs(got[:17] - want[:17]).max() <= 1e-14
    assert np.abs(got[17:] - want[17:]).max() <= 1e-13 * np.abs(want[17:]).max()
    # IMU factor
    w = golden_io.window_from_dict({k[5:]: d[k] for k in d.files if k.startswith("imuw_")})
    h = abi.WindowHandle(w)
    G = np.array([0, 0, synth.G_NORM])
    for j in range(1, 4):
        pre_c = C.cast(h.pre.ctypes.data + j * 467 * 8, C.POINTER(abi.Preint))
        res, J = np.zeros(15), [np.zeros(105), np.zeros(135), np.zeros(105), np.zeros(135)]
        oracle.oracle_imu_factor(pre_c, abi.dptr(G), abi.dptr(w.para_pose[j - 1].copy()), abi.dptr(w.para_speed_bias[j - 1].copy()),
                                 abi.dptr(w.para_pose[j].copy()), abi.dptr(w.para_speed_bias[j].copy()), abi.dptr(res), *(abi.dptr(x) for x in J))
        got, want = np.concatenate([res] + J), d["imu_out"][j - 1]
        assert np.abs(got - want).max() <= 2e-7 * np.abs(want).max()


def _marg_case():
    d = np.load(os.path.join(GOLD, "ref_marg_k8_l60.npz"))
    return d, golden_io.window_from_dict(d)


def test_oracle_matches_reference_marginalization_vector(pkg, oracle):
    d, w = _marg_case()
    p = run_marg(pkg.abi, oracle.oracle_marginalize, w, 0)
    H, g = info_in_state_coords(p, w.K, lambda f: f + 1)
    assert p["n"] == int(d["out_n"][0])
    assert np.abs(H - d["out_H"]).max() <= 1e-7 * np.abs(d["out_H"]).max()
    assert np.abs(g - d["out_g"]).max() <= 5e-5 * max(np.abs(d["out_g"]).max(), 1.0)


@pytest.mark.parametrize("name", SEL)
def test_oracle_matches_reference_selection_vector(pkg, oracle, name):
    abi = pkg.abi
    d = np.load(os.path.join(GOLD, name))
    prob = golden_io.select_from_dict(d)
    hs, ss = abi.SelectHandle(prob), abi.SelectSummary()
    out = np.zeros(prob.kappa, np.int32)
    assert oracle.oracle_select(C.byref(hs.s), abi.iptr(out), None, C.byref(ss)) == 0
    assert out[:ss.n_selected].tolist() == d["out_ids"].tolist()
