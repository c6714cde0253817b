#!/usr/bin/env python
"""Generates tests/golden/ref_*.npz: outputs of the REFERENCE's own code (oracle/_ref/libvins_ref.so = the reference's
factor / marginalization / selector sources compiled from /root/reference) on stored inputs.
Unlike the library, these vectors travel: tests/test_ref_golden.py checks the oracle (CPU suite) and the CUDA path (GPU
suite) against them on any machine.  Needs /root/reference.  Run from the repo root:
    python tests/golden/make_ref_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402
import ref_lib  # noqa: E402
from golden_io import select_to_dict, window_to_dict  # noqa: E402
from test_oracle_marg import info_in_state_coords, run_marg  # noqa: E402
import test_reference_pin as T  # noqa: E402


def main():
    pkg = g.load_package()
    abi, synth = pkg.abi, pkg.synth
    ref = ref_lib.load()
    assert ref is not None, "oracle/_ref/libvins_ref.so missing and /root/reference absent"
    rng = np.random.default_rng(2024)
    # ---- factors -------------------------------------------------------------------------------------------------------------
    n = 24
    ex = T._ex(synth)
    inp = dict(pts_i=np.zeros((n, 3)), pts_j=np.zeros((n, 3)), pose_i=np.zeros((n, 7)), pose_j=np.zeros((n, 7)), lam=np.zeros(n),
               vel_i=np.zeros((n, 2)), vel_j=np.zeros((n, 2)), tds=np.zeros((n, 3)), rows=np.zeros((n, 2)), TR=np.zeros(n))
    out_p, out_td = np.zeros((n, 2 + 14 * 3 + 2)), np.zeros((n, 2 + 14 * 3 + 2 + 2))
    for k in range(n):
        pi, pj = T._rand_pose(rng, 0.3), T._rand_pose(rng, 0.3)
        pj[3:] = T.np_ref.pose_plus(pi, np.concatenate([np.zeros(3), rng.normal(size=3) * 0.2]))[3:]
        a = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4), 1.0])
        b = np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.4, 0.4), 1.0])
        lam = rng.uniform(0.1, 0.5)
        vi, vj = rng.normal(0, 0.3, 2), rng.normal(0, 0.3, 2)
        tds, rows, TR = rng.normal(0, 0.01, 3), rng.uniform(0, 480, 2), float(rng.choice([0.0, 0.02]))
        for key, v in (("pts_i", a), ("pts_j", b), ("pose_i", pi), ("pose_j", pj), ("lam", lam), ("vel_i", vi), ("vel_j", vj),
                       ("tds", tds), ("rows", rows), ("TR", TR)):
            inp[key][k] = v
        res, Ji, Jj, Jex, Jf, Jtd = np.zeros(2), np.zeros(14), np.zeros(14), np.zeros(14), np.zeros(2), np.zeros(2)
        ref.ref_projection_factor(abi.dptr(a), abi.dptr(b), abi.dptr(pi), abi.dptr(pj), abi.dptr(ex), lam, 460 / 1.5, abi.dptr(res),
                                  abi.dptr(Ji), abi.dptr(Jj), abi.dptr(Jex), abi.dptr(Jf))
        out_p[k] = np.concatenate([res, Ji, Jj, Jex, Jf])
        ref.ref_projection_td_factor(abi.dptr(a), abi.dptr(b), abi.dptr(vi), abi.dptr(vj), tds[0], tds[1], rows[0], rows[1], TR, 480.0,
                                     abi.dptr(pi), abi.dptr(pj), abi.dptr(ex), lam, tds[2], 460 / 1.5, abi.dptr(res), abi.dptr(Ji),
                                     abi.dptr(Jj), abi.dptr(Jex), abi.dptr(Jf), abi.dptr(Jtd))
        out_td[k] = np.concatenate([res, Ji, Jj, Jex, Jf, Jtd])
    # ---- preintegration + IMU factor ----------------------------------------------------------------------------------------
    m = 20
    ba, bg = rng.normal(0, 0.02, 3), rng.normal(0, 0.002, 3)
    acc, gyr, dt = rng.normal(0, 1, (m + 1, 3)) + [0, 0, 9.8], rng.normal(0, 0.3, (m + 1, 3)), np.full(m, 0.005)
    pre = abi.Preint()
    ref.ref_preintegrate(m, abi.dptr(dt), abi.dptr(acc.reshape(-1).copy()), abi.dptr(gyr.reshape(-1).copy()), abi.dptr(ba), abi.dptr(bg),
                         synth.ACC_N, synth.GYR_N, synth.ACC_W, synth.GYR_W, None, None, C.byref(pre))
    w = synth.make_window(seed=7, K=4, L=10)
    G = np.array([0, 0, synth.G_NORM])
    h = abi.WindowHandle(w)
    imu_out = np.zeros((3, 15 + 105 + 135 + 105 + 135))
    for j in range(1, 4):
        pre_c = C.cast(h.pre.ctypes.data + j * 467 * 8, C.POINTER(abi.Preint))
        res, J = np.zeros(15), [np.zeros(105), np.zeros(135), np.zeros(105), np.zeros(135)]
        ref.ref_imu_factor(pre_c, abi.dptr(G), abi.dptr(w.para_pose[j - 1].copy()), abi.dptr(w.para_speed_bias[j - 1].copy()),
                           abi.dptr(w.para_pose[j].copy()), abi.dptr(w.para_speed_bias[j].copy()), abi.dptr(res), *(abi.dptr(x) for x in J))
        imu_out[j - 1] = np.concatenate([res] + J)
    d = {"ex": ex, "proj_out": out_p, "td_out": out_td, "pre_ba": ba, "pre_bg": bg, "pre_acc": acc, "pre_gyr": gyr, "pre_dt": dt,
         "pre_out": np.frombuffer(bytes(pre), dtype=np.float64).copy(), "imu_out": imu_out}
    d.update({"f_" + k: v for k, v in inp.items()})
    d.update({"imuw_" + k: v for k, v in window_to_dict(w).items()})
    np.savez_compressed(os.path.join(HERE, "ref_factors.npz"), **d)
    # ---- marginalization: the reference's new prior as (J^T J, J^T r) in state coordinates --------------------------------------
    K = 8
    w = synth.make_window(seed=9, K=K, L=60)
    p1 = run_marg(abi, ref.ref_marginalize, w, 0)
    H1, g1 = info_in_state_coords(p1, K, lambda f: f + 1)
    d = window_to_dict(w)
    d.update(out_n=np.array([p1["n"]], np.int32), out_H=H1, out_g=g1)
    np.savez_compressed(os.path.join(HERE, "ref_marg_k8_l60.npz"), **d)
    # ---- FeatureSelector::select ------------------------------------------------------------------------------------------------
    for name, args in (("ref_sel_n150_u12", (1, 150, 12, 80, 30)), ("ref_sel_n120_u0", (0, 120, 0, 60, 25))):
        ids, prob = T.reference_select_case(pkg, ref, *args)
        d = select_to_dict(prob)
        d["out_ids"] = ids.astype(np.int32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.startswith("ref_")))


if __name__ == "__main__":
    main()
