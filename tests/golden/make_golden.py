#!/usr/bin/env python
"""Generates the committed golden vectors under tests/golden/ from the CPU oracle on seeded
synthetic inputs (the reference ships no fixtures for this path and cannot be imported or compiled
here -- DESIGN.md section 2).  Inputs are stored next to the outputs, so the fixtures do not depend
on numpy's RNG stream.  Run from the repo root:  python tests/golden/make_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402
import oracle_lib  # noqa: E402
from golden_io import window_to_dict, select_to_dict  # noqa: E402

TIGHT = dict(max_iters=50, function_tolerance=1e-14, gradient_tolerance=1e-12, parameter_tolerance=1e-14)


def main():
    pkg = g.load_package()
    abi, synth = pkg.abi, pkg.synth
    orc = oracle_lib.load()
    for name, kw in (("ba_k5_l12", dict(seed=41, K=5, L=12)), ("ba_k11_l150", dict(seed=42, K=11, L=150)),
                     ("ba_k11_l40_noprior", dict(seed=43, K=11, L=40, prior="none"))):
        w = synth.make_window(**kw)
        d = window_to_dict(w)
        np_ = 15 * w.K
        h = abi.WindowHandle(w)
        S, gv, hh, bb, c = np.zeros((np_, np_)), np.zeros(np_), np.zeros(w.L), np.zeros(w.L), np.zeros(1)
        o = abi.default_opts()
        assert orc.oracle_linearize(C.byref(h.s), C.byref(o), abi.dptr(S), abi.dptr(gv), abi.dptr(hh), abi.dptr(bb), abi.dptr(c)) == 0
        d.update(out_S=S, out_g=gv, out_h=hh, out_b=bb, out_cost=c)
        # reference budget (8 iterations, Ceres default tolerances)
        h8, s8 = abi.WindowHandle(w), abi.Summary()
        assert orc.oracle_optimize(C.byref(h8.s), C.byref(o), C.byref(s8)) == 0
        d.update(out8_state=h8.state_vector(), out8_summary=np.array([s8.iterations, s8.num_accepted, s8.num_rejected, s8.termination], np.int32),
                 out8_cost=np.array([s8.initial_cost, s8.final_cost]))
        # converged
        hc, sc = abi.WindowHandle(w), abi.Summary()
        assert orc.oracle_optimize(C.byref(hc.s), C.byref(abi.default_opts(**TIGHT)), C.byref(sc)) == 0
        d.update(outc_state=hc.state_vector(), outc_cost=np.array([sc.final_cost]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "cost", c[0], "->", s8.final_cost, "->", sc.final_cost)
    # option variants at the reference budget: {LM, DOGLEG} x estimate_extrinsic x estimate_td on one window whose
    # extrinsics are slightly off and whose observations are taken 4 ms late
    w = synth.make_window(seed=44, K=11, L=80, td_true=0.004)
    rng = np.random.default_rng(44)
    w.para_ex_pose[:3] += rng.normal(0, 0.02, 3)
    q = synth.quat_mul(w.para_ex_pose[3:], np.concatenate([0.5 * rng.normal(0, 0.01, 3), [1.0]]))
    w.para_ex_pose[3:] = q / np.linalg.norm(q)
    d = window_to_dict(w)
    for strategy in (0, 1):
        for ex in (0, 1):
            for td in (0, 1):
                o = abi.default_opts(strategy=strategy, estimate_extrinsic=ex, estimate_td=td, TR=0.01)
                h, s = abi.WindowHandle(w), abi.Summary()
                assert orc.oracle_optimize(C.byref(h.s), C.byref(o), C.byref(s)) == 0
                key = f"out_s{strategy}_e{ex}_t{td}"
                d[key + "_state"] = np.concatenate([h.state_vector(), h.td])
                d[key + "_summary"] = np.array([s.iterations, s.num_accepted, s.num_rejected, s.termination], np.int32)
                d[key + "_cost"] = np.array([s.initial_cost, s.final_cost, s.final_radius])
                print("variants", key, s.as_dict())
    np.savez_compressed(os.path.join(HERE, "opt_variants_k11_l80.npz"), **d)
    for name, kw in (("sel_n40_h10", dict(seed=51, N=40, H=10, kappa=8)), ("sel_n60_h13_u5", dict(seed=52, N=60, H=13, U=5, kappa=10)),
                     ("sel_n150_h10", dict(seed=53, N=150, H=10, kappa=30))):
        p = synth.make_select_problem(**kw)
        d = select_to_dict(p)
        h = abi.SelectHandle(p)
        T, D = 3 * p.H, 9 * (p.H + 1)
        Cc, valid, Om = np.zeros((p.N, T, T)), np.zeros(p.N, np.int32), np.zeros((D, D))
        orc.oracle_build_delta(C.byref(h.s), 0, None, abi.dptr(Cc), abi.iptr(valid), None)
        orc.oracle_omega_imu(C.byref(h.s), abi.dptr(Om))
        ids, vals, s = np.full(p.kappa, -1, np.int32), np.zeros(p.kappa), abi.SelectSummary()
        assert orc.oracle_select(C.byref(h.s), abi.iptr(ids), abi.dptr(vals), C.byref(s)) == 0
        d.update(out_C=Cc, out_valid=valid, out_omega=Om, out_ids=ids, out_vals=vals,
                 out_summary=np.array([s.n_selected, s.n_candidates_valid], np.int32),
                 out_final_logdet=np.array([s.final_logdet]), out_min_margin=np.array([s.min_margin]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, ids[:s.n_selected].tolist(), "margin", s.min_margin)


if __name__ == "__main__":
    main()
