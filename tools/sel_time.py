#!/usr/bin/env python
"""Device time of one bvio_select for a few (N, H, kappa) shapes (single GPU)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
pkg = g.load_package()
abi, synth = pkg.abi, pkg.synth
ctx = pkg.lib.Context(0)
for N, H, kappa in ((2000, 10, 150), (2000, 13, 150), (333, 13, 40), (300, 13, 20), (2000, 16, 150)):
    p = synth.make_select_problem(seed=1, N=N, H=H, kappa=kappa)
    h = abi.SelectHandle(p)
    ids, s = np.zeros(kappa, np.int32), abi.SelectSummary()
    for _ in range(3):
        ctx.check(ctx.L.bvio_select(ctx.h, C.byref(h.s), abi.iptr(ids), None, C.byref(s)), "select")
    print(f"N={N} H={H} kappa={kappa}: {s.device_ms:.3f} ms, {s.candidates_scored / s.device_ms / 1e3:.1f} M cand/s")
ctx.close()
