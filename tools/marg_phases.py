"""Where bvio_marginalize's time goes: device phases (BVIO_DEBUG stamps inside ba_marginalize_kernel) and the host-side
call, on windows with a real chained n = 75 prior.  Usage: BVIO_DEBUG=1 python tools/marg_phases.py"""
import dataclasses
import sys
import time

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
import test_oracle_marg as tm

pkg = g.load_package(); abi, synth = pkg.abi, pkg.synth
ctx = pkg.lib.Context(0)
keys = ("n", "block_kind", "block_frame", "block_idx", "x0", "lin_jac", "lin_res")
for L in (150, 1500):
    p0 = tm.run_marg(abi, ctx.L.bvio_marginalize, synth.make_window(seed=300, K=11, L=L), 0, ctx=ctx.h)
    w = dataclasses.replace(synth.make_window(seed=301, K=11, L=L), prior={k: p0[k] for k in keys})
    for flag in (0, 1):
        for i in range(4):
            t = time.perf_counter()
            p = tm.run_marg(abi, ctx.L.bvio_marginalize, w, flag, ctx=ctx.h)
            print(f"L {L} flag {flag}: call {(time.perf_counter() - t) * 1e3:.3f} ms (ffi {abi.call_marginalize.t_call * 1e3:.3f} ms) n {p['n'] if p else None}", flush=True)
