"""Single-window latency (BASELINE configs 2 / 5: the reference's actual use): one 8-iteration dogleg solve of ONE
11-keyframe window through bvio_optimize with host buffers, the per-kernel device split at B = 1, and the CPU oracle
timed beside it.  Usage: python tools/single_window.py [--json]"""
import ctypes as C
import dataclasses
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g


KEYS = ("n", "block_kind", "block_frame", "block_idx", "x0", "lin_jac", "lin_res")
_cache = {}


def window_with_prior(pkg, L, marginalize):
    """The second of two consecutive windows of one session, carrying the n = 75 prior marginalized from the first
    (first window's state = its initial guess: the latency does not depend on it)."""
    synth = pkg.synth
    if L not in _cache:
        ses = synth.make_session(40, K=11, L=L)
        first, idx = synth.slice_window(ses, 0, 11)
        p0 = marginalize(first)
        _cache[L] = synth.consecutive_window(ses, first, idx, {k: p0[k] for k in KEYS})
    return _cache[L]


def measure_cpu(pkg, oracle, L, strategy=1):
    abi = pkg.abi
    w = window_with_prior(pkg, L, lambda a: abi.call_marginalize(oracle.oracle_marginalize, a, 0))
    o = abi.default_opts(strategy=strategy, max_iters=8)
    tc = []
    for r in range(3):
        h, so = abi.WindowHandle(w), abi.Summary()
        t0 = time.perf_counter()
        assert oracle.oracle_optimize(C.byref(h.s), C.byref(o), C.byref(so)) == 0
        tc.append(time.perf_counter() - t0)
    return {"cpu_oracle_ms": float(np.min(tc) * 1e3), "cpu_iterations": int(so.iterations)}


def measure(pkg, ctx, oracle, L, reps=30, strategy=1):
    abi, synth = pkg.abi, pkg.synth
    w = window_with_prior(pkg, L, lambda a: abi.call_marginalize(ctx.L.bvio_marginalize, a, 0, ctx=ctx.h))
    o = abi.default_opts(strategy=strategy, max_iters=8)
    ts, s = [], abi.Summary()
    for r in range(reps + 5):
        h = abi.WindowHandle(w)
        t0 = time.perf_counter()
        ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(o), C.byref(s)), "bvio_optimize")
        ts.append(time.perf_counter() - t0)
    ts = np.array(ts[5:]) * 1e3
    out = dict(L=L, n_factors=int(w.n_factors), iterations=int(s.iterations), e2e_ms_p50=float(np.percentile(ts, 50)),
               e2e_ms_p99=float(np.percentile(ts, 99)), device_ms=float(s.device_ms))
    # per-kernel split, resident
    hb = abi.WindowHandle(w)
    arr = (abi.WindowS * 1)(hb.s)
    ph = C.c_void_p()
    ctx.check(ctx.L.bvio_batch_upload(ctx.h, arr, 1, C.byref(o), C.byref(ph)), "upload")
    ms, nl = (C.c_double * 4)(), (C.c_int32 * 3)()
    for _ in range(3):
        ctx.check(ctx.L.bvio_batch_solve_timed(ctx.h, ph, ms, nl), "timed")
    out["kernel_ms_per_pass"] = dict(linearize=ms[0] / nl[0], solve=ms[1] / nl[1], cost=ms[2] / max(nl[2], 1))
    ctx.L.bvio_batch_free(ctx.h, ph)
    if oracle is not None:
        out.update(measure_cpu(pkg, oracle, L, strategy))
    return out


if __name__ == "__main__":
    import oracle_lib
    pkg = g.load_package()
    ctx = pkg.lib.Context(0)
    res = [measure(pkg, ctx, oracle_lib.load(), L) for L in (150, 1500)]
    ctx.close()
    print(json.dumps(res) if "--json" in sys.argv else "\n".join(str(r) for r in res))
