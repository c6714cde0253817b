#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native BA solver + anticipated feature selector.

  python bench.py --gpus N --steps K --warmup W            # our arm (libbvio.so, CUDA)
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

Metric (BASELINE.json): GN iterations/s on 11-keyframe / 1500-feature windows (+ candidates-scored/s
of the selector on 2000 candidates, H=10, kappa=150, reported under "selector").  A BA "step" is one
solve (8 iterations of the traditional dogleg the reference configures, estimator.cpp:798; tolerances off so
every window runs all 8) of a batch of independent windows that is larger than L2; every window of the pool is
the SECOND of two consecutive windows of its own synthetic session and carries the real n = 75 prior that
bvio_marginalize produced from the first.  A selector "step" is one full greedy selection.  Multi-GPU: BA =
independent replicas (no collective, weak scaling); selector = candidates sharded across ranks, winner
records exchanged once per greedy round over peer memory.

One JSON line on stdout (rank 0).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

K_FRAMES, L_FEATS, TRACK_MIN = 11, 1500, 6
SEL_N, SEL_H, SEL_KAPPA = 2000, 10, 150
POOL = 64                     # distinct synthetic windows (each with its own marginalized prior), tiled to the batch size
FP64_PEAK_TFLOPS = 36.4       # measured DFMA ceiling of this pool's B200 (tools/fp64_peak.cu; DMMA: 37.2)
BENCH_OPTS = dict(max_iters=8, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
PRIOR_KEYS = ("n", "block_kind", "block_frame", "block_idx", "x0", "lin_jac", "lin_res")
# the workload, identical for both arms (the driver compares the two lines' config)
CONFIG = {"workload": f"configs[2]: BA on {K_FRAMES}-keyframe / {L_FEATS}-feature windows (~10 500 projection factors), each the "
                      "second of two consecutive windows of its own session with the n = 75 prior marginalized from the first; "
                      "8 iterations of traditional dogleg per solve (DENSE_SCHUR + DOGLEG as estimator.cpp:794-806), tolerances off",
          "strategy": "dogleg", "iterations_per_solve": 8, "pool_windows": POOL, "prior": "marginalized, n = 75",
          "l2_policy": "inputs larger than L2: a GPU step solves 4 windows per SM (592 on a B200, ~0.7 GB resident); the "
                       "selector's 7 MB working set is L2-resident by design"}


def ba_algorithmic_bytes(w, n_prior, np_dim):
    """bytes one GN iteration of one window must move."""
    nf, L, K = w.n_factors, w.L, w.K
    return 20 * nf + 40 * L + 2296 * (K - 1) + 8 * n_prior * n_prior + 16 * n_prior + 128 * K + 8 * np_dim


def ba_linearize_flops(w):
    """Algorithmic FP64 flops of one linearization of one window: per projection factor
    ~600 (residual + Jacobians + Cauchy) + ~700 (its J^T J / J^T r blocks), per landmark the Schur outer product
    2 (6 n_l)^2.  The reduced solve (n_p^3 / 3) belongs to ba_solve and is not counted here."""
    nobs = np.diff(np.asarray(w.lm_obs_offset))
    return 1300.0 * w.n_factors + float((2.0 * (6.0 * nobs) ** 2).sum())


def sel_algorithmic_bytes_round(n_remaining, H):
    TT = (3 * H) * (3 * H + 1) // 2
    return n_remaining * (8 * TT + 8) + 8 * (9 * (H + 1)) ** 2


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def _session(i):
    import __graft_entry__ as gg
    return gg.load_package().synth.make_session(100 + i, K=K_FRAMES, L=L_FEATS, track_min=TRACK_MIN)


def make_sessions(n, workers):
    """n synthetic (K+1)-frame sessions (pure numpy; forked workers -- call before CUDA is initialised)."""
    g.load_package()                      # the parent unpickles the workers' Window objects: the class must be importable
    if workers <= 1:
        return [_session(i) for i in range(n)]
    import multiprocessing as mp
    with mp.get_context("fork").Pool(workers) as pool:
        return pool.map(_session, range(n))


def make_pool(pkg, sessions, solve, marginalize):
    """Every pool window = the second window of its session, started from the first window's solved state, with the
    prior marginalized from the first window (bvio_marginalize on the GPU arm, the oracle on the CPU arm)."""
    import dataclasses
    synth = pkg.synth
    pool = []
    for ses in sessions:
        first, idx = synth.slice_window(ses, 0, K_FRAMES)
        solved = solve(first)
        prior = marginalize(solved)
        pool.append(synth.consecutive_window(ses, solved, idx, {k: prior[k] for k in PRIOR_KEYS}, K=K_FRAMES))
    return pool


def pool_builders(pkg, ctx=None, orc=None):
    """(solve, marginalize) closures: through libbvio (ctx) or through the oracle (orc, CPU arm only)."""
    import dataclasses
    abi = pkg.abi
    o = abi.default_opts(strategy=1)

    def solve(w):
        h, sm = abi.WindowHandle(w), abi.Summary()
        if ctx is not None:
            ctx.check(ctx.L.bvio_optimize(ctx.h, C.byref(h.s), C.byref(o), C.byref(sm)), "bvio_optimize")
        else:
            assert orc.oracle_optimize(C.byref(h.s), C.byref(o), C.byref(sm)) == 0
        return dataclasses.replace(w, para_pose=h.pose, para_speed_bias=h.sb, inv_depth=h.inv)

    def marginalize(w):
        if ctx is not None:
            return abi.call_marginalize(ctx.L.bvio_marginalize, w, 0, ctx=ctx.h, opts=o)
        return abi.call_marginalize(orc.oracle_marginalize, w, 0, opts=o)
    return solve, marginalize


def window_array(abi, pool, B):
    hs = [abi.WindowHandle(pool[i % len(pool)]) for i in range(B)]
    arr = (abi.WindowS * B)(*[h.s for h in hs])
    return hs, arr


# --------------------------------------------------------------------------------------------
# CPU legs (oracle = CPU restatement of the reference's Ceres/Eigen path; the reference itself cannot
# be compiled in this image: no Eigen/Ceres/ROS -- DESIGN.md)
# --------------------------------------------------------------------------------------------
def cpu_ba(abi, orc, pool, n_solves, threads):
    o = abi.default_opts(strategy=1, **BENCH_OPTS)      # DOGLEG: what the reference configures (estimator.cpp:798)
    hs = [abi.WindowHandle(pool[i % len(pool)]) for i in range(n_solves)]
    sums = [abi.Summary() for _ in range(n_solves)]

    def run(i):
        orc.oracle_optimize(C.byref(hs[i].s), C.byref(o), C.byref(sums[i]))

    t0 = time.perf_counter()
    if threads <= 1:
        for i in range(n_solves):
            run(i)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:       # ctypes releases the GIL during the call
            list(ex.map(run, range(n_solves)))
    dt = time.perf_counter() - t0
    iters = sum(s.iterations for s in sums)
    return iters, dt


def cpu_select(abi, synth, orc, kappa):
    p = synth.make_select_problem(seed=0, N=SEL_N, H=SEL_H, kappa=kappa)
    h = abi.SelectHandle(p)
    ids = np.zeros(max(kappa, 1), np.int32)
    s = abi.SelectSummary()
    t0 = time.perf_counter()
    orc.oracle_select(C.byref(h.s), abi.iptr(ids), None, C.byref(s))
    return s.candidates_scored, time.perf_counter() - t0


def run_reference(args, rank, world):
    """The reference's path on the host cores: the CPU restatement (oracle/) of Ceres DENSE_SCHUR + DOGLEG on every host
    thread, same workload and strategy as the GPU arm; each step a bounded sample of the pool."""
    if rank != 0:
        return
    os.environ["ORACLE_EIGH"] = "ql"      # timed CPU legs: the eigen-solver class Eigen uses (oracle/ba_oracle.cpp)
    import oracle_lib
    pkg = g.load_package()
    abi, synth = pkg.abi, pkg.synth
    orc = oracle_lib.load()
    threads = os.cpu_count() or 1
    sessions = make_sessions(min(POOL, max(16, threads)), min(threads, 16))
    pool = make_pool(pkg, sessions, *pool_builders(pkg, orc=orc))
    per_step = 2 * threads
    for _ in range(args.warmup):
        cpu_ba(abi, orc, pool, threads, threads)
    iters, dt = 0, 0.0
    for _ in range(args.steps):
        i, d = cpu_ba(abi, orc, pool, per_step, threads)
        iters += i
        dt += d
    val = iters / dt
    sc, sdt = cpu_select(abi, synth, orc, 16)
    line = {
        "impl": "reference", "metric": "GN iters/sec (BA, 11-kf/1500-feat windows)", "value": val, "unit": "iters/s",
        "config": dict(CONFIG),
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "cpu_baseline": {"value": val, "unit": "iters/s", "cores": threads, "kind": "port",
                         "sample": f"{per_step} window solves per step x {args.steps} steps drawn from {len(pool)} pool windows; "
                                   "CPU restatement (oracle/, -O3) of the reference's Ceres DENSE_SCHUR + DOGLEG path -- the "
                                   "reference itself needs Eigen/Ceres/ROS and does not compile here"},
        "e2e": {"value": val, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "selector": {"metric": "candidates-scored/sec", "value": sc / sdt, "unit": "cand/s", "cores": 1,
                     "ms_per_select_extrapolated": 1e3 * sdt * SEL_KAPPA / 16,
                     "sample": f"N={SEL_N}, H={SEL_H}, first 16 greedy rounds, lazy-greedy dense {9*(SEL_H+1)}^2 LLT"},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
def time_resident_solve(ctx, abi, pool, B, opts, steps, warmup, stream, barrier, max_over_ranks, sum_over_ranks, timed=False):
    """B windows resident in HBM, `steps` solves timed with CUDA events on the library's stream.
    -> dict(ms_total, iters_total, launches, final_costs[, lin_ms, solve_ms, cost_ms])"""
    import torch
    L = ctx.L
    hs, arr = window_array(abi, pool, B)
    sums = (abi.Summary * B)()
    bh = C.c_void_p()
    ctx.check(L.bvio_batch_upload(ctx.h, arr, B, C.byref(opts), C.byref(bh)), "batch_upload")
    for _ in range(max(warmup, 3)):
        ctx.check(L.bvio_batch_solve(ctx.h, bh), "batch_solve")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = L.bvio_launch_count(ctx.h)
    e0.record(stream)
    for _ in range(steps):
        ctx.check(L.bvio_batch_solve(ctx.h, bh), "batch_solve")
    e1.record(stream)
    barrier()
    out = {"ms_total": max_over_ranks(e0.elapsed_time(e1)), "launches": L.bvio_launch_count(ctx.h) - l0}
    ctx.check(L.bvio_batch_download(ctx.h, bh, arr, sums), "batch_download")
    out["iters_total"] = sum_over_ranks(float(sum(s.iterations for s in sums))) * steps
    out["final_costs"] = [s.final_cost for s in sums[:4]]
    if timed:
        kms, kl = (C.c_double * 4)(), (C.c_int32 * 3)()
        for _ in range(2):
            ctx.check(L.bvio_batch_solve_timed(ctx.h, bh, kms, kl), "solve_timed")
        out.update(lin_ms=kms[0] / kl[0], solve_ms=kms[1] / kl[1], cost_ms=kms[2] / max(kl[2], 1))
    L.bvio_batch_free(ctx.h, bh)
    return out


def time_selector(ctx, abi, synth, N, H, kappa, steps, warmup, stream, barrier, max_over_ranks, sharded):
    """One resident selection problem: ms per select (CUDA events, max over ranks), summary, ids."""
    import torch
    L = ctx.L
    p = synth.make_select_problem(seed=0, N=N, H=H, kappa=kappa)
    sh = abi.SelectHandle(p)
    ph = C.c_void_p()
    ctx.check(L.bvio_select_upload_mode(ctx.h, C.byref(sh.s), 1 if sharded else 0, C.byref(ph)), "select_upload")
    for _ in range(max(warmup, 2)):
        ctx.check(L.bvio_select_run(ctx.h, ph), "select_run")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l1 = L.bvio_launch_count(ctx.h)
    e0.record(stream)
    for _ in range(steps):
        ctx.check(L.bvio_select_run(ctx.h, ph), "select_run")
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    launches = L.bvio_launch_count(ctx.h) - l1
    ids, ss = np.zeros(kappa, np.int32), abi.SelectSummary()
    ctx.check(L.bvio_select_fetch(ctx.h, ph, abi.iptr(ids), None, C.byref(ss)), "select_fetch")
    L.bvio_select_free(ctx.h, ph)
    return {"ms": ms, "launches": launches, "summary": ss, "ids": ids, "handle": sh, "problem": p}


TRANSPORT = {0: "single-gpu", 1: "fused (peer-memory mailboxes inside the persistent kernel)", 2: "nccl (one ncclAllGather per round)"}


def selector_block(r):
    ss = r["summary"]
    return {"ms_per_select": r["ms"], "transport": TRANSPORT.get(ss.transport, "?"), "world": ss.world, "grid": ss.grid, "cpw": ss.cpw,
            "per_round_us": {"score_merge_update": ss.round_score_us, "grid_barrier": ss.round_barrier_us,
                             "exchange": ss.round_exchange_us}}


def run_stream(pkg, ctx, abi, frames, max_feats, H, label, track_loss=0.0, overlap=False):
    """BASELINE configs[4]: closed-loop sequence of CONSECUTIVE sliding windows (slider.py), one window per call through the
    one-shot C-ABI with host buffers.  *_call = the FFI call alone; the others include the ctypes packing around it."""
    sim = pkg.slider.SlidingWindowSim(seed=7, max_feats=max_feats, max_cand=300, H=H, frame_dt=1.0 / 30.0)
    gb = pkg.slider.GpuBackend(ctx, abi)          # the reference's budget: 8 iterations, Ceres default tolerances, dogleg
    sim.opts = dict(strategy=1)
    sim.track_loss = track_loss                   # fraction of live tracks the front end loses per frame
    sim.overlap_marginalize = overlap             # bvio_marginalize_begin ... select ... bvio_marginalize_end
    keys = ("optimize", "marginalize", "select", "optimize_call", "marginalize_call", "select_call")
    lat = {k: [] for k in keys + ("frame", "frame_call")}
    cnt = {"L": [], "n_factors": [], "N": [], "kappa": [], "iterations": []}
    warm = sim.K + 4
    t_wall = time.perf_counter()
    for f in range(frames + warm):
        r = sim.step(gb)
        if r is None or f < warm:
            continue
        for k in keys:
            lat[k].append(r[k])
        lat["frame"].append(r["optimize"] + r["marginalize"] + r["select"])
        lat["frame_call"].append(r["optimize_call"] + r["marginalize_call"] + r["select_call"])
        for k in cnt:
            cnt[k].append(r.get(k, 0))
    errs = np.array([h[1] for h in sim.history])
    out = {"frames": frames, "label": label, "wall_s": time.perf_counter() - t_wall, "track_loss_per_frame": track_loss,
           "marginalization": "bvio_marginalize_begin / _end around bvio_select (second stream)" if overlap else "bvio_marginalize (synchronous)",
           "workload": f"closed-loop 30 Hz sequence of consecutive 11-keyframe windows, budget {max_feats} features: per frame "
                       "bvio_optimize (8 dogleg iterations, Ceres default tolerances, prior = previous marginalization) + "
                       f"bvio_marginalize (eigen route) + bvio_select (H={H}), host buffers, wall clock",
           "budget_ms_30hz": 33.3,
           "mean_landmarks": float(np.mean(cnt["L"])), "mean_factors": float(np.mean(cnt["n_factors"])),
           "mean_candidates": float(np.mean(cnt["N"])), "mean_kappa": float(np.mean(cnt["kappa"])),
           "mean_iterations": float(np.mean(cnt["iterations"])),
           "position_error_m": {"mean": float(errs.mean()), "last": float(errs[-1])}}
    for k, v in lat.items():
        a = np.array(v) * 1e3
        out[k + "_ms"] = {"p50": float(np.percentile(a, 50)), "p99": float(np.percentile(a, 99)), "max": float(a.max())}
    return out


def cpu_stream(pkg, orc, abi, frames, max_feats, H, track_loss=0.0):
    """The same closed loop on the CPU oracle (cpu_baseline leg): per-frame latency of optimize + marginalize + select."""
    from slider_backends import OracleBackend
    sim = pkg.slider.SlidingWindowSim(seed=7, max_feats=max_feats, max_cand=300, H=H, frame_dt=1.0 / 30.0)
    sim.opts = dict(strategy=1)
    sim.track_loss = track_loss
    be = OracleBackend(orc, abi)
    lat, warm = [], sim.K + 2
    for f in range(frames + warm):
        r = sim.step(be)
        if r is not None and f >= warm:
            lat.append((r["optimize"], r["marginalize"], r["select"]))
    a = np.array(lat) * 1e3
    return {"frames": len(lat), "optimize_ms_p50": float(np.percentile(a[:, 0], 50)), "marginalize_ms_p50": float(np.percentile(a[:, 1], 50)),
            "select_ms_p50": float(np.percentile(a[:, 2], 50)), "frame_ms_p50": float(np.percentile(a.sum(1), 50))}


def run_ours(args, rank, world, local_rank):
    # synthetic sessions first: forked numpy workers must not inherit a CUDA context
    t_pool = time.perf_counter()
    sessions = make_sessions(args.pool, max(1, min(16, (os.cpu_count() or 1) // max(world, 1))))
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = g.load_package()
    abi, synth = pkg.abi, pkg.synth
    ctx = pkg.lib.Context(local_rank)
    L = ctx.L
    stream = torch.cuda.ExternalStream(L.bvio_stream(ctx.h), device=torch.device("cuda", local_rank))
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
    B = args.batch if args.batch > 0 else 4 * n_sm    # whole waves of windows: 4 per SM (592 on a B200)
    pool = make_pool(pkg, sessions, *pool_builders(pkg, ctx=ctx))
    t_pool = time.perf_counter() - t_pool
    o_dl = abi.default_opts(strategy=1, **BENCH_OPTS)  # headline: the reference's strategy
    o_lm = abi.default_opts(strategy=0, **BENCH_OPTS)

    # ---- multi-GPU plumbing for the sharded selector
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (C.c_char * 128)()
            assert L.bvio_nccl_unique_id(buf) == 0
            uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        raw = bytes(uid.cpu().numpy().tobytes())
        ctx.check(L.bvio_comm_init(ctx.h, raw, rank, world), "bvio_comm_init")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())
    tools = (stream, barrier, max_over_ranks, sum_over_ranks)

    # ---- BA, inputs resident in HBM: headline (dogleg) with the per-kernel split, LM beside it
    sampler = ClockSampler(local_rank)
    sampler.start()
    dl = time_resident_solve(ctx, abi, pool, B, o_dl, args.steps, args.warmup, *tools, timed=True)
    lm = time_resident_solve(ctx, abi, pool, B, o_lm, max(1, min(args.steps, 5)), 3, *tools)
    ba_value = dl["iters_total"] / (dl["ms_total"] * 1e-3)
    lin_ms, solve_ms, cost_ms = dl["lin_ms"], dl["solve_ms"], dl["cost_ms"]
    n_prior = pool[0].prior["n"]
    alg_bytes_iter = sum(ba_algorithmic_bytes(pool[i % len(pool)], pool[i % len(pool)].prior["n"], 15 * K_FRAMES) for i in range(B))
    lin_bytes = alg_bytes_iter - 8 * 15 * K_FRAMES * B      # everything but the delta-x write is read by linearize
    lin_flops = sum(ba_linearize_flops(pool[i % len(pool)]) for i in range(B))

    # ---- selector, inputs resident in HBM: config 4 (H = 10) and the reference's compile-time horizon (H = 13)
    sel = time_selector(ctx, abi, synth, SEL_N, SEL_H, SEL_KAPPA, args.steps, args.warmup, stream, barrier, max_over_ranks, world > 1)
    clocks = sampler.stop()
    sel13 = time_selector(ctx, abi, synth, SEL_N, 13, SEL_KAPPA, max(1, min(args.steps, 5)), 2, stream, barrier, max_over_ranks, world > 1)
    ss, ids, sh = sel["summary"], sel["ids"], sel["handle"]
    sel_value = ss.candidates_scored / (sel["ms"] * 1e-3)
    nv = ss.n_candidates_valid
    sel_bytes = sum(sel_algorithmic_bytes_round(nv - i, SEL_H) for i in range(ss.n_selected))
    # strong scaling of the sharded selector where sharding has something to share (multi-GPU runs only)
    sel_scaling = None
    if world > 1:
        sel_scaling = []
        for N_, H_ in ((2000, 10), (16000, 10), (64000, 10), (16000, 13), (64000, 13)):
            one = time_selector(ctx, abi, synth, N_, H_, SEL_KAPPA, 2, 1, stream, barrier, max_over_ranks, False)
            shd = time_selector(ctx, abi, synth, N_, H_, SEL_KAPPA, 2, 1, stream, barrier, max_over_ranks, True)
            assert (one["ids"] == shd["ids"]).all(), f"sharded selector disagrees with one GPU at N={N_} H={H_}"
            sel_scaling.append({"N": N_, "H": H_, "kappa": SEL_KAPPA, "one_gpu_ms": one["ms"], "sharded_ms": shd["ms"],
                                "speedup": one["ms"] / shd["ms"], "efficiency": one["ms"] / shd["ms"] / world,
                                "one_gpu": selector_block(one), "sharded": selector_block(shd)})

    # ---- end to end through the public C-ABI with HOST buffers (H2D + D2H inside the timed region)
    e2e_steps = max(1, min(args.steps, 5))
    sums = (abi.Summary * B)()
    hs2, arr2 = window_array(abi, pool, B)
    state0 = [(h.pose.copy(), h.sb.copy(), h.inv.copy()) for h in hs2]
    for _ in range(2):
        ctx.check(L.bvio_optimize_batch(ctx.h, arr2, B, C.byref(o_dl), sums), "optimize_batch")
    barrier()
    t_e2e, it_e2e = 0.0, 0
    for _ in range(e2e_steps):
        for h, (a, b_, c) in zip(hs2, state0):     # restore inputs (not timed)
            h.pose[...] = a
            h.sb[...] = b_
            h.inv[...] = c
        t0 = time.perf_counter()
        ctx.check(L.bvio_optimize_batch(ctx.h, arr2, B, C.byref(o_dl), sums), "optimize_batch")
        t_e2e += time.perf_counter() - t0
        it_e2e += sum(s.iterations for s in sums)
    t_e2e = max_over_ranks(t_e2e)
    e2e_value = sum_over_ranks(float(it_e2e)) / t_e2e
    h2d = sum(h.pose.nbytes + h.sb.nbytes + h.ex.nbytes + h.inv.nbytes + h.off.nbytes + h.frame.nbytes +
              h.xy.nbytes + h.pre.nbytes + (h._pj.nbytes + h._pr.nbytes + h._px0.nbytes if h.prior_s else 0) for h in hs2)
    d2h = sum(h.pose.nbytes + h.sb.nbytes + h.inv.nbytes for h in hs2) + B * 104
    del hs2, arr2, state0
    # selector end to end
    ids2 = np.zeros(SEL_KAPPA, np.int32)
    ss2 = abi.SelectSummary()
    sel_fn = L.bvio_select_sharded if world > 1 else L.bvio_select
    ctx.check(sel_fn(ctx.h, C.byref(sh.s), abi.iptr(ids2), None, C.byref(ss2)), "select")
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.check(sel_fn(ctx.h, C.byref(sh.s), abi.iptr(ids2), None, C.byref(ss2)), "select")
    t_sel_e2e = max_over_ranks(time.perf_counter() - t0)
    sel_e2e = ss2.candidates_scored * e2e_steps / t_sel_e2e
    sel_h2d = sh.pos.nbytes + sh.quat.nbytes + sh.cxy.nbytes + sh.cp.nbytes + sh.clxy.nbytes + sh.cld.nbytes
    assert (ids2 == ids).all(), "one-shot and resident selector disagree"

    # ---- the reference's actual use: ONE window per call (configs 2 and 5), rank 0 only
    single, streams = None, None
    if rank == 0 and not args.no_latency:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import single_window
        single = [single_window.measure(pkg, ctx, None, L_) for L_ in (150, 1500)]
    if rank == 0 and args.stream_frames > 0:
        streams = [run_stream(pkg, ctx, abi, args.stream_frames, 150, SEL_H, "configs[4]: feature budget 150"),
                   run_stream(pkg, ctx, abi, min(args.stream_frames, 2000), 150, 13,
                              "feature budget 150, 20 % of the tracks lost per frame (kappa ~ 30 new features per frame), H = 13 "
                              "(the reference's compile-time HORIZON, state_defs.h:8)", track_loss=0.2),
                   run_stream(pkg, ctx, abi, min(args.stream_frames, 2000), 150, SEL_H,
                              "configs[4] with the marginalization overlapped with select()", overlap=True)]

    # ---- CPU baseline on the host cores (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import oracle_lib
        orc = oracle_lib.load()
        os.environ["ORACLE_EIGH"] = "ql"  # timed CPU legs: the eigen-solver class Eigen uses (oracle/ba_oracle.cpp)
        it, dt = cpu_ba(abi, orc, pool, args.cpu_solves, 1)
        sc, sdt = cpu_select(abi, synth, orc, args.cpu_kappa)
        cpu = {"value": it / dt, "unit": "iters/s", "cores": 1, "kind": "port",
               "sample": f"{args.cpu_solves} solves drawn from the same pool ({K_FRAMES}-kf/{L_FEATS}-feat windows, n = 75 priors), "
                         f"8 dogleg iterations each, 1 thread ({dt:.1f} s); selector: first {args.cpu_kappa} rounds of "
                         f"N={SEL_N} ({sdt:.1f} s)",
               "selector_value": sc / sdt, "selector_unit": "cand/s", "selector_ms_per_select_extrapolated": 1e3 * sdt * SEL_KAPPA / args.cpu_kappa}
        if single is not None:
            import single_window
            for blk in single:
                ref = single_window.measure_cpu(pkg, orc, blk["L"])
                blk.update(cpu_oracle_ms=ref["cpu_oracle_ms"], cpu_iterations=ref["cpu_iterations"])
        if streams is not None:
            streams[0]["cpu_oracle"] = cpu_stream(pkg, orc, abi, 40, 150, SEL_H)
            streams[1]["cpu_oracle"] = cpu_stream(pkg, orc, abi, 40, 150, 13, track_loss=0.2)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("hbm_gbs", 6650.0)
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            # the linearization of a throughput batch is two launches: the visual tiles and the IMU / prior companion
            traffic = tj["ba_linearize_ws_kernel"] + tj.get("ba_imu_prior_kernel", 0)
        except Exception:
            pass
        achieved = lin_bytes / (lin_ms * 1e-3) / 1e9
        cfg = dict(CONFIG)
        cfg["selector"] = selector_block(sel)        # compact, up front: survives a truncated tail
        line = {
            "metric": "GN iters/sec (BA, 11-kf/1500-feat windows)", "value": ba_value, "unit": "iters/s",
            "config": cfg,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dl["ms_total"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "e2e": {"value": e2e_value, "unit": "iters/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "bvio_optimize_batch (host buffers, pack + H2D + solve + D2H), dogleg"},
            "gpu_launches": int(dl["launches"] + sel["launches"]),
            "clocks": clocks,
            "run": {"batch_windows_per_gpu": B, "mean_factors_per_window": float(np.mean([w.n_factors for w in pool])),
                    "mean_landmarks_per_window": float(np.mean([w.L for w in pool])), "prior_n": int(n_prior),
                    "pool_build_s": t_pool, "parallelism": "replicas (BA) + candidate-sharded selector" if world > 1 else "single GPU",
                    "final_cost_check": dl["final_costs"]},
            "roofline": {"bound": "hbm", "kernel": "ba_linearize_ws_kernel (+ ba_imu_prior_kernel, its companion launch)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650",
                         "algorithmic_bytes_per_launch": int(lin_bytes), "launch_ms": lin_ms,
                         "kernel_ms_per_pass": {"ba_linearize_ws_kernel + ba_imu_prior_kernel": lin_ms, "ba_solve_kernel (+ ba_dogleg_kernel)": solve_ms,
                                                "ba_cost_kernel": cost_ms},
                         "fp64": {"achieved": lin_flops / (lin_ms * 1e-3) / 1e12, "peak": FP64_PEAK_TFLOPS,
                                  "unit": "TFLOP/s", "frac": lin_flops / (lin_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                                  "peak_source": "tools/fp64_peak.cu on this pool's B200: DFMA 36.4 / DMMA 37.2 TFLOP/s "
                                                 "(profiles/r01_fp64_peak.json)",
                                  "algorithmic_flops_per_launch": lin_flops},
                         "note": "FP64 compute/latency-bound path (~95 flop/byte): the HBM fraction is reported as "
                                 "BASELINE.json asks; the binding ceiling is the FP64 pipe, reported under fp64"},
            "lm": {"value": lm["iters_total"] / (lm["ms_total"] * 1e-3), "unit": "iters/s",
                   "note": "same windows, BVIO_STRATEGY_LM (north_star's 'LM damping loop'): 3 launches per iteration instead of 4"},
            "selector": {"metric": "candidates-scored/sec", "value": sel_value, "unit": "cand/s",
                         "ms_per_select": sel["ms"], "workload": f"configs[3]: N={SEL_N}, H={SEL_H}, "
                         f"kappa={SEL_KAPPA}, {nv} valid candidates, every remaining candidate scored each round",
                         "scaling": "strong" if world > 1 else None,
                         "l2_resident_gbs": sel_bytes / (sel["ms"] * 1e-3) / 1e9,
                         "h13": dict(selector_block(sel13), workload=f"N={SEL_N}, H=13 (the reference's compile-time HORIZON), kappa={SEL_KAPPA}"),
                         "e2e": {"ms_per_select": 1e3 * t_sel_e2e / e2e_steps, "value": sel_e2e, "unit": "cand/s",
                                 "h2d_bytes_per_step": int(sel_h2d), "d2h_bytes_per_step": int(SEL_KAPPA * 12 + 64)},
                         "strong_scaling": sel_scaling,
                         "selected_head": ids[:8].tolist()},
            "single_window": single,
            "stream": streams,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="independent windows per GPU per BA step (0 = 4 per SM)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-solves", type=int, default=64)
    ap.add_argument("--cpu-kappa", type=int, default=16)
    ap.add_argument("--stream-frames", type=int, default=10000,
                    help="frames of the closed-loop latency leg (0 = skip; BASELINE configs[4] asks for 10000)")
    ap.add_argument("--no-latency", action="store_true", help="skip the single-window latency leg")
    ap.add_argument("--pool", type=int, default=POOL, help="distinct windows in the pool (profiling runs use a small one)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
