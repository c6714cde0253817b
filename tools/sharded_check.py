#!/usr/bin/env python
"""Run under torchrun (one process per GPU): bvio_select_sharded over NCCL must return exactly what
the single-GPU bvio_select returns on every rank (ids and log-det values bit-identical)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    pkg = g.load_package()
    abi, synth = pkg.abi, pkg.synth
    ctx = pkg.lib.Context(lr)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (C.c_char * 128)()
        assert ctx.L.bvio_nccl_unique_id(buf) == 0
        uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    ctx.check(ctx.L.bvio_comm_init(ctx.h, bytes(uid.cpu().numpy().tobytes()), rank, world), "comm_init")
    ok = True
    # BVIO_SEL_FORCE_SHARD=1 (set by the test): share even the small problems that the planner would keep on one GPU
    for seed, N, H, U, kappa, twins in ((0, 2000, 10, 0, 150, 0), (1, 333, 13, 4, 40, 0), (2, 7, 10, 0, 5, 0), (3, 64, 5, 2, 64, 0),
                                        (4, 400, 10, 0, 60, 12), (5, 240, 13, 3, 30, 8), (6, 12000, 10, 0, 30, 0), (7, 9000, 13, 2, 20, 6)):
        p = synth.make_select_problem(seed=seed, N=N, H=H, U=U, kappa=kappa)
        # exact duplicates (the reference's UB-map collision, feature_selector.cpp:697,724): twins inside one shard and
        # twins that straddle shard boundaries must resolve to the larger id on every rank
        rng = np.random.default_rng(seed)
        for k in range(twins):
            a, b = (2 * k, 2 * k + 1) if k % 2 == 0 else sorted(rng.choice(N, 2, replace=False).tolist())
            p.cand_xy[b], p.cand_prob[b] = p.cand_xy[a], p.cand_prob[a]
        if twins:
            p.state_k1_pos = p.horizon_pos[1] + 0.03          # ABI v2 inputs ride along on the sharded path
            p.omega_prior = np.diag(np.linspace(1.0, 9.0, 9))
        h = abi.SelectHandle(p)
        i1, v1, s1 = np.full(kappa, -1, np.int32), np.zeros(kappa), abi.SelectSummary()
        i2, v2, s2 = np.full(kappa, -1, np.int32), np.zeros(kappa), abi.SelectSummary()
        ctx.check(ctx.L.bvio_select(ctx.h, C.byref(h.s), abi.iptr(i1), abi.dptr(v1), C.byref(s1)), "select")
        ctx.check(ctx.L.bvio_select_sharded(ctx.h, C.byref(h.s), abi.iptr(i2), abi.dptr(v2), C.byref(s2)), "select_sharded")
        same = (np.array_equal(i1, i2) and np.array_equal(v1, v2) and s1.n_selected == s2.n_selected and
                s1.candidates_scored == s2.candidates_scored and s1.n_candidates_valid == s2.n_candidates_valid and
                s1.final_logdet == s2.final_logdet)
        print(f"[rank {rank}/{world}] seed {seed} N {N} H {H} kappa {kappa}: n_sel {s2.n_selected} scored {s2.candidates_scored} "
              f"single {s1.device_ms:.2f} ms (cpw {s1.cpw}) sharded {s2.device_ms:.2f} ms (transport {s2.transport}, cpw {s2.cpw}, "
              f"round us {s2.round_score_us:.1f}/{s2.round_barrier_us:.1f}/{s2.round_exchange_us:.1f})  {'OK' if same else 'MISMATCH'}", flush=True)
        ok &= same
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ctx.close()
    dist.destroy_process_group()
    if not int(t.item()):
        raise SystemExit(1)
    if rank == 0:
        print("SHARDED_OK")


if __name__ == "__main__":
    main()
