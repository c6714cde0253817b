"""world_size > 1 on CPU (gloo): the sharded-selector protocol -- contiguous candidate blocks per
rank, one all-gather of winner records per greedy round, identical update on every rank -- gives the
same selection as the single-process oracle for any world size.  Scoring here is numpy on the
oracle's compact blocks; the CUDA path implements the same protocol with NCCL (csrc/sel_api.cu)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed, N, H, kappa, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import __graft_entry__ as g
    import oracle_lib
    pkg = g.load_package()
    abi, synth, shard = pkg.abi, pkg.synth, pkg.shard
    orc = oracle_lib.load()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = synth.make_select_problem(seed=seed, N=N, H=H, kappa=kappa)
    h = abi.SelectHandle(p)
    T, D = 3 * H, 9 * (H + 1)
    c0, c1 = shard.shard_range(N, rank, world)
    Cc, valid, Om = np.zeros((N, T, T)), np.zeros(N, np.int32), np.zeros((D, D))
    orc.oracle_build_delta(C.byref(h.s), 0, None, abi.dptr(Cc), abi.iptr(valid), None)
    orc.oracle_omega_imu(C.byref(h.s), abi.dptr(Om))
    pos = np.array([9 * (1 + t // 3) + t % 3 for t in range(T)])
    oth = np.array([i for i in range(D) if i not in set(pos.tolist())])
    Moo, Mop, Mpp = Om[np.ix_(oth, oth)], Om[np.ix_(oth, pos)], Om[np.ix_(pos, pos)]
    R = Mpp - Mop.T @ np.linalg.solve(Moo, Mop)
    ld_oo = np.linalg.slogdet(Moo)[1]
    tril = np.tril_indices(T)
    taken = np.zeros(N, bool)
    RS = shard.record_size(H)
    sel = []
    for _ in range(kappa):
        best, second, bidx = -1.0, -np.inf, -1
        for i in range(c0, c1):                     # only this rank's shard is scored
            if not valid[i] or taken[i]:
                continue
            sign, ld = np.linalg.slogdet(R + p.cand_prob[i] * Cc[i])
            v = ld_oo + ld if sign > 0 else np.nan
            if v > best:
                second, best, bidx = best, v, i
            elif v > second:
                second = v
        rec = torch.zeros(RS, dtype=torch.float64)
        rec[0], rec[1], rec[2] = best, second, float(bidx)
        if bidx >= 0:
            rec[3] = p.cand_prob[bidx]
            rec[shard.REC_HDR:] = torch.from_numpy(Cc[bidx][tril])
        allr = [torch.zeros(RS, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allr, rec)                   # the one exchange per greedy round
        b, s, ix, prob, Cp = shard.pick_winner(torch.stack(allr).numpy())
        if ix >= 0:
            Cw = np.zeros((T, T))
            Cw[tril] = Cp
            Cw = Cw + np.tril(Cw, -1).T
            R = R + prob * Cw
            taken[ix] = True
            sel.append(int(p.cand_id[ix]))
    gathered = [None] * world
    dist.all_gather_object(gathered, sel)
    assert all(x == gathered[0] for x in gathered), "ranks disagree"
    if rank == 0:
        np.save(out_path, np.array(sel, np.int32))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_protocol_matches_oracle(pkg, oracle, tmp_path, world):
    abi, synth = pkg.abi, pkg.synth
    seed, N, H, kappa = 3, 45, 10, 9
    out = str(tmp_path / "sel.npy")
    mp.spawn(_worker, args=(world, _free_port(), seed, N, H, kappa, out), nprocs=world, join=True)
    got = np.load(out).tolist()
    p = synth.make_select_problem(seed=seed, N=N, H=H, kappa=kappa)
    h = abi.SelectHandle(p)
    ids, s = np.zeros(kappa, np.int32), abi.SelectSummary()
    assert oracle.oracle_select(C.byref(h.s), abi.iptr(ids), None, C.byref(s)) == 0
    assert got == ids[:s.n_selected].tolist()


def test_shard_ranges_cover_everything(pkg):
    shard = pkg.shard
    for N in (0, 1, 7, 2000, 2001):
        for world in (1, 2, 3, 4, 8):
            r = [shard.shard_range(N, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == N
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))


def test_pick_winner_total_order(pkg):
    shard = pkg.shard
    RS = shard.record_size(2)
    recs = np.zeros((3, RS))
    recs[:, 0], recs[:, 1], recs[:, 2] = [5.0, 7.0, 7.0], [4.0, 1.0, 6.5], [10, 30, 20]
    b, s, ix, prob, Cp = shard.pick_winner(recs)
    assert (b, ix) == (7.0, 30)        # exact tie = UB collision in the reference -> the larger candidate index
    assert s == 6.5                    # the collided twin is not a runner-up
    recs[:, 0], recs[:, 2] = -1.0, -1  # nobody has a candidate
    assert shard.pick_winner(recs)[2] == -1
