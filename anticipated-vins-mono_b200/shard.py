"""Host-side restatement of the multi-GPU selector protocol (csrc/sel_api.cu, csrc/sel_kernels.cu):
candidate partition, winner-record layout and the total order every rank applies to the gathered
records.  Used by the world_size > 1 CPU tests (gloo) and by bench.py's bookkeeping; the product
path itself runs these steps in CUDA.  Both device transports implement this same protocol: the fused
one (records stored straight into every rank's mailbox over peer memory inside the persistent
kernel; 32-byte header only, every rank holds all information blocks) and the NCCL one
(ncclAllGather of header + packed block between sel_round_kernel and sel_apply_kernel)."""
from __future__ import annotations

import numpy as np

REC_HDR = 4  # value, second, candidate index, prob   (SEL_REC_HDR in csrc/sel.h)


def shard_range(N: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous blocks of ceil(N/world) candidates per rank (sel_upload_impl)."""
    per = (N + world - 1) // world
    return min(N, rank * per), min(N, (rank + 1) * per)


def record_size(H: int) -> int:
    T = 3 * H
    return REC_HDR + T * (T + 1) // 2


def merge_best(best, second, idx, ob, os_, oidx):
    """merge_best() of sel_kernels.cu: order by value; an exact tie is the reference's `UBs[ub] = feature_id` collision
    (feature_selector.cpp:697,724): the larger candidate index takes the slot and the twin is not a runner-up."""
    if ob > best:
        return ob, max(best, second, os_), oidx
    if ob == best and oidx >= 0 and idx >= 0:
        return best, max(second, os_), max(idx, oidx)
    return best, max(second, ob, os_), idx


def pick_winner(records: np.ndarray):
    """sel_apply_kernel: records [world, record_size]; returns (value, second, idx, prob, C_packed)."""
    b, s, ix, who = -1.0, -np.inf, -1, -1
    for r in range(records.shape[0]):
        nb, ns, nix = merge_best(b, s, ix, records[r, 0], records[r, 1], int(records[r, 2]))
        if nix != ix:
            who = r
        b, s, ix = nb, ns, nix
    if ix < 0:
        return b, s, -1, 0.0, None
    return b, s, ix, records[who, 3], records[who, REC_HDR:]
