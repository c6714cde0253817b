"""Ground-truth-mode horizon generation, host side.

Restates HorizonGenerator::{loadGroundTruth, groundTruth, getNextFrameTruth}
(vins_estimator/src/utility/horizon_generator.cpp:74-123, 169-210): the EuRoC ground-truth csv reader
(`timestamp[ns], p, q(wxyz), v, bw, ba` per row, one header line) and the horizon built from relative ground-truth
motion between future frames.  It is sequential host bookkeeping (a seek cursor over the csv rows, H quaternion
products per frame): it feeds `bvio_select_in.horizon_pos / horizon_quat` the same way `bvio_horizon_imu` does for the
IMU mode, and is pinned against the reference's own class in tests/test_reference_pin.py.

Reference behaviour reproduced on purpose:
  * both time searches post-increment past the first row whose timestamp exceeds the target, and `getNextFrameTruth`
    returns the row AFTER that one (horizon_generator.cpp:85-87, 203-209);
  * the relative translation is rotated by the NEW frame's ground-truth attitude, `relP = q_next^-1 (p_next - p_prev)`
    (:110), then applied with the PREVIOUS horizon attitude (:116);
  * a first call with a timestamp beyond the end of the csv restarts from the first row (:82);
  * only position and attitude of the future states are produced (:114-117).
"""
from __future__ import annotations

import numpy as np

from . import synth as S


def load_groundtruth_csv(path):
    """-> dict(t [n] seconds, p [n,3], q [n,4] xyzw, v, w, a [n,3]).  Column order of
    benchmark_publisher/config/*/data.csv; the quaternion is stored w-first in the file."""
    rows = []
    with open(path) as f:
        next(f, None)                                   # header line (horizon_generator.cpp:176)
        for line in f:
            cells = line.rstrip("\n").split(",")
            if len(cells) < 17:
                continue
            rows.append([float(c) for c in cells[:17]])
    a = np.array(rows, dtype=np.float64).reshape(-1, 17)
    return dict(t=a[:, 0] * 1e-9, p=a[:, 1:4].copy(), q=np.column_stack([a[:, 5:8], a[:, 4]]), v=a[:, 8:11].copy(),
                w=a[:, 11:14].copy(), a=a[:, 14:17].copy())


def _conj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


class GroundTruthHorizon:
    def __init__(self, truth, H):
        self.truth, self.H = truth, int(H)
        self.seek_idx = 0

    def _next(self, idx, delta_frame):
        t = self.truth["t"]
        nxt = t[idx] + delta_frame
        while idx < len(t):
            idx += 1
            if not t[idx - 1] <= nxt:
                break
        if idx >= len(t):
            raise IndexError("ground truth exhausted (the reference reads past the end here)")
        return idx

    def generate(self, timestamp, pos0, quat0, delta_frame):
        """horizon_pos [H+1,3], horizon_quat [H+1,4] xyzw for the frame with the given timestamp / pose."""
        t = self.truth["t"]
        if timestamp > t[-1]:
            timestamp = t[0]
        while self.seek_idx < len(t):
            self.seek_idx += 1
            if not t[self.seek_idx - 1] <= timestamp:
                break
        idx = self.seek_idx - 1
        pos, quat = np.zeros((self.H + 1, 3)), np.zeros((self.H + 1, 4))
        pos[0], quat[0] = pos0, quat0
        prev_p, prev_q = self.truth["p"][idx], self.truth["q"][idx]
        for h in range(1, self.H + 1):
            idx = self._next(idx, delta_frame)
            gp, gq = self.truth["p"][idx], self.truth["q"][idx]
            rel_q = S.quat_mul(_conj(prev_q) / (prev_q @ prev_q), gq)
            rel_p = S._eigen_quat_rotate(_conj(gq) / (gq @ gq), gp - prev_p)
            pos[h] = pos[h - 1] + S._eigen_quat_rotate(quat[h - 1], rel_p)
            quat[h] = S.quat_mul(quat[h - 1], rel_q)
            prev_p, prev_q = gp, gq
        return pos, quat
