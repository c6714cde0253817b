"""Closed-loop sliding-window driver (BASELINE configs[4]).

Host-side restatement of the bookkeeping that surrounds the hot path in the reference, so that consecutive
windows -- each one's prior produced by the previous one's marginalization -- can be streamed through the
C-ABI without ROS:

  * Estimator::processIMU / processImage skeleton (vins_estimator/src/estimator.cpp:86-186): IMU-predicted state
    of the incoming frame, preintegration with the newest bias estimate;
  * FeatureManager::{addFeatureCheckParallax, triangulate, setDepth, removeFailures, removeBackShiftDepth}
    (vins_estimator/src/feature_manager.cpp:46-113, 141-159, 202-257, 275-311);
  * the filter that decides which features become parameter blocks (used_num >= 2 && start_frame < WINDOW_SIZE - 2,
    estimator.cpp:712-718, feature_manager.cpp:28-43);
  * Estimator::double2vector's yaw / position gauge re-anchoring (estimator.cpp:521-555);
  * Estimator::slideWindow, both branches (estimator.cpp:996-1078): MARGIN_OLD with removeBackShiftDepth, and
    MARGIN_SECOND_NEW (the second-newest frame is not a keyframe: its IMU samples are appended to the previous
    preintegration, the newest frame takes its slot, FeatureManager::removeFront, feature_manager.cpp:330-351);
  * the new-feature budget of FeatureSelector::select (feature_selector.cpp:157-170): kappa = maxFeatures - tracked.

The arithmetic of optimize / marginalize / select is NOT here: it is delegated to a backend (the CUDA library in the
product; the CPU oracle only in tests).  `keyframes="always"` treats every frame as a keyframe (MARGIN_OLD, the
stream benchmark's setting); `keyframes="parallax"` applies the reference's addFeatureCheckParallax rule.  Pure numpy.
"""
from __future__ import annotations

import dataclasses
import time

import numpy as np

from . import synth as S

WINDOW_SIZE = 10           # parameters.h:15 ; K = WINDOW_SIZE + 1 frames in the window
INIT_DEPTH = 5.0           # parameters.cpp:3
MIN_PARALLAX = 10.0 / S.FOCAL_LENGTH   # keyframe_parallax 10 px (config/euroc/euroc_config.yaml) / FOCAL_LENGTH, parameters.cpp:88
MARGIN_OLD, MARGIN_SECOND_NEW = 0, 1


@dataclasses.dataclass
class Track:
    lid: int                       # world landmark id (= feature id)
    start: int                     # index of the first observation's frame inside the window
    xy: list                       # normalized-plane observations, one per consecutive frame from `start`
    depth: float = -1.0            # estimated depth in the anchor frame (<= 0: not triangulated yet)
    alive: bool = True             # still tracked in the newest frame


def _R(q):
    return S.quat_to_rot(q)


def _ypr(R):
    """Utility::R2ypr (vins_estimator/src/utility/utility.h:69-85), degrees."""
    n, o, a = R[:, 0], R[:, 1], R[:, 2]
    y = np.arctan2(n[1], n[0])
    p = np.arctan2(-n[2], n[0] * np.cos(y) + n[1] * np.sin(y))
    r = np.arctan2(a[0] * np.sin(y) - a[1] * np.cos(y), -o[0] * np.sin(y) + o[1] * np.cos(y))
    return np.array([y, p, r]) / np.pi * 180.0


def _yaw_R(deg):
    y = deg / 180.0 * np.pi
    return np.array([[np.cos(y), -np.sin(y), 0.0], [np.sin(y), np.cos(y), 0.0], [0.0, 0.0, 1.0]])


def regauge(pose0_before, pose, sb):
    """Estimator::double2vector (estimator.cpp:521-555): rotate the solution about z so that frame 0 keeps its
    pre-solve yaw, and translate it so that frame 0 keeps its pre-solve position."""
    R0, P0 = _R(pose0_before[3:]), pose0_before[:3].copy()
    R00 = _R(pose[0, 3:])
    a, b = _ypr(R0), _ypr(R00)
    rot = _yaw_R(a[0] - b[0])
    if abs(abs(a[1]) - 90) < 1.0 or abs(abs(b[1]) - 90) < 1.0:
        rot = R0 @ R00.T
    p00 = pose[0, :3].copy()
    for i in range(len(pose)):
        Ri = rot @ _R(pose[i, 3:] / np.linalg.norm(pose[i, 3:]))
        pose[i, :3] = rot @ (pose[i, :3] - p00) + P0
        pose[i, 3:] = S.rot_to_quat(Ri)
        sb[i, :3] = rot @ sb[i, :3]


def triangulate(track, pose, ric, tic):
    """FeatureManager::triangulate (feature_manager.cpp:202-257): DLT over all observations, relative to the
    anchor camera frame; depth = svd_V[2] / svd_V[3]; below 0.1 falls back to INIT_DEPTH."""
    i0 = track.start
    R0c = _R(pose[i0, 3:]) @ ric
    t0 = pose[i0, :3] + _R(pose[i0, 3:]) @ tic
    rows = []
    for k, xy in enumerate(track.xy):
        j = i0 + k
        Rjc = _R(pose[j, 3:]) @ ric
        tj = pose[j, :3] + _R(pose[j, 3:]) @ tic
        t = R0c.T @ (tj - t0)
        Rr = R0c.T @ Rjc
        P = np.hstack([Rr.T, (-Rr.T @ t)[:, None]])
        f = np.array([xy[0], xy[1], 1.0])
        f = f / np.linalg.norm(f)
        rows.append(f[0] * P[2] - f[2] * P[0])
        rows.append(f[1] * P[2] - f[2] * P[1])
    A = np.array(rows)
    v = np.linalg.svd(A, full_matrices=True)[2][-1]
    d = v[2] / v[3] if v[3] != 0 else -1.0
    return d if d >= 0.1 else INIT_DEPTH


class World:
    """Static landmark field around the analytic trajectory of synth.Trajectory."""

    def __init__(self, rng, n=40000):
        self.pts = np.column_stack([rng.uniform(-12, 12, n), rng.uniform(-12, 12, n), rng.uniform(-3, 4, n)])
        self.score = rng.uniform(0.05, 1.0, n)          # GFTT score / max score (feature_tracker.cpp:313-322)

    def observe(self, cam, Rwc, twc, zmin=0.5, zmax=15.0):
        """ids and normalized-plane coordinates of the landmarks that project inside the image."""
        pc = (self.pts - twc) @ Rwc                     # rows: Rwc^T (p - t)
        z = pc[:, 2]
        ok = (z > zmin) & (z < zmax)
        mx, my = pc[:, 0] / np.where(ok, z, 1.0), pc[:, 1] / np.where(ok, z, 1.0)
        dx, dy = S.distortion(cam, mx, my)
        u = np.round(cam["fx"] * (mx + dx) + cam["cx"])
        v = np.round(cam["fy"] * (my + dy) + cam["cy"])
        ok &= (u >= 0) & (u < cam["width"]) & (v >= 0) & (v < cam["height"]) & (np.abs(mx) < 1.2) & (np.abs(my) < 0.9)
        ids = np.nonzero(ok)[0]
        return ids, np.column_stack([mx[ids], my[ids]]), z[ids]


class SlidingWindowSim:
    """One simulated VIO session.  `step(backend)` ingests one camera frame and, once the window is full, runs
    optimize -> (regauge, setDepth, removeFailures) -> marginalize -> select -> slideWindow, returning the
    wall time of each backend call."""

    def __init__(self, seed=0, max_feats=150, max_cand=300, H=10, frame_dt=0.1, imu_rate=200, px_sigma=0.5,
                 opts=None, keyframes="always", speed=1.0):
        self.rng = np.random.default_rng(seed)
        self.K = WINDOW_SIZE + 1
        self.cam = S.EUROC_CAM
        U_, _, Vt_ = np.linalg.svd(S.EUROC_RIC)
        self.ric, self.tic = U_ @ Vt_, S.EUROC_TIC.copy()
        self.qic = S.rot_to_quat(self.ric)
        self.traj = S.Trajectory(phase=self.rng.uniform(0, 10.0), scale=speed)
        assert keyframes in ("always", "parallax")
        self.keyframes = keyframes
        self.margin_flag = MARGIN_OLD
        self.sum_of_back = self.sum_of_front = 0       # estimator.h:89 counters
        self.world = World(self.rng)
        self.max_feats, self.max_cand, self.H = max_feats, max_cand, H
        self.frame_dt, self.n_imu = frame_dt, int(round(frame_dt * imu_rate))
        self.sig = px_sigma / S.FOCAL_LENGTH
        self.ba_true = self.rng.normal(0, 0.02, 3)
        self.bg_true = self.rng.normal(0, 0.002, 3)
        self.g = np.array([0, 0, S.G_NORM])
        self.opts = opts or {}
        self.t = 1.0                                   # time of the newest frame
        self.frame = 0                                 # frames ingested so far
        # window state (lists grow to K entries)
        self.pose = np.zeros((0, 7))
        self.sb = np.zeros((0, 9))
        self.preint = [np.zeros(S.PREINT_DOUBLES)]
        self.pre_obj = [None]                          # the live IntegrationBase of every interval ...
        self.imu_buf = [[]]                            # ... and its samples (dt_buf / *_buf, estimator.h:79-81)
        self.tracks: dict[int, Track] = {}
        self.prior = None
        self.last_selected = np.zeros(0, np.int32)
        self.history = []                              # (frame, position error, cost) per optimised window

    # ---- sensors -------------------------------------------------------------------------------------------
    def _imu(self, t):
        R = self.traj.rot(t)
        acc = R.T @ (self.traj.acc(t) + self.g) + self.ba_true + self.rng.normal(0, S.ACC_N, 3)
        gyr = self.traj.omega_body(t) + self.bg_true + self.rng.normal(0, S.GYR_N, 3)
        return acc, gyr

    def _gt(self, t):
        return np.concatenate([self.traj.pos(t), S.rot_to_quat(self.traj.rot(t))]), \
            np.concatenate([self.traj.vel(t), self.ba_true, self.bg_true])

    def _cam_pose(self, pose):
        R = _R(pose[3:])
        return R @ self.ric, pose[:3] + R @ self.tic

    # ---- one frame -----------------------------------------------------------------------------------------
    def _ingest(self):
        """processIMU + the tracking half of processImage for the next frame."""
        K = self.K
        first = len(self.pose) == 0
        if first:
            p, sbv = self._gt(self.t)
            sbv[3:6] += self.rng.normal(0, 0.01, 3)
            sbv[6:9] += self.rng.normal(0, 0.001, 3)
            self.pose, self.sb = p[None, :].copy(), sbv[None, :].copy()
        else:
            t0, dt = self.t, self.frame_dt / self.n_imu
            ba, bg = self.sb[-1, 3:6].copy(), self.sb[-1, 6:9].copy()
            a0, g0 = self._imu(t0)
            pre = S.Preintegration(a0, g0, ba, bg)
            buf = []
            for i in range(1, self.n_imu + 1):
                a1, g1 = self._imu(t0 + i * dt)
                pre.push_back(dt, a1, g1)
                buf.append((dt, a1, g1))
            self.t = t0 + self.frame_dt
            Ri, Pi, Vi, T = _R(self.pose[-1, 3:]), self.pose[-1, :3], self.sb[-1, :3], pre.sum_dt
            Pj = Pi + Vi * T - 0.5 * self.g * T * T + Ri @ pre.delta_p
            Vj = Vi - self.g * T + Ri @ pre.delta_v
            qj = S.quat_mul(self.pose[-1, 3:], pre.delta_q)
            qj /= np.linalg.norm(qj)
            if len(self.pose) < K:                     # bootstrap: stand in for the initializer with noisy GT
                p, sbv = self._gt(self.t)
                Pj = p[:3] + self.rng.normal(0, 0.02, 3)
                Vj = sbv[:3] + self.rng.normal(0, 0.05, 3)
                dth = self.rng.normal(0, np.deg2rad(0.3), 3)
                qj = S.quat_mul(p[3:], np.array([dth[0] / 2, dth[1] / 2, dth[2] / 2, 1.0]))
                qj /= np.linalg.norm(qj)
            self.pose = np.vstack([self.pose, np.concatenate([Pj, qj])])
            self.sb = np.vstack([self.sb, np.concatenate([Vj, ba, bg])])
            self.preint.append(S.pack_preint(pre))
            self.pre_obj.append(pre)
            self.imu_buf.append(buf)
        self.frame += 1
        # tracking against the true camera pose of the new frame
        gp, _ = self._gt(self.t)
        Rwc, twc = self._cam_pose(gp)
        ids, xy, _ = self.world.observe(self.cam, Rwc, twc)
        vis = dict(zip(ids.tolist(), range(len(ids))))
        new_idx = len(self.pose) - 1
        n_tracked = 0
        for tr in self.tracks.values():
            if not tr.alive:
                continue
            k = vis.get(tr.lid)
            if k is None or (getattr(self, "track_loss", 0.0) > 0 and self.rng.uniform() < self.track_loss):
                tr.alive = False                                # out of view, or the front end lost the track
            else:
                tr.xy.append(xy[k] + self.rng.normal(0, self.sig, 2))
                n_tracked += 1
        self.margin_flag = self._keyframe_decision(new_idx, n_tracked)
        cand = [i for i in ids.tolist() if i not in self.tracks]
        if len(cand) > self.max_cand:
            cand = sorted(self.rng.choice(cand, self.max_cand, replace=False).tolist())
        cxy = np.array([xy[vis[i]] for i in cand]).reshape(-1, 2)
        return new_idx, n_tracked, np.array(cand, np.int32), cxy

    def _keyframe_decision(self, frame_count, last_track_num):
        """FeatureManager::addFeatureCheckParallax (feature_manager.cpp:46-97): the second-newest frame is a keyframe
        (MARGIN_OLD) when tracking is weak or when the mean parallax between the third-newest and second-newest
        frames (compensatedParallax2, :353-385: no rotation compensation) reaches MIN_PARALLAX."""
        if self.keyframes == "always" or frame_count < 2 or last_track_num < 20:
            return MARGIN_OLD
        par = []
        for tr in self.tracks.values():
            if tr.start <= frame_count - 2 and tr.start + len(tr.xy) - 1 >= frame_count - 1:
                a, b = tr.xy[frame_count - 2 - tr.start], tr.xy[frame_count - 1 - tr.start]
                par.append(np.hypot(a[0] - b[0], a[1] - b[1]))
        if not par:
            return MARGIN_OLD
        return MARGIN_OLD if sum(par) / len(par) >= MIN_PARALLAX else MARGIN_SECOND_NEW

    def _start_tracks(self, new_idx, ids, cand, cxy):
        pos = {int(c): k for k, c in enumerate(cand)}
        for i in ids:
            k = pos[int(i)]
            self.tracks[int(i)] = Track(lid=int(i), start=new_idx, xy=[cxy[k] + self.rng.normal(0, self.sig, 2)])

    def _optimised(self):
        """features that become parameter blocks (estimator.cpp:712-718)"""
        return [tr for tr in self.tracks.values() if len(tr.xy) >= 2 and tr.start < WINDOW_SIZE - 2]

    def build_window(self, backend=None):
        K = self.K
        feats = self._optimised()
        offs, fr, xy = [0], [], []
        for tr in feats:
            n = min(len(tr.xy), K - tr.start)
            fr += list(range(tr.start, tr.start + n))
            xy += tr.xy[:n]
            offs.append(len(fr))
        need = [i for i, tr in enumerate(feats) if tr.depth <= 0]
        if need:                                   # FeatureManager::triangulate for the features without a depth yet
            if backend is not None and hasattr(backend, "triangulate"):
                wt = S.Window(K=K, para_pose=self.pose.copy(), para_speed_bias=self.sb.copy(),
                              para_ex_pose=np.concatenate([self.tic, self.qic]), para_td=np.array([getattr(self, "td_est", 0.0)]),
                              inv_depth=np.ones(len(feats)), lm_obs_offset=np.array(offs, np.int32),
                              obs_frame=np.array(fr, np.int32), obs_xy=np.array(xy, float).reshape(-1, 2),
                              preint=np.array(self.preint), prior=None)
                d = backend.triangulate(wt, INIT_DEPTH)
                for i in need:
                    feats[i].depth = float(d[i])
            else:
                for i in need:
                    feats[i].depth = triangulate(feats[i], self.pose, self.ric, self.tic)
        w = S.Window(K=K, para_pose=self.pose.copy(), para_speed_bias=self.sb.copy(),
                     para_ex_pose=np.concatenate([self.tic, self.qic]), para_td=np.array([getattr(self, "td_est", 0.0)]),
                     inv_depth=np.array([1.0 / tr.depth for tr in feats]), lm_obs_offset=np.array(offs, np.int32),
                     obs_frame=np.array(fr, np.int32), obs_xy=np.array(xy, float).reshape(-1, 2),
                     preint=np.array(self.preint), prior=self.prior)
        return w, feats

    def build_select(self, cand, cxy, kappa):
        """Inputs of FeatureSelector::select for the newest frame k: ground-truth horizon (HorizonGenerator GT mode,
        horizon_generator.cpp:95-123), depth cloud = optimised landmarks seen from frame k (initKDTree,
        feature_selector.cpp:380-432), tracked set = features observed in frame k."""
        pos, quat = self._horizon()
        new_idx = len(self.pose) - 1
        Rc, tc = self._cam_pose(self.pose[new_idx])
        cl_xy, cl_d, used_id, used_xy = [], [], [], []
        for tr in self.tracks.values():
            if tr.alive and tr.start + len(tr.xy) - 1 == new_idx and len(tr.xy) >= 2:
                used_id.append(tr.lid)
                used_xy.append(tr.xy[-1])
            if tr.depth > 0 and len(tr.xy) >= 2 and tr.start < WINDOW_SIZE - 2:
                Ra, ta = self._cam_pose(self.pose[tr.start])
                pw = Ra @ (np.array([tr.xy[0][0], tr.xy[0][1], 1.0]) * tr.depth) + ta
                pc = Rc.T @ (pw - tc)
                if pc[2] > 0.1:
                    cl_xy.append(pc[:2] / pc[2])
                    cl_d.append(tr.depth)
        order = np.argsort(cand)
        omega_prior = None
        if getattr(self, "use_omega_prior", False) and hasattr(self._backend, "omega_prior") and len(self.pose) == self.K:
            # opt-in (not the reference's behaviour): Omega_PRIOR on x_k from the back end's window instead of I9
            omega_prior = self._backend.omega_prior(self.build_window(self._backend)[0], self.opts)
        return S.SelectProblem(H=self.H, horizon_pos=pos, horizon_quat=quat, q_ic=self.qic, t_ic=self.tic, omega_prior=omega_prior,
                               cam=dict(self.cam), nr_imu=self.n_imu, delta_imu=self.frame_dt / self.n_imu,
                               acc_var=S.ACC_N, acc_bias_var=S.ACC_W,
                               cand_id=cand[order], cand_xy=cxy[order], cand_prob=self.world.score[cand[order]],
                               used_id=np.array(used_id, np.int32), used_xy=np.array(used_xy, float).reshape(-1, 2),
                               cloud_xy=np.array(cl_xy, float).reshape(-1, 2), cloud_depth=np.array(cl_d, float),
                               kappa=int(kappa))

    def _horizon(self):
        """Ground-truth horizon of the simulated trajectory (HorizonGenerator GT mode)."""
        # index 0 = x_k, the previous frame; index 1 = x_k+1, the newest frame at self.t, in which the candidates were
        # detected (feature_selector.cpp:247-248) -- the same convention as ReplaySession._horizon
        th = self.t + self.frame_dt * (np.arange(self.H + 1) - 1)
        return np.stack([self.traj.pos(t) for t in th]), np.stack([S.rot_to_quat(self.traj.rot(t)) for t in th])

    def _slide(self):
        """Estimator::slideWindow MARGIN_OLD (estimator.cpp:1000-1040) + removeBackShiftDepth."""
        R0c, P0c = self._cam_pose(self.pose[0])
        R1c, P1c = self._cam_pose(self.pose[1])
        self.pose, self.sb = self.pose[1:].copy(), self.sb[1:].copy()
        self.preint = [np.zeros(S.PREINT_DOUBLES)] + self.preint[2:]
        self.pre_obj = [None] + self.pre_obj[2:]
        self.imu_buf = [[]] + self.imu_buf[2:]
        self.sum_of_back += 1
        dead = []
        for lid, tr in self.tracks.items():
            if tr.start != 0:
                tr.start -= 1
                continue
            uv = tr.xy.pop(0)
            if len(tr.xy) < 2:
                dead.append(lid)
                continue
            if tr.depth > 0:
                pw = R0c @ (np.array([uv[0], uv[1], 1.0]) * tr.depth) + P0c
                dj = (R1c.T @ (pw - P1c))[2]
                tr.depth = dj if dj > 0 else INIT_DEPTH
        for lid in dead:
            del self.tracks[lid]

    def _slide_new(self):
        """Estimator::slideWindow MARGIN_SECOND_NEW (estimator.cpp:1042-1076) + FeatureManager::removeFront."""
        fc = self.K - 1
        pre = self.pre_obj[fc - 1]
        for dt, a, g in self.imu_buf[fc]:
            pre.push_back(dt, a, g)
        self.imu_buf[fc - 1] = self.imu_buf[fc - 1] + self.imu_buf[fc]
        self.preint[fc - 1] = S.pack_preint(pre)
        self.preint, self.pre_obj, self.imu_buf = self.preint[:fc], self.pre_obj[:fc], self.imu_buf[:fc]
        self.pose[fc - 1], self.sb[fc - 1] = self.pose[fc], self.sb[fc]
        self.pose, self.sb = self.pose[:fc].copy(), self.sb[:fc].copy()
        self.sum_of_front += 1
        dead = []
        for lid, tr in self.tracks.items():
            if tr.start == fc:
                tr.start -= 1
                continue
            if tr.start + len(tr.xy) - 1 < fc - 1:
                continue
            del tr.xy[WINDOW_SIZE - 1 - tr.start]
            if not tr.xy:
                dead.append(lid)
        for lid in dead:
            del self.tracks[lid]

    def step(self, backend):
        """Returns None while the window fills, else a dict of per-call wall times (s) and counters."""
        self._backend = backend
        new_idx, n_tracked, cand, cxy = self._ingest()
        lat, job, prob_early = None, None, None
        if len(self.pose) == self.K:
            w, feats = self.build_window(backend)
            pose0 = w.para_pose[0].copy()
            t0 = time.perf_counter()
            wsol, summ = backend.optimize(w, self.opts)
            t1 = time.perf_counter()
            c_opt = getattr(backend, "t_call", 0.0)
            self.pose, self.sb = wsol.para_pose.copy(), wsol.para_speed_bias.copy()
            if self.opts.get("estimate_extrinsic"):             # double2vector: tic / ric follow para_Ex_Pose (estimator.cpp:580-588)
                self.tic, self.qic = np.array(wsol.para_ex_pose[:3], float), np.array(wsol.para_ex_pose[3:], float)
                self.ric = S.quat_to_rot(self.qic)
            if self.opts.get("estimate_td"):                    # td = para_Td[0][0] (estimator.cpp:598-599)
                self.td_est = float(np.atleast_1d(wsol.para_td)[0])
            regauge(pose0, self.pose, self.sb)
            for tr, lam in zip(feats, wsol.inv_depth):          # setDepth + removeFailures
                tr.depth = 1.0 / lam
                if tr.depth < 0:
                    del self.tracks[tr.lid]
            wpost = dataclasses.replace(w, para_pose=self.pose.copy(), para_speed_bias=self.sb.copy(),
                                        para_ex_pose=np.array(wsol.para_ex_pose, float).copy(),
                                        para_td=np.atleast_1d(np.array(wsol.para_td, float)).copy(),
                                        inv_depth=wsol.inv_depth.copy())
            c_marg = 0.0
            overlap = getattr(self, "overlap_marginalize", False) and hasattr(backend, "marginalize_begin")
            prob_early = None
            if overlap and max(0, self.max_feats - n_tracked) > 0 and len(cand) > 0:
                # overlapped mode: the selector's inputs are packed BEFORE the marginalization is enqueued, so that
                # begin -> select -> end run back to back and the timings below hide no host-side work behind the GPU
                prob_early = self.build_select(cand, cxy, max(0, self.max_feats - n_tracked))
            t2 = time.perf_counter()
            if self.margin_flag == MARGIN_OLD:
                if overlap:
                    job = backend.marginalize_begin(wpost, MARGIN_OLD, self.opts)
                else:
                    self.prior = backend.marginalize(wpost, MARGIN_OLD, self.opts)
                c_marg = getattr(backend, "t_call", 0.0)
            elif self.prior is not None:                        # estimator.cpp:926: only if there is a prior; it is
                if overlap:                                     # replaced only if it involves Pose[WINDOW_SIZE-1]
                    job = backend.marginalize_begin(wpost, MARGIN_SECOND_NEW, self.opts)
                else:
                    p = backend.marginalize(wpost, MARGIN_SECOND_NEW, self.opts)
                    if p is not None:
                        self.prior = p
                c_marg = getattr(backend, "t_call", 0.0)
            t3 = time.perf_counter()
            lat = {"optimize": t1 - t0, "marginalize": t3 - t2, "optimize_call": c_opt, "flag": self.margin_flag,
                   "marginalize_call": c_marg, "L": w.L, "n_factors": w.n_factors,
                   "iterations": summ["iterations"], "final_cost": summ["final_cost"]}
            gp, _ = self._gt(self.t)
            self.history.append((self.frame, float(np.linalg.norm(self.pose[-1, :3] - gp[:3])), summ["final_cost"]))
        kappa = max(0, self.max_feats - n_tracked)
        t4 = time.perf_counter()
        if kappa > 0 and len(cand) > 0:
            if len(self.pose) == self.K:
                prob = prob_early if (lat is not None and prob_early is not None) else self.build_select(cand, cxy, kappa)
                t4 = time.perf_counter()
                sel = backend.select(prob)
                t5 = time.perf_counter()
                if lat is not None:
                    lat["select"] = t5 - t4
                    lat["select_call"] = getattr(backend, "t_call", 0.0)
                    lat["N"] = len(cand)
                    lat["kappa"] = kappa
            else:                                               # not initialised: take the strongest corners
                sel = cand[np.argsort(-self.world.score[cand])[:kappa]]
            self.last_selected = np.array(sel, np.int32)
            self._start_tracks(new_idx, self.last_selected, cand, cxy)
        if lat is not None and "select" not in lat:
            lat["select"] = lat["select_call"] = 0.0
        if len(self.pose) == self.K and lat is not None and job is not None:
            # the marginalization that has been running on the second stream while select() worked
            t6 = time.perf_counter()
            p = backend.marginalize_end(job)
            if p is not None or self.margin_flag == MARGIN_OLD:
                self.prior = p if p is not None else self.prior
            lat["marginalize"] += time.perf_counter() - t6
            lat["marginalize_call"] += getattr(backend, "t_call", 0.0)
        if len(self.pose) == self.K:
            self._slide() if self.margin_flag == MARGIN_OLD else self._slide_new()
        return lat


def process_imu(pose, sb, acc_0, gyr_0, samples, g):
    """Estimator::processIMU (estimator.cpp:86-119) for one frame interval: the reference's own prediction of the
    incoming frame -- sample-by-sample midpoint propagation in the world frame, rotation updated by the UN-normalised
    Utility::deltaQ increment through toRotationMatrix() and never re-orthonormalised inside the frame.
    -> (P, R 3x3, V) of the new frame before it is optimised."""
    P, V, R = np.array(pose[:3], float), np.array(sb[:3], float), _R(np.asarray(pose[3:], float))
    ba, bg = np.asarray(sb[3:6], float), np.asarray(sb[6:9], float)
    a0, g0 = np.asarray(acc_0, float), np.asarray(gyr_0, float)
    for dt, a1, g1 in samples:
        un_acc_0 = R @ (a0 - ba) - g
        un_gyr = 0.5 * (g0 + g1) - bg
        q = np.array([un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2, 1.0])      # not normalised
        R = R @ S.quat_to_rot(q)                                    # Eigen's toRotationMatrix formula, no normalisation
        un_acc_1 = R @ (a1 - ba) - g
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        P = P + dt * V + 0.5 * dt * dt * un_acc
        V = V + dt * un_acc
        a0, g0 = np.asarray(a1, float), np.asarray(g1, float)
    return P, R, V


class _Scores:
    """feature id -> detector score of the current image, indexable like World.score"""

    def __init__(self):
        self.d = {}

    def __getitem__(self, idx):
        return np.array([self.d[int(i)] for i in np.atleast_1d(idx)], dtype=float)


class ReplaySession(SlidingWindowSim):
    """The same window bookkeeping driven by recorded front-end traffic instead of the
    simulator: feed it the decoded `sensor_msgs/Imu` and feature `sensor_msgs/PointCloud` messages in arrival order
    (replay.read_dump); every feature message that `replay.get_measurements` can pair with its IMU samples becomes one
    frame.  `init` supplies the states of the first K frames (the stand-in for the reference's visual-inertial
    initializer, which is out of scope).  New features follow FeatureSelector's rule: only ids above the largest id seen
    so far are candidates (feature_selector.cpp:208-219); the horizon is the IMU-mode one (backend.horizon_imu)."""

    def __init__(self, cam, ric, tic, init, max_feats=150, H=10, opts=None, keyframes="parallax", td=0.0, gt=None):
        from . import replay as RP
        self.RP = RP
        self.K = WINDOW_SIZE + 1
        self.cam, self.ric, self.tic = dict(cam), np.array(ric, float), np.array(tic, float)
        self.qic = S.rot_to_quat(self.ric)
        self.max_feats, self.H, self.opts, self.td = max_feats, H, opts or {}, td
        assert keyframes in ("always", "parallax")
        self.keyframes, self.margin_flag = keyframes, MARGIN_OLD
        self.sum_of_back = self.sum_of_front = 0
        self.g = np.array([0, 0, S.G_NORM])
        self.init, self.gt = init, gt
        self.world = type("W", (), {})()
        self.world.score = _Scores()
        self.t, self.frame, self.frame_dt, self.n_imu = 0.0, 0, 0.1, 20
        self.pose, self.sb = np.zeros((0, 7)), np.zeros((0, 9))
        self.preint, self.pre_obj, self.imu_buf = [np.zeros(S.PREINT_DOUBLES)], [None], [[]]
        self.tracks, self.prior = {}, None
        self.last_selected, self.history = np.zeros(0, np.int32), []
        self.last_feature_id = 0
        self.acc_0 = self.gyr_0 = None
        self.imu_q, self.feat_q, self.clock = [], [], RP.ImuClock()
        self._pending = None

    def feed(self, topic, msg, backend):
        """-> list of per-frame latency dicts (see step) for the frames this message completed."""
        (self.imu_q if topic == "imu" else self.feat_q).append(msg)
        out = []
        for ms, img in self.RP.get_measurements(self.imu_q, self.feat_q, self.td):
            self._pending = (ms, img)
            out.append(self.step(backend))
        return out

    def _gt(self, t):
        if self.gt is not None:
            return self.gt(t)
        return self.pose[-1].copy(), self.sb[-1].copy()

    def _ingest(self):
        ms, img = self._pending
        dt, acc, gyr = self.RP.imu_segment(self.clock, ms, img.stamp, self.td)
        self.last_segment = (dt, acc, gyr)                 # what Estimator::processIMU receives for this frame
        image = self.RP.image_from_pointcloud(img)
        K = self.K
        if len(self.pose) == 0:
            p, sbv = self.init[0]
            self.pose, self.sb = np.array(p, float)[None, :].copy(), np.array(sbv, float)[None, :].copy()
        else:
            ba, bg = self.sb[-1, 3:6].copy(), self.sb[-1, 6:9].copy()
            pre = S.Preintegration(self.acc_0, self.gyr_0, ba, bg)
            buf = []
            for k in range(len(dt)):
                pre.push_back(dt[k], acc[k], gyr[k])
                buf.append((dt[k], acc[k].copy(), gyr[k].copy()))
            Pj, Rj, Vj = process_imu(self.pose[-1], self.sb[-1], self.acc_0, self.gyr_0, buf, self.g)
            pj, sj = np.concatenate([Pj, S.rot_to_quat(Rj)]), np.concatenate([Vj, ba, bg])    # vector2double: Quaterniond{Rs}
            if len(self.pose) < K and self.frame in self.init:
                pj, sj = (np.array(x, float) for x in self.init[self.frame])
            self.pose, self.sb = np.vstack([self.pose, pj]), np.vstack([self.sb, sj])
            self.preint.append(S.pack_preint(pre))
            self.pre_obj.append(pre)
            self.imu_buf.append(buf)
            self.frame_dt, self.n_imu = img.stamp - self.t, len(ms)
        self.acc_0, self.gyr_0 = acc[-1].copy(), gyr[-1].copy()
        self.t = img.stamp
        self.frame += 1
        new_idx = len(self.pose) - 1
        n_tracked = 0
        for tr in self.tracks.values():
            if not tr.alive:
                continue
            obs = image.get(tr.lid)
            if obs is None:
                tr.alive = False
            else:
                tr.xy.append(obs[0][1][:2].copy())
                n_tracked += 1
        self.margin_flag = self._keyframe_decision(new_idx, n_tracked)
        cand = np.array(sorted(i for i in image if i > self.last_feature_id), np.int32)
        if len(cand):
            self.last_feature_id = int(cand[-1])
        self.world.score.d = {int(i): float(image[int(i)][0][1][7]) for i in cand}
        cxy = np.array([image[int(i)][0][1][:2] for i in cand]).reshape(-1, 2)
        return new_idx, n_tracked, cand, cxy

    def _start_tracks(self, new_idx, ids, cand, cxy):
        pos = {int(c): k for k, c in enumerate(cand)}
        for i in ids:
            self.tracks[int(i)] = Track(lid=int(i), start=new_idx, xy=[cxy[pos[int(i)]].copy()])

    def _horizon(self):
        """IMU-mode horizon (HorizonGenerator::imu): x_k = the previous frame, x_k+1 = the newest frame, latest IMU
        sample held constant (estimator_node.cpp:326-331)."""
        nr = max(int(self.n_imu), 1)
        return self._backend.horizon_imu(self.H, self.pose[-2, :3], self.pose[-2, 3:], self.sb[-2, 3:6], self.pose[-1, :3],
                                         self.pose[-1, 3:], self.sb[-1, :3], self.acc_0, self.gyr_0 - self.sb[-1, 6:9], nr,
                                         self.frame_dt / nr)


class GpuBackend:
    """optimize / marginalize / select through the C-ABI of libbvio.so (host buffers in, host buffers out).
    `t_call` holds the wall time of the last FFI call alone (without the ctypes packing around it)."""

    def __init__(self, ctx, abi):
        import ctypes
        self.C, self.ctx, self.abi, self.L = ctypes, ctx, abi, ctx.L
        self.t_call = 0.0

    def optimize(self, w, opts):
        C, abi = self.C, self.abi
        h, o, s = abi.WindowHandle(w), abi.default_opts(**opts), abi.Summary()
        t0 = time.perf_counter()
        self.ctx.check(self.L.bvio_optimize(self.ctx.h, C.byref(h.s), C.byref(o), C.byref(s)), "bvio_optimize")
        self.t_call = time.perf_counter() - t0
        return dataclasses.replace(w, para_pose=h.pose, para_speed_bias=h.sb, para_ex_pose=h.ex, para_td=h.td,
                                   inv_depth=h.inv), s.as_dict()

    def triangulate(self, w, init_depth):
        h = self.abi.WindowHandle(w)
        d = np.zeros(w.L)
        self.ctx.check(self.L.bvio_triangulate(self.ctx.h, self.C.byref(h.s), init_depth, self.abi.dptr(d)), "bvio_triangulate")
        return d

    def horizon_imu(self, H, pos0, quat0, ba0, pos1, quat1, vel1, acc, gyr, nr_imu, delta_imu):
        f = lambda a: np.ascontiguousarray(a, np.float64)
        pos, quat = np.zeros((H + 1, 3)), np.zeros((H + 1, 4))
        args = [self.abi.dptr(f(a)) for a in (pos0, quat0, ba0, pos1, quat1, vel1, acc, gyr)]
        self.ctx.check(self.L.bvio_horizon_imu(self.ctx.h, H, *args, nr_imu, delta_imu, self.abi.dptr(pos), self.abi.dptr(quat)),
                       "bvio_horizon_imu")
        return pos, quat

    def marginalize(self, w, flag, opts=None):
        out = self.abi.call_marginalize(self.L.bvio_marginalize, w, flag, ctx=self.ctx.h,
                                        opts=self.abi.default_opts(**(opts or {})))
        self.t_call = self.abi.call_marginalize.t_call
        return out

    def omega_prior(self, w, opts=None):
        h, om = self.abi.WindowHandle(w), np.zeros(81)
        self.ctx.check(self.L.bvio_window_omega_prior(self.ctx.h, self.C.byref(h.s), self.C.byref(self.abi.default_opts(**(opts or {}))),
                                                      self.abi.dptr(om)), "bvio_window_omega_prior")
        return om.reshape(9, 9)

    def marginalize_begin(self, w, flag, opts=None):
        job = self.abi.MarginalizeJob(self.L, self.ctx.h, w, flag, self.abi.default_opts(**(opts or {})))
        self.t_call = job.t_begin
        return job

    def marginalize_end(self, job):
        out = job.end()
        self.t_call = job.t_end
        return out

    def select(self, prob):
        C, abi = self.C, self.abi
        h, ss = abi.SelectHandle(prob), abi.SelectSummary()
        ids = np.zeros(max(prob.kappa, 1), np.int32)
        t0 = time.perf_counter()
        self.ctx.check(self.L.bvio_select(self.ctx.h, C.byref(h.s), abi.iptr(ids), None, C.byref(ss)), "bvio_select")
        self.t_call = time.perf_counter() - t0
        return ids[:ss.n_selected].copy()
