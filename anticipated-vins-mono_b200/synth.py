"""Synthetic sliding windows and selector problems.

Pure numpy input preparation: nothing here is on the product path.  The IMU
preintegration restates IntegrationBase::{push_back,propagate,
midPointIntegration} (vins_estimator/src/factor/integration_base.h:30-158),
which in the reference runs in Estimator::processIMU (not in optimization()),
so it is host-side input prep here as well.

Quaternions are stored x y z w (Eigen coeffs order) like the reference's
para_Pose (vins_estimator/src/estimator.cpp:481-488).
"""
from __future__ import annotations

import dataclasses
import numpy as np

# EuRoC calibration, /root/reference/config/euroc/euroc_config.yaml:11-42
EUROC_CAM = dict(fx=461.6, fy=460.3, cx=363.0, cy=248.1,
                 k1=-2.917e-01, k2=8.228e-02, p1=5.333e-05, p2=-1.578e-04,
                 width=752, height=480)
EUROC_RIC = np.array([[0.0148655429818, -0.999880929698, 0.00414029679422],
                      [0.999557249008, 0.0149672133247, 0.025715529948],
                      [-0.0257744366974, 0.00375618835797, 0.999660727178]])
EUROC_TIC = np.array([-0.0216401454975, -0.064676986768, 0.00981073058949])
ACC_N, GYR_N, ACC_W, GYR_W, G_NORM = 0.08, 0.004, 4.0e-5, 2.0e-6, 9.81007
FOCAL_LENGTH = 460.0


# ----------------------------------------------------------------------------
# small SO(3) helpers
# ----------------------------------------------------------------------------
def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def quat_mul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_inv(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def quat_to_rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rot_to_quat(R):
    """Eigen's Quaterniond(Matrix3d) (trace / largest-diagonal branches)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        return np.array([(R[2, 1] - R[1, 2]) * s, (R[0, 2] - R[2, 0]) * s, (R[1, 0] - R[0, 1]) * s, w])
    i = 0
    if R[1, 1] > R[0, 0]:
        i = 1
    if R[2, 2] > R[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
    q = np.zeros(4)
    q[i] = 0.5 * s
    s = 0.5 / s
    q[3] = (R[k, j] - R[j, k]) * s
    q[j] = (R[j, i] + R[i, j]) * s
    q[k] = (R[k, i] + R[i, k]) * s
    return q


def rot_exp(phi):
    th = np.linalg.norm(phi)
    if th < 1e-12:
        return np.eye(3) + skew(phi)
    a = phi / th
    K = skew(a)
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def euler_zyx(y, p, r):
    cy, sy, cp, sp, cr, sr = np.cos(y), np.sin(y), np.cos(p), np.sin(p), np.cos(r), np.sin(r)
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    Ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    Rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return Rz @ Ry @ Rx


# ----------------------------------------------------------------------------
# camera (PinholeCamera.cc:520-542, 672-688 forward; :491-504 inverse)
# ----------------------------------------------------------------------------
def distortion(cam, mx, my):
    k1, k2, p1, p2 = cam["k1"], cam["k2"], cam["p1"], cam["p2"]
    mx2, my2, mxy = mx * mx, my * my, mx * my
    rho2 = mx2 + my2
    rad = k1 * rho2 + k2 * rho2 * rho2
    return (mx * rad + 2.0 * p1 * mxy + p2 * (rho2 + 2.0 * mx2),
            my * rad + 2.0 * p2 * mxy + p1 * (rho2 + 2.0 * my2))


def space_to_plane(cam, P):
    mx, my = P[0] / P[2], P[1] / P[2]
    dx, dy = distortion(cam, mx, my)
    return np.array([cam["fx"] * (mx + dx) + cam["cx"], cam["fy"] * (my + dy) + cam["cy"]])


def lift_projective(cam, u, v):
    mxd = (u - cam["cx"]) / cam["fx"]
    myd = (v - cam["cy"]) / cam["fy"]
    dx, dy = distortion(cam, mxd, myd)
    mx, my = mxd - dx, myd - dy
    for _ in range(7):
        dx, dy = distortion(cam, mx, my)
        mx, my = mxd - dx, myd - dy
    return np.array([mx, my, 1.0])


def in_fov(cam, px):
    u, v = int(np.round(px[0])), int(np.round(px[1]))
    return 0 <= u < cam["width"] and 0 <= v < cam["height"]


# ----------------------------------------------------------------------------
# trajectory
# ----------------------------------------------------------------------------
class Trajectory:
    """Analytic body trajectory: circle r=2 m at ~1 m/s, yaw rate 0.5 rad/s, with
    vertical motion and small roll/pitch so every IMU axis is excited."""

    def __init__(self, phase=0.0, scale=1.0):
        self.ph = phase
        self.sc = scale

    def pos(self, t):
        t = t + self.ph
        s = self.sc
        return np.array([2.0 * s * np.cos(0.5 * t), 2.0 * s * np.sin(0.5 * t), 0.3 * np.sin(1.1 * t)])

    def vel(self, t):
        t = t + self.ph
        s = self.sc
        return np.array([-1.0 * s * np.sin(0.5 * t), 1.0 * s * np.cos(0.5 * t), 0.33 * np.cos(1.1 * t)])

    def acc(self, t):
        t = t + self.ph
        s = self.sc
        return np.array([-0.5 * s * np.cos(0.5 * t), -0.5 * s * np.sin(0.5 * t), -0.363 * np.sin(1.1 * t)])

    def rot(self, t):
        t = t + self.ph
        # body x forward along the tangent (yaw = heading + pi/2), camera looks along body x
        return euler_zyx(0.5 * t + np.pi / 2, 0.1 * np.cos(0.9 * t), 0.1 * np.sin(0.7 * t))

    def omega_body(self, t, h=1e-6):
        Rm, Rp = self.rot(t - h), self.rot(t + h)
        W = self.rot(t).T @ (Rp - Rm) / (2 * h)
        return np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) * 0.5


# ----------------------------------------------------------------------------
# IMU preintegration (integration_base.h:54-158)
# ----------------------------------------------------------------------------
class Preintegration:
    def __init__(self, acc_0, gyr_0, lin_ba, lin_bg,
                 acc_n=ACC_N, gyr_n=GYR_N, acc_w=ACC_W, gyr_w=GYR_W):
        self.acc_0, self.gyr_0 = np.array(acc_0, float), np.array(gyr_0, float)
        self.linearized_acc, self.linearized_gyr = self.acc_0.copy(), self.gyr_0.copy()   # integration_base.h:16-17
        self.lin_ba, self.lin_bg = np.array(lin_ba, float), np.array(lin_bg, float)
        self.jacobian = np.eye(15)
        self.covariance = np.zeros((15, 15))
        self.sum_dt = 0.0
        self.delta_p = np.zeros(3)
        self.delta_q = np.array([0, 0, 0, 1.0])
        self.delta_v = np.zeros(3)
        n = np.zeros(18)
        n[0:3] = acc_n ** 2
        n[3:6] = gyr_n ** 2
        n[6:9] = acc_n ** 2
        n[9:12] = gyr_n ** 2
        n[12:15] = acc_w ** 2
        n[15:18] = gyr_w ** 2
        self.noise = np.diag(n)

    def push_back(self, dt, acc_1, gyr_1):
        acc_1, gyr_1 = np.array(acc_1, float), np.array(gyr_1, float)
        a0, g0, ba, bg = self.acc_0, self.gyr_0, self.lin_ba, self.lin_bg
        dq, dp, dv = self.delta_q, self.delta_p, self.delta_v
        Rq = quat_to_rot(dq)
        un_acc_0 = Rq @ (a0 - ba)
        un_gyr = 0.5 * (g0 + gyr_1) - bg
        rq = quat_mul(dq, np.array([un_gyr[0] * dt / 2, un_gyr[1] * dt / 2, un_gyr[2] * dt / 2, 1.0]))
        # NOTE: the reference rotates with the un-normalized result_delta_q (Eigen's
        # operator* on a vector uses the formula v + 2w(u x v) + 2u x (u x v), which is
        # only a rotation for unit q) and builds F/V from toRotationMatrix() of the
        # un-normalized quaternion; it normalizes afterwards (integration_base.h:149).
        Rr = quat_to_rot(rq)
        un_acc_1 = _eigen_quat_rotate(rq, acc_1 - ba)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        res_p = dp + dv * dt + 0.5 * un_acc * dt * dt
        res_v = dv + un_acc * dt
        w_x = 0.5 * (g0 + gyr_1) - bg
        a0x, a1x = a0 - ba, acc_1 - ba
        Rw, Ra0, Ra1 = skew(w_x), skew(a0x), skew(a1x)
        I3 = np.eye(3)
        F = np.zeros((15, 15))
        F[0:3, 0:3] = I3
        F[0:3, 3:6] = -0.25 * Rq @ Ra0 * dt * dt + -0.25 * Rr @ Ra1 @ (I3 - Rw * dt) * dt * dt
        F[0:3, 6:9] = I3 * dt
        F[0:3, 9:12] = -0.25 * (Rq + Rr) * dt * dt
        F[0:3, 12:15] = -0.25 * Rr @ Ra1 * dt * dt * -dt
        F[3:6, 3:6] = I3 - Rw * dt
        F[3:6, 12:15] = -1.0 * I3 * dt
        F[6:9, 3:6] = -0.5 * Rq @ Ra0 * dt + -0.5 * Rr @ Ra1 @ (I3 - Rw * dt) * dt
        F[6:9, 6:9] = I3
        F[6:9, 9:12] = -0.5 * (Rq + Rr) * dt
        F[6:9, 12:15] = -0.5 * Rr @ Ra1 * dt * -dt
        F[9:12, 9:12] = I3
        F[12:15, 12:15] = I3
        V = np.zeros((15, 18))
        V[0:3, 0:3] = 0.25 * Rq * dt * dt
        V[0:3, 3:6] = 0.25 * -Rr @ Ra1 * dt * dt * 0.5 * dt
        V[0:3, 6:9] = 0.25 * Rr * dt * dt
        V[0:3, 9:12] = V[0:3, 3:6]
        V[3:6, 3:6] = 0.5 * I3 * dt
        V[3:6, 9:12] = 0.5 * I3 * dt
        V[6:9, 0:3] = 0.5 * Rq * dt
        V[6:9, 3:6] = 0.5 * -Rr @ Ra1 * dt * 0.5 * dt
        V[6:9, 6:9] = 0.5 * Rr * dt
        V[6:9, 9:12] = V[6:9, 3:6]
        V[9:12, 12:15] = I3 * dt
        V[12:15, 15:18] = I3 * dt
        self.jacobian = F @ self.jacobian
        self.covariance = F @ self.covariance @ F.T + V @ self.noise @ V.T
        self.delta_p, self.delta_v = res_p, res_v
        self.delta_q = rq / np.linalg.norm(rq)
        self.sum_dt += dt
        self.acc_0, self.gyr_0 = acc_1, gyr_1


def _eigen_quat_rotate(q, v):
    """Eigen QuaternionBase::_transformVector: v + w*uv + u x uv with uv = 2 u x v."""
    u = q[:3]
    uv = 2.0 * np.cross(u, v)
    return v + q[3] * uv + np.cross(u, uv)


# ----------------------------------------------------------------------------
# window container
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class Window:
    K: int
    para_pose: np.ndarray        # [K,7]
    para_speed_bias: np.ndarray  # [K,9]
    para_ex_pose: np.ndarray     # [7]
    para_td: np.ndarray          # [1]
    inv_depth: np.ndarray        # [L]
    lm_obs_offset: np.ndarray    # [L+1] int32
    obs_frame: np.ndarray        # [n_obs] int32
    obs_xy: np.ndarray           # [n_obs,2]
    preint: np.ndarray           # [K,467] (row 0 unused) -- bvio_preint layout
    prior: dict | None           # n, block_kind, block_frame, block_idx, x0, lin_jac(col-major n*n), lin_res
    gt_pose: np.ndarray | None = None
    gt_speed_bias: np.ndarray | None = None
    gt_inv_depth: np.ndarray | None = None
    # ProjectionTdFactor inputs (estimate_td): per observation feature velocity on the normalized plane, the td in
    # force when the observation was taken (cur_td) and its pixel row (feature_manager.h FeaturePerFrame)
    obs_vel: np.ndarray | None = None    # [n_obs,2]
    obs_td: np.ndarray | None = None     # [n_obs]
    obs_row: np.ndarray | None = None    # [n_obs]
    gt_td: float = 0.0
    # relocalization matches (estimator.cpp:760-792): the loop-closure frame's pose block and, per matched landmark,
    # its observation there
    relo_pose: np.ndarray | None = None  # [7]
    relo_lm: np.ndarray | None = None    # [n_relo] int32, ascending landmark indices
    relo_xy: np.ndarray | None = None    # [n_relo,2]

    @property
    def L(self):
        return len(self.inv_depth)

    @property
    def n_factors(self):
        return len(self.obs_frame) - self.L

    def copy(self):
        return dataclasses.replace(
            self, para_pose=self.para_pose.copy(), para_speed_bias=self.para_speed_bias.copy(),
            para_ex_pose=self.para_ex_pose.copy(), para_td=self.para_td.copy(),
            inv_depth=self.inv_depth.copy())


PREINT_DOUBLES = 3 + 4 + 3 + 3 + 3 + 1 + 225 + 225


def slice_window(w: "Window", f0: int, K: int):
    """The K-frame window that starts at frame f0 of a longer session `w`, as FeatureManager would hold it: observations
    outside [f0, f0 + K) are gone, landmarks with fewer than two left (or anchored too late to be optimised,
    estimator.cpp:712-718) are dropped, and a landmark whose anchor observation fell out is re-anchored at its first
    remaining observation with its depth carried over through the current pose estimates (removeBackShiftDepth,
    feature_manager.cpp:275-311).  Returns (window, idx) with idx[l] = index of landmark l in `w`."""
    U_, _, Vt_ = np.linalg.svd(EUROC_RIC)
    ric, tic = U_ @ Vt_, EUROC_TIC
    Rs = [quat_to_rot(q / np.linalg.norm(q)) for q in w.para_pose[:, 3:]]
    Rg = [quat_to_rot(q) for q in w.gt_pose[:, 3:]] if w.gt_pose is not None else None
    offs, fr, xy, inv, ginv, idx = [0], [], [], [], [], []
    for l in range(w.L):
        o0, o1 = int(w.lm_obs_offset[l]), int(w.lm_obs_offset[l + 1])
        sel = [k for k in range(o0, o1) if f0 <= w.obs_frame[k] < f0 + K]
        if len(sel) < 2 or w.obs_frame[sel[0]] - f0 >= K - 3:
            continue
        lam, glam = float(w.inv_depth[l]), float(w.gt_inv_depth[l]) if w.gt_inv_depth is not None else 0.0
        if sel[0] != o0:                      # re-anchor: depth in the old anchor camera -> world -> new anchor camera
            fa, fn = int(w.obs_frame[o0]), int(w.obs_frame[sel[0]])
            pa = np.array([w.obs_xy[o0, 0], w.obs_xy[o0, 1], 1.0])
            pw = Rs[fa] @ (ric @ (pa / lam) + tic) + w.para_pose[fa, :3]
            lam = 1.0 / (ric.T @ (Rs[fn].T @ (pw - w.para_pose[fn, :3]) - tic))[2]
            if Rg is not None:
                pw = Rg[fa] @ (ric @ (pa / glam) + tic) + w.gt_pose[fa, :3]
                glam = 1.0 / (ric.T @ (Rg[fn].T @ (pw - w.gt_pose[fn, :3]) - tic))[2]
        fr += [int(w.obs_frame[k]) - f0 for k in sel]
        xy += [w.obs_xy[k] for k in sel]
        offs.append(len(fr))
        inv.append(lam)
        ginv.append(glam)
        idx.append(l)
    pre = np.zeros((K, PREINT_DOUBLES))
    pre[1:] = w.preint[f0 + 1:f0 + K]
    out = Window(K=K, para_pose=w.para_pose[f0:f0 + K].copy(), para_speed_bias=w.para_speed_bias[f0:f0 + K].copy(),
                 para_ex_pose=w.para_ex_pose.copy(), para_td=w.para_td.copy(), inv_depth=np.array(inv),
                 lm_obs_offset=np.array(offs, np.int32), obs_frame=np.array(fr, np.int32),
                 obs_xy=np.array(xy, float).reshape(-1, 2), preint=pre, prior=w.prior if f0 == 0 else None,
                 gt_pose=w.gt_pose[f0:f0 + K].copy() if w.gt_pose is not None else None,
                 gt_speed_bias=w.gt_speed_bias[f0:f0 + K].copy() if w.gt_speed_bias is not None else None,
                 gt_inv_depth=np.array(ginv) if w.gt_inv_depth is not None else None)
    return out, np.array(idx, np.int64)


def make_session(seed, K=11, L=1500, track_min=6):
    """A (K+1)-frame session from which two CONSECUTIVE K-frame windows are cut (bench.py builds its pool of windows with
    real marginalized priors from these): ~L landmarks in either window."""
    return feature_manager_order(make_window(seed=seed, K=K + 1, L=int(L * 1.04), track_min=track_min, track_max=K + 1))


def feature_manager_order(w: "Window") -> "Window":
    """The same window with its landmarks in the order FeatureManager's list holds them: features are appended when
    they are first seen (feature_manager.cpp:46-97), so start frames never decrease along the list (stable: ties keep
    their order).  make_window() draws start frames at random; a session that stands for the reference's traffic is
    sorted once here and every window sliced from it inherits the order."""
    start = w.obs_frame[w.lm_obs_offset[:-1]]
    order = np.argsort(start, kind="stable")
    n = np.diff(w.lm_obs_offset)[order]
    offs = np.concatenate([[0], np.cumsum(n)]).astype(np.int32)
    obs = np.concatenate([np.arange(w.lm_obs_offset[l], w.lm_obs_offset[l + 1]) for l in order]) if len(order) else np.zeros(0, int)
    out = w.copy()
    out.inv_depth = w.inv_depth[order].copy()
    out.lm_obs_offset = offs
    out.obs_frame = w.obs_frame[obs].copy()
    out.obs_xy = w.obs_xy[obs].copy()
    for name in ("obs_vel", "obs_td", "obs_row"):
        if getattr(w, name, None) is not None:
            setattr(out, name, getattr(w, name)[obs].copy())
    if w.gt_inv_depth is not None:
        out.gt_inv_depth = w.gt_inv_depth[order].copy()
    assert w.relo_lm is None or len(w.relo_lm) == 0, "sort before adding relocalization matches"
    return out


def consecutive_window(session: "Window", first_solved: "Window", first_idx, prior: dict, K=11):
    """The second window of make_session(): frames 1 .. K, starting from the first window's SOLVED state (frames and
    landmarks they share) and carrying `prior`, the marginalization of the first window's frame 0 -- the situation of
    every Estimator::optimization() call after the first (estimator.cpp:694-700)."""
    ses = session.copy()
    ses.para_pose[:K] = first_solved.para_pose
    ses.para_speed_bias[:K] = first_solved.para_speed_bias
    ses.inv_depth[first_idx] = first_solved.inv_depth
    w, _ = slice_window(ses, 1, K)
    w.prior = prior
    return w


def add_relocalization(w: "Window", seed=0, local_index=4, max_matches=40, pose_sigma=(0.05, 0.01)) -> "Window":
    """Relocalization inputs as Estimator::setReloFrame leaves them (estimator.cpp:1109-1127) for a window made by
    make_window: the loop-closure frame is an earlier visit near window frame `local_index`; every landmark anchored at
    a frame <= local_index may have a match there (estimator.cpp:774).  relo_pose starts as the window's own pose of
    that frame (:1124); the matches are the landmarks' ground-truth positions seen from a nearby true pose, plus pixel
    noise.  Returns a copy of `w` with relo_pose / relo_lm / relo_xy set."""
    rng = np.random.default_rng(7000 + seed)
    U_, _, Vt_ = np.linalg.svd(EUROC_RIC)
    ric, tic = U_ @ Vt_, EUROC_TIC
    gp = w.gt_pose if w.gt_pose is not None else w.para_pose
    ginv = w.gt_inv_depth if w.gt_inv_depth is not None else w.inv_depth
    # true pose of the old visit: the ground-truth pose of frame local_index, moved a little
    Pt = gp[local_index, :3] + rng.normal(0, pose_sigma[0], 3)
    qt = quat_mul(gp[local_index, 3:], np.concatenate([0.5 * rng.normal(0, pose_sigma[1], 3), [1.0]]))
    Rt = quat_to_rot(qt / np.linalg.norm(qt))
    lm, xy = [], []
    eligible = [l for l in range(w.L) if int(w.obs_frame[w.lm_obs_offset[l]]) <= local_index]
    for l in range(w.L):
        o0 = w.lm_obs_offset[l]
        fi = int(w.obs_frame[o0])
        # the last eligible landmark is always matched: the reference's walk over match_points (estimator.cpp:776-779)
        # has no end check and would read past the vector for an eligible feature id above every matched id
        last = bool(eligible) and l == eligible[-1]
        if fi > local_index or (not last and (len(lm) >= max_matches or rng.uniform() < 0.3)):
            continue
        Ri, Pi = quat_to_rot(gp[fi, 3:]), gp[fi, :3]
        pw = Ri @ (ric @ (np.array([w.obs_xy[o0, 0], w.obs_xy[o0, 1], 1.0]) / ginv[l]) + tic) + Pi
        pc = ric.T @ (Rt.T @ (pw - Pt) - tic)
        if pc[2] < 0.2 and not last:
            continue
        lm.append(l)
        xy.append(pc[:2] / pc[2] + rng.normal(0, 1.5 / FOCAL_LENGTH, 2))
    return dataclasses.replace(w.copy(), relo_pose=w.para_pose[local_index].copy(), relo_lm=np.array(lm, np.int32),
                               relo_xy=np.array(xy, float).reshape(-1, 2))


def pack_preint(p: Preintegration) -> np.ndarray:
    return np.concatenate([p.delta_p, p.delta_q, p.delta_v, p.lin_ba, p.lin_bg, [p.sum_dt],
                           p.jacobian.reshape(-1), p.covariance.reshape(-1)])


def make_window(seed=0, K=11, L=150, track_min=3, track_max=11, frame_dt=0.1, imu_rate=200,
                prior="frame0", noise=True, perturb=True, depth_range=(2.0, 10.0), td_true=None) -> Window:
    """Configs 1-3 of SURVEY.md section 8d.  prior: 'frame0' (15-dim full-rank prior on
    frame 0, Lambda = 1e4 I), 'none'."""
    rng = np.random.default_rng(seed)
    traj = Trajectory(phase=rng.uniform(0, 10.0), scale=rng.uniform(0.8, 1.2))
    cam = EUROC_CAM
    ric, tic = EUROC_RIC, EUROC_TIC
    # re-orthonormalise the 13-digit calibration matrix so that qic is a unit quaternion
    U_, _, Vt_ = np.linalg.svd(ric)
    ric = U_ @ Vt_
    qic = rot_to_quat(ric)
    t0 = 1.0
    tk = t0 + frame_dt * np.arange(K)
    gvec = np.array([0, 0, G_NORM])
    ba_true = rng.normal(0, 0.02, 3)
    bg_true = rng.normal(0, 0.002, 3)
    n_imu = int(round(frame_dt * imu_rate))
    dt = frame_dt / n_imu

    def imu_at(t):
        R = traj.rot(t)
        acc = R.T @ (traj.acc(t) + gvec) + ba_true
        gyr = traj.omega_body(t) + bg_true
        if noise:
            acc = acc + rng.normal(0, ACC_N, 3)
            gyr = gyr + rng.normal(0, GYR_N, 3)
        return acc, gyr

    gt_pose = np.zeros((K, 7))
    gt_sb = np.zeros((K, 9))
    for k in range(K):
        gt_pose[k, :3] = traj.pos(tk[k])
        gt_pose[k, 3:] = rot_to_quat(traj.rot(tk[k]))
        gt_sb[k, :3] = traj.vel(tk[k])
        gt_sb[k, 3:6] = ba_true
        gt_sb[k, 6:9] = bg_true

    # initial bias estimates (= linearization point of the preintegrations)
    ba_est = ba_true + (rng.normal(0, 0.01, 3) if perturb else 0)
    bg_est = bg_true + (rng.normal(0, 0.001, 3) if perturb else 0)

    preint = np.zeros((K, PREINT_DOUBLES))
    for k in range(1, K):
        a0, g0 = imu_at(tk[k - 1])
        pre = Preintegration(a0, g0, ba_est, bg_est)
        for i in range(1, n_imu + 1):
            a1, g1 = imu_at(tk[k - 1] + i * dt)
            pre.push_back(dt, a1, g1)
        preint[k] = pack_preint(pre)

    # landmarks
    offs = [0]
    obs_frame, obs_xy, inv_depth_gt = [], [], []
    obs_vel, obs_row = [], []
    sig_px = 1.5 / FOCAL_LENGTH
    tries = 0
    while len(inv_depth_gt) < L and tries < 100 * L + 1000:
        tries += 1
        nl = int(rng.integers(track_min, track_max + 1))
        nl = min(nl, K)
        max_start = min(K - nl, K - 4 if K >= 4 else 0)   # start_frame < WINDOW_SIZE-2 = K-3
        if K < 4:
            max_start = 0
        start = int(rng.integers(0, max_start + 1))
        u, v = rng.uniform(20, cam["width"] - 20), rng.uniform(20, cam["height"] - 20)
        ray = lift_projective(cam, u, v)
        depth = rng.uniform(*depth_range)
        pc = ray * depth
        Ri, Pi = traj.rot(tk[start]), traj.pos(tk[start])
        pw = Ri @ (ric @ pc + tic) + Pi
        frames, pts, vels, rows = [], [], [], []
        ok = True
        for j in range(start, start + nl):
            Rj, Pj = traj.rot(tk[j]), traj.pos(tk[j])
            pcj = ric.T @ (Rj.T @ (pw - Pj) - tic)
            if pcj[2] < 0.2 or not in_fov(cam, space_to_plane(cam, pcj)):
                ok = False
                break
            xy = pcj[:2] / pcj[2]
            if td_true is not None:
                # feature velocity on the normalized plane (feature_tracker publishes it, estimator_node.cpp:318-320);
                # the image is taken td_true late: observed = xy(t_k) + td_true * velocity
                hh = 1e-4
                pa = ric.T @ (traj.rot(tk[j] + hh).T @ (pw - traj.pos(tk[j] + hh)) - tic)
                pb = ric.T @ (traj.rot(tk[j] - hh).T @ (pw - traj.pos(tk[j] - hh)) - tic)
                vel = (pa[:2] / pa[2] - pb[:2] / pb[2]) / (2 * hh)
                vels.append(vel)
                rows.append(space_to_plane(cam, pcj)[1])
                xy = xy + td_true * vel
            if noise:
                xy = xy + rng.normal(0, sig_px, 2)
            frames.append(j)
            pts.append(xy)
        if not ok and len(frames) < 2:
            continue
        # keep the visible prefix (contiguous track, like FeatureManager)
        obs_frame += frames
        obs_xy += pts
        obs_vel += vels[:len(frames)]
        obs_row += rows[:len(frames)]
        offs.append(len(obs_frame))
        inv_depth_gt.append(1.0 / depth)
    assert len(inv_depth_gt) == L, "could not place enough landmarks"
    inv_depth_gt = np.array(inv_depth_gt)

    pose0 = gt_pose.copy()
    sb0 = gt_sb.copy()
    sb0[:, 3:6] = ba_est
    sb0[:, 6:9] = bg_est
    inv0 = inv_depth_gt.copy()
    if perturb:
        for k in range(K):
            pose0[k, :3] += rng.normal(0, 0.05, 3)
            dth = rng.normal(0, np.deg2rad(1.0), 3)
            q = quat_mul(pose0[k, 3:], np.array([dth[0] / 2, dth[1] / 2, dth[2] / 2, 1.0]))
            pose0[k, 3:] = q / np.linalg.norm(q)
            sb0[k, :3] += rng.normal(0, 0.1, 3)
        inv0 = inv_depth_gt * rng.uniform(0.8, 1.25, L)

    pr = None
    if prior == "frame0":
        n = 15
        pr = dict(n=n, block_kind=np.array([0, 1], np.int32), block_frame=np.array([0, 0], np.int32),
                  block_idx=np.array([0, 6], np.int32),
                  x0=np.concatenate([gt_pose[0], gt_sb[0, :3], ba_est, bg_est]),
                  # sigma: p,theta,v 1e-2; ba 1e-3; bg 1e-4 (a marginalization prior carries
                  # accumulated bias information; a flat 1e4 leaves the bias common mode at
                  # a relative curvature of 1e-12 against the reference's bias random walk)
                  lin_jac=np.diag([100.0] * 9 + [1e3] * 3 + [1e4] * 3).reshape(-1, order="F").copy(),
                  lin_res=(rng.normal(0, 0.1, n) if noise else np.zeros(n)))
    return Window(K=K, para_pose=pose0, para_speed_bias=sb0,
                  para_ex_pose=np.concatenate([tic, qic]), para_td=np.zeros(1),
                  inv_depth=inv0, lm_obs_offset=np.array(offs, np.int32),
                  obs_frame=np.array(obs_frame, np.int32), obs_xy=np.array(obs_xy, float).reshape(-1, 2),
                  preint=preint, prior=pr, gt_pose=gt_pose, gt_speed_bias=gt_sb, gt_inv_depth=inv_depth_gt,
                  obs_vel=np.array(obs_vel, float).reshape(-1, 2) if td_true is not None else None,
                  obs_td=np.zeros(len(obs_frame)) if td_true is not None else None,
                  obs_row=np.array(obs_row, float) if td_true is not None else None,
                  gt_td=td_true or 0.0)


# ----------------------------------------------------------------------------
# selector problem (config 4)
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class SelectProblem:
    H: int
    horizon_pos: np.ndarray   # [H+1,3]
    horizon_quat: np.ndarray  # [H+1,4]
    q_ic: np.ndarray
    t_ic: np.ndarray
    cam: dict
    nr_imu: int
    delta_imu: float
    acc_var: float
    acc_bias_var: float
    cand_id: np.ndarray
    cand_xy: np.ndarray
    cand_prob: np.ndarray
    used_id: np.ndarray
    used_xy: np.ndarray
    cloud_xy: np.ndarray
    cloud_depth: np.ndarray
    kappa: int
    # ABI v2, optional: the reference's state_k1_ when it differs from horizon[1] (ground-truth horizon mode), and an
    # Omega_PRIOR on x_k replacing I9
    state_k1_pos: np.ndarray = None
    state_k1_quat: np.ndarray = None
    omega_prior: np.ndarray = None

    @property
    def N(self):
        return len(self.cand_id)


def make_select_problem(seed=0, N=2000, H=10, U=0, C=150, kappa=150, frame_dt=0.1, nr_imu=20,
                        first_id=1000) -> SelectProblem:
    rng = np.random.default_rng(1000 + seed)
    traj = Trajectory(phase=rng.uniform(0, 10.0), scale=rng.uniform(0.8, 1.2))
    cam = EUROC_CAM
    U_, _, Vt_ = np.linalg.svd(EUROC_RIC)
    ric = U_ @ Vt_
    t0 = 2.0
    th = t0 + frame_dt * np.arange(H + 1)      # x_k, x_k+1, ... x_k+H
    pos = np.stack([traj.pos(t) for t in th])
    quat = np.stack([rot_to_quat(traj.rot(t)) for t in th])

    def sample_xy(n):
        out = np.zeros((n, 2))
        for i in range(n):
            u, v = rng.uniform(0, cam["width"] - 1), rng.uniform(0, cam["height"] - 1)
            out[i] = lift_projective(cam, u, v)[:2]
        return out

    cand_xy = sample_xy(N)
    cand_prob = rng.uniform(0.05, 1.0, N)
    cand_id = first_id + np.sort(rng.choice(4 * N, size=N, replace=False)).astype(np.int32)
    used_xy = sample_xy(U)
    used_id = np.arange(U, dtype=np.int32)
    cloud_xy = sample_xy(C)
    cloud_depth = rng.uniform(2.0, 10.0, C)
    return SelectProblem(H=H, horizon_pos=pos, horizon_quat=quat, q_ic=rot_to_quat(ric), t_ic=EUROC_TIC.copy(),
                         cam=dict(cam), nr_imu=nr_imu, delta_imu=frame_dt / nr_imu,
                         acc_var=ACC_N, acc_bias_var=ACC_W,
                         cand_id=cand_id, cand_xy=cand_xy, cand_prob=cand_prob,
                         used_id=used_id, used_xy=used_xy, cloud_xy=cloud_xy, cloud_depth=cloud_depth,
                         kappa=kappa)
