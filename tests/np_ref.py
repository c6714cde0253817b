"""Independent numpy/scipy restatement of the reference hot path (second opinion for
the C++ oracle; SURVEY.md section 8c "Oracle plan").  Written with rotation
matrices and dense linear algebra on purpose, so that it shares no code structure
with oracle/*.cpp.  Reference citations as in oracle/ba_oracle.cpp / sel_oracle.cpp.
"""
import numpy as np


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def R_of(q):  # q = x y z w, unit
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def qinv(q):
    return np.array([-q[0], -q[1], -q[2], q[3]]) / np.dot(q, q)


def Qleft(q):  # w-first 4x4, utility.h:49-57
    x, y, z, w = q
    m = np.zeros((4, 4))
    m[0, 0] = w
    m[0, 1:] = -np.array([x, y, z])
    m[1:, 0] = [x, y, z]
    m[1:, 1:] = w * np.eye(3) + skew([x, y, z])
    return m


def Qright(q):
    x, y, z, w = q
    m = np.zeros((4, 4))
    m[0, 0] = w
    m[0, 1:] = -np.array([x, y, z])
    m[1:, 0] = [x, y, z]
    m[1:, 1:] = w * np.eye(3) - skew([x, y, z])
    return m


def pose_plus(x, d):
    out = np.array(x, float).copy()
    out[:3] += d[:3]
    q = qmul(x[3:], np.array([d[3] / 2, d[4] / 2, d[5] / 2, 1.0]))
    out[3:] = q / np.linalg.norm(q)
    return out


# ---------------------------------------------------------------------------
def projection(pts_i, pts_j, pose_i, pose_j, ex, inv_dep, sqrt_info=460.0 / 1.5, jac=True):
    Pi, Ri = pose_i[:3], R_of(pose_i[3:])
    Pj, Rj = pose_j[:3], R_of(pose_j[3:])
    tic, ric = ex[:3], R_of(ex[3:])
    pci = np.array([pts_i[0], pts_i[1], 1.0]) / inv_dep
    pii = ric @ pci + tic
    pw = Ri @ pii + Pi
    pij = Rj.T @ (pw - Pj)
    pcj = ric.T @ (pij - tic)
    dep = pcj[2]
    r = sqrt_info * (pcj[:2] / dep - np.asarray(pts_j[:2]))
    if not jac:
        return r
    red = sqrt_info * np.array([[1 / dep, 0, -pcj[0] / dep ** 2], [0, 1 / dep, -pcj[1] / dep ** 2]])
    Ji = red @ np.hstack([ric.T @ Rj.T, ric.T @ Rj.T @ Ri @ -skew(pii)])
    Jj = red @ np.hstack([ric.T @ -Rj.T, ric.T @ skew(pij)])
    tmp_r = ric.T @ Rj.T @ Ri @ ric
    Jex = red @ np.hstack([ric.T @ (Rj.T @ Ri - np.eye(3)),
                           -tmp_r @ skew(pci) + skew(tmp_r @ pci) + skew(ric.T @ (Rj.T @ (Ri @ tic + Pi - Pj) - tic))])
    Jf = red @ (tmp_r @ np.array([pts_i[0], pts_i[1], 1.0])) * -1.0 / inv_dep ** 2
    return r, Ji, Jj, Jex, Jf


def unpack_preint(row):
    return dict(delta_p=row[0:3], delta_q=row[3:7], delta_v=row[7:10], lin_ba=row[10:13], lin_bg=row[13:16],
                sum_dt=row[16], jacobian=row[17:242].reshape(15, 15), covariance=row[242:467].reshape(15, 15))


def imu_sqrt_info(cov):
    return np.linalg.cholesky(np.linalg.inv(cov)).T


def imu(pre, G, pose_i, sb_i, pose_j, sb_j, jac=True):
    Pi, Qi, Pj, Qj = pose_i[:3], pose_i[3:], pose_j[:3], pose_j[3:]
    Vi, Bai, Bgi = sb_i[:3], sb_i[3:6], sb_i[6:9]
    Vj, Baj, Bgj = sb_j[:3], sb_j[3:6], sb_j[6:9]
    J = pre["jacobian"]
    dp_dba, dp_dbg, dq_dbg, dv_dba, dv_dbg = J[0:3, 9:12], J[0:3, 12:15], J[3:6, 12:15], J[6:9, 9:12], J[6:9, 12:15]
    dba, dbg = Bai - pre["lin_ba"], Bgi - pre["lin_bg"]
    th = dq_dbg @ dbg
    cq = qmul(pre["delta_q"], np.array([th[0] / 2, th[1] / 2, th[2] / 2, 1.0]))
    cv = pre["delta_v"] + dv_dba @ dba + dv_dbg @ dbg
    cp = pre["delta_p"] + dp_dba @ dba + dp_dbg @ dbg
    dt = pre["sum_dt"]
    Ri = R_of(Qi)
    G = np.asarray(G)
    r = np.zeros(15)
    r[0:3] = Ri.T @ (0.5 * G * dt * dt + Pj - Pi - Vi * dt) - cp
    r[3:6] = 2 * qmul(qinv(cq), qmul(qinv(Qi), Qj))[:3]
    r[6:9] = Ri.T @ (G * dt + Vj - Vi) - cv
    r[9:12] = Baj - Bai
    r[12:15] = Bgj - Bgi
    si = imu_sqrt_info(pre["covariance"])
    if not jac:
        return si @ r
    Jpi = np.zeros((15, 6))
    Jpi[0:3, 0:3] = -Ri.T
    Jpi[0:3, 3:6] = skew(Ri.T @ (0.5 * G * dt * dt + Pj - Pi - Vi * dt))
    Jpi[3:6, 3:6] = -(Qleft(qmul(qinv(Qj), Qi)) @ Qright(cq))[1:, 1:]
    Jpi[6:9, 3:6] = skew(Ri.T @ (G * dt + Vj - Vi))
    Jsi = np.zeros((15, 9))
    Jsi[0:3, 0:3] = -Ri.T * dt
    Jsi[0:3, 3:6] = -dp_dba
    Jsi[0:3, 6:9] = -dp_dbg
    Jsi[3:6, 6:9] = -Qleft(qmul(qmul(qinv(Qj), Qi), pre["delta_q"]))[1:, 1:] @ dq_dbg
    Jsi[6:9, 0:3] = -Ri.T
    Jsi[6:9, 3:6] = -dv_dba
    Jsi[6:9, 6:9] = -dv_dbg
    Jsi[9:12, 3:6] = -np.eye(3)
    Jsi[12:15, 6:9] = -np.eye(3)
    Jpj = np.zeros((15, 6))
    Jpj[0:3, 0:3] = Ri.T
    Jpj[3:6, 3:6] = Qleft(qmul(qmul(qinv(cq), qinv(Qi)), Qj))[1:, 1:]
    Jsj = np.zeros((15, 9))
    Jsj[6:9, 0:3] = Ri.T
    Jsj[9:12, 3:6] = np.eye(3)
    Jsj[12:15, 6:9] = np.eye(3)
    return si @ r, si @ Jpi, si @ Jsi, si @ Jpj, si @ Jsj


def prior_residual(prior, w):
    n = prior["n"]
    dx = np.zeros(n)
    off = 0
    for kind, frame, idx in zip(prior["block_kind"], prior["block_frame"], prior["block_idx"]):
        if kind == 0:
            x, x0 = w.para_pose[frame], prior["x0"][off:off + 7]
            dx[idx:idx + 3] = x[:3] - x0[:3]
            dq = qmul(qinv(x0[3:]), x[3:])
            v = 2 * dq[:3]
            dx[idx + 3:idx + 6] = v if dq[3] >= 0 else -v
            off += 7
        elif kind == 1:
            dx[idx:idx + 9] = w.para_speed_bias[frame] - prior["x0"][off:off + 9]
            off += 9
        elif kind == 2:
            x, x0 = w.para_ex_pose, prior["x0"][off:off + 7]
            dx[idx:idx + 3] = x[:3] - x0[:3]
            dq = qmul(qinv(x0[3:]), x[3:])
            v = 2 * dq[:3]
            dx[idx + 3:idx + 6] = v if dq[3] >= 0 else -v
            off += 7
        else:
            dx[idx] = w.para_td[0] - prior["x0"][off]
            off += 1
    Jm = np.asarray(prior["lin_jac"]).reshape(n, n, order="F")
    return prior["lin_res"] + Jm @ dx, dx, Jm


def projection_td(pts_i, pts_j, vel_i, vel_j, td_i, td_j, row_i, row_j, TR, ROW, pose_i, pose_j, ex, inv_dep, td,
                  sqrt_info=460.0 / 1.5):
    """ProjectionTdFactor::Evaluate (projection_td_factor.cpp:34-141): the projection chain on the time-shifted
    points, plus the 2x1 Jacobian w.r.t. td."""
    si = td - td_i + TR / ROW * (row_i - ROW / 2)
    sj = td - td_j + TR / ROW * (row_j - ROW / 2)
    pi_td = np.asarray(pts_i[:2]) - si * np.asarray(vel_i)
    pj_td = np.asarray(pts_j[:2]) - sj * np.asarray(vel_j)
    r, Ji, Jj, Jex, Jf = projection(pi_td, pj_td, pose_i, pose_j, ex, inv_dep, sqrt_info)
    Ri, Rj, ric = R_of(pose_i[3:]), R_of(pose_j[3:]), R_of(ex[3:])
    pcj = ric.T @ (Rj.T @ (Ri @ (ric @ (np.array([pi_td[0], pi_td[1], 1.0]) / inv_dep) + ex[:3]) + pose_i[:3] - pose_j[:3]) - ex[:3])
    dep = pcj[2]
    red = sqrt_info * np.array([[1 / dep, 0, -pcj[0] / dep ** 2], [0, 1 / dep, -pcj[1] / dep ** 2]])
    Jtd = red @ (ric.T @ Rj.T @ Ri @ ric @ np.array([vel_i[0], vel_i[1], 0.0])) / inv_dep * -1.0 + sqrt_info * np.asarray(vel_j)
    return r, Ji, Jj, Jex, Jf, Jtd


def full_system_ext(w, est_ex=False, est_td=False, TR=0.0, ROW=480.0, G=(0, 0, 9.81007), sqrt_info=460.0 / 1.5, cauchy_a=1.0):
    """full_system with the extrinsic pose and / or td as free parameters: columns [15K | 6 ex | 1 td | L]."""
    K, L = w.K, len(w.inv_depth)
    n_ex, n_td = (6 if est_ex else 0), (1 if est_td else 0)
    o_ex, o_td = 15 * K, 15 * K + n_ex
    npar = o_td + n_td
    J0, r0, cost = full_system(w, G=G, sqrt_info=sqrt_info, cauchy_a=cauchy_a) if not est_td else (None, None, None)
    rows, res = [], []
    cost_v = 0.0
    for l in range(L):
        o0, o1 = w.lm_obs_offset[l], w.lm_obs_offset[l + 1]
        fi = w.obs_frame[o0]
        for k in range(o0 + 1, o1):
            fj = w.obs_frame[k]
            if est_td:
                r, Ji, Jj, Jex, Jf, Jtd = projection_td(w.obs_xy[o0], w.obs_xy[k], w.obs_vel[o0], w.obs_vel[k], w.obs_td[o0],
                                                        w.obs_td[k], w.obs_row[o0], w.obs_row[k], TR, ROW, w.para_pose[fi],
                                                        w.para_pose[fj], w.para_ex_pose, w.inv_depth[l], w.para_td[0], sqrt_info)
            else:
                r, Ji, Jj, Jex, Jf = projection(w.obs_xy[o0], w.obs_xy[k], w.para_pose[fi], w.para_pose[fj], w.para_ex_pose,
                                                w.inv_depth[l], sqrt_info)
            sq = r @ r
            cost_v += 0.5 * cauchy_a ** 2 * np.log1p(sq / cauchy_a ** 2)
            sr = np.sqrt(1.0 / (1.0 + sq / cauchy_a ** 2))
            blk = np.zeros((2, npar + L))
            blk[:, 15 * fi:15 * fi + 6] += sr * Ji
            blk[:, 15 * fj:15 * fj + 6] += sr * Jj
            if est_ex:
                blk[:, o_ex:o_ex + 6] = sr * Jex
            if est_td:
                blk[:, o_td] = sr * Jtd
            blk[:, npar + l] = sr * Jf
            rows.append(blk)
            res.append(sr * r)
    # IMU and prior rows: reuse the fixed-extrinsic builder on a window without landmarks, widened to the new layout
    import dataclasses
    bare = dataclasses.replace(w, inv_depth=np.zeros(0), lm_obs_offset=np.zeros(1, np.int32), obs_frame=np.zeros(0, np.int32),
                               obs_xy=np.zeros((0, 2)))
    Jb, rb, cb = full_system(bare, G=G, sqrt_info=sqrt_info, cauchy_a=cauchy_a)
    wide = np.zeros((Jb.shape[0], npar + L))
    wide[:, :15 * K] = Jb
    if w.prior is not None:      # prior columns on the extrinsic / td blocks
        rp, dx, Jm = prior_residual(w.prior, w)
        n = w.prior["n"]
        for kind, frame, idx in zip(w.prior["block_kind"], w.prior["block_frame"], w.prior["block_idx"]):
            if kind == 2 and est_ex:
                wide[-n:, o_ex:o_ex + 6] = Jm[:, idx:idx + 6]
            if kind == 3 and est_td:
                wide[-n:, o_td] = Jm[:, idx]
    rows.append(wide)
    res.append(rb)
    return np.vstack(rows), np.concatenate(res), cost_v + cb


def reduced_system_ext(w, **kw):
    J, r, cost = full_system_ext(w, **kw)
    npar = J.shape[1] - len(w.inv_depth)
    Hf, gf = J.T @ J, J.T @ r
    Hpp, Hpl, hl = Hf[:npar, :npar], Hf[:npar, npar:], np.diag(Hf[npar:, npar:])
    return Hpp - (Hpl / hl) @ Hpl.T, gf[:npar] - (Hpl / hl) @ gf[npar:], hl, gf[npar:], cost


def full_system(w, G=(0, 0, 9.81007), sqrt_info=460.0 / 1.5, cauchy_a=1.0):
    """Dense weighted Jacobian/residual of the whole window in local coordinates
    [15*K pose/speed-bias | L inverse depths]; extrinsics fixed.  Returns J, r, cost."""
    K, L = w.K, len(w.inv_depth)
    npar = 15 * K
    rows, res = [], []
    cost = 0.0
    for l in range(L):
        o0, o1 = w.lm_obs_offset[l], w.lm_obs_offset[l + 1]
        fi = w.obs_frame[o0]
        for k in range(o0 + 1, o1):
            fj = w.obs_frame[k]
            r, Ji, Jj, _, Jf = projection(w.obs_xy[o0], w.obs_xy[k], w.para_pose[fi], w.para_pose[fj],
                                          w.para_ex_pose, w.inv_depth[l], sqrt_info)
            s = r @ r
            rho0 = cauchy_a ** 2 * np.log1p(s / cauchy_a ** 2)
            rho1 = 1.0 / (1.0 + s / cauchy_a ** 2)
            cost += 0.5 * rho0
            sr = np.sqrt(rho1)
            blk = np.zeros((2, npar + L))
            blk[:, 15 * fi:15 * fi + 6] += sr * Ji
            blk[:, 15 * fj:15 * fj + 6] += sr * Jj
            blk[:, npar + l] = sr * Jf
            rows.append(blk)
            res.append(sr * r)
    for j in range(1, K):
        pre = unpack_preint(w.preint[j])
        if pre["sum_dt"] > 10.0:
            continue
        i = j - 1
        r, Jpi, Jsi, Jpj, Jsj = imu(pre, G, w.para_pose[i], w.para_speed_bias[i], w.para_pose[j], w.para_speed_bias[j])
        blk = np.zeros((15, npar + L))
        blk[:, 15 * i:15 * i + 6] = Jpi
        blk[:, 15 * i + 6:15 * i + 15] = Jsi
        blk[:, 15 * j:15 * j + 6] = Jpj
        blk[:, 15 * j + 6:15 * j + 15] = Jsj
        rows.append(blk)
        res.append(r)
        cost += 0.5 * r @ r
    if w.prior is not None:
        r, dx, Jm = prior_residual(w.prior, w)
        n = w.prior["n"]
        blk = np.zeros((n, npar + L))
        for kind, frame, idx in zip(w.prior["block_kind"], w.prior["block_frame"], w.prior["block_idx"]):
            if kind == 0:
                blk[:, 15 * frame:15 * frame + 6] = Jm[:, idx:idx + 6]
            elif kind == 1:
                blk[:, 15 * frame + 6:15 * frame + 15] = Jm[:, idx:idx + 9]
        rows.append(blk)
        res.append(r)
        cost += 0.5 * r @ r
    return np.vstack(rows), np.concatenate(res), cost


def reduced_system(w, **kw):
    J, r, cost = full_system(w, **kw)
    npar = 15 * w.K
    Hf = J.T @ J
    gf = J.T @ r
    Hpp, Hpl, hl = Hf[:npar, :npar], Hf[:npar, npar:], np.diag(Hf[npar:, npar:])
    S = Hpp - (Hpl / hl) @ Hpl.T
    g = gf[:npar] - (Hpl / hl) @ gf[npar:]
    return S, g, hl, gf[npar:], cost


def apply_delta(w, d):
    """x (+) d in local coordinates; returns a new window copy."""
    K, L = w.K, len(w.inv_depth)
    c = w.copy()
    for i in range(K):
        c.para_pose[i] = pose_plus(w.para_pose[i], d[15 * i:15 * i + 6])
        c.para_speed_bias[i] = w.para_speed_bias[i] + d[15 * i + 6:15 * i + 15]
    c.inv_depth = w.inv_depth + d[15 * K:15 * K + L]
    return c


def solve_gn(w, iters=60, tol=1e-13, **kw):
    """Plain damped Gauss-Newton on the dense system (completely different solver from
    the oracle's LM/dogleg): pins the FIXED POINT the other solvers must reach."""
    lam = 1e-4
    J, r, cost = full_system(w, **kw)
    for _ in range(iters):
        H = J.T @ J
        g = J.T @ r
        if np.max(np.abs(g)) < tol:
            break
        while True:
            d = -np.linalg.solve(H + lam * np.diag(np.diag(H)), g)
            c = apply_delta(w, d)
            Jc, rc, costc = full_system(c, **kw)
            if costc < cost or lam > 1e8:
                break
            lam *= 10
        w, J, r, cost = c, Jc, rc, costc
        lam = max(lam / 10, 1e-12)
    return w, cost, np.max(np.abs(J.T @ r))


def solve_trust_region(w, strategy=0, max_iters=8, initial_radius=1e4, function_tolerance=1e-6,
                       gradient_tolerance=1e-10, parameter_tolerance=1e-8, min_relative_decrease=1e-3, **kw):
    """trust_region_loop on this module's own dense system [poses | depths] of a synth.Window.  Second opinion on the
    oracle's Schur-based loop: the two must take the same accept / reject path and produce the same iterates."""
    flat = lambda s: np.concatenate([s.para_pose.ravel(), s.para_speed_bias.ravel(), s.inv_depth])
    return trust_region_loop(w, lambda s: full_system(s, **kw), apply_delta, flat, strategy=strategy, max_iters=max_iters,
                             initial_radius=initial_radius, function_tolerance=function_tolerance,
                             gradient_tolerance=gradient_tolerance, parameter_tolerance=parameter_tolerance,
                             min_relative_decrease=min_relative_decrease)


def trust_region_loop(w, evaluate, plus, flat, strategy=0, max_iters=8, initial_radius=1e4, function_tolerance=1e-6,
                      gradient_tolerance=1e-10, parameter_tolerance=1e-8, min_relative_decrease=1e-3):
    """Ceres' trust-region control flow (TrustRegionMinimizer with LevenbergMarquardtStrategy, strategy 0, or
    DoglegStrategy / TRADITIONAL_DOGLEG, strategy 1) on DENSE, unreduced normal equations -- no Schur elimination,
    numpy's solver -- with Jacobi scaling fixed at the first linearization.  The problem is abstract:
    evaluate(state) -> (J, r, cost) with J, r already loss-corrected, plus(state, local step) -> state,
    flat(state) -> the free parameters as one vector (parameter tolerance).  Returns (state, trace, termination) with
    trace = [(cost, candidate cost, accepted, radius after)] per iteration."""
    J, r, cost = evaluate(w)
    H, g = J.T @ J, J.T @ r
    scale = 1.0 / (1.0 + np.sqrt(np.diag(H)))
    radius, decrease, mu, invalid_run = initial_radius, 2.0, 1e-8, 0
    trace, term, reuse = [], 0, False
    gn = gr = Dd = None
    alpha = step_norm = 0.0
    if np.abs(g).max() <= gradient_tolerance:
        return w, trace, 2
    it = 0
    while it < max_iters:
        it += 1
        cl = np.clip(scale * scale * np.diag(H), 1e-6, 1e32)
        valid = True
        if strategy == 0:
            dd = cl / (radius * scale * scale)
            try:
                np.linalg.cholesky(H + np.diag(dd))
                d = -np.linalg.solve(H + np.diag(dd), g)
            except np.linalg.LinAlgError:
                valid = False
        else:
            if not reuse:
                Dd = np.sqrt(cl)
                gr = scale * g / Dd                                  # gradient in the scaled y space
                t = scale * gr / Dd
                alpha = (gr @ gr) / (t @ H @ t)
                while True:
                    dd = mu * Dd * Dd / (scale * scale)
                    try:
                        np.linalg.cholesky(H + np.diag(dd))
                        dgn = -np.linalg.solve(H + np.diag(dd), g)
                        break
                    except np.linalg.LinAlgError:
                        mu *= 10.0
                        if mu > 1.0:
                            valid = False
                            break
                if valid:
                    gn = Dd * dgn / scale
                reuse = True
            if valid:
                gn_norm, g_norm = np.linalg.norm(gn), np.linalg.norm(gr)
                if gn_norm <= radius:
                    ca, cb, step_norm = 0.0, 1.0, gn_norm
                elif g_norm * alpha >= radius:
                    ca, cb, step_norm = -(radius / g_norm), 0.0, radius
                else:
                    b_dot_a = -alpha * (gr @ gn)
                    a_sq = alpha * alpha * g_norm * g_norm
                    bma_sq = a_sq - 2 * b_dot_a + gn_norm * gn_norm
                    c = b_dot_a - a_sq
                    dq = np.sqrt(c * c + bma_sq * (radius * radius - a_sq))
                    beta = (dq - c) / bma_sq if c <= 0 else (radius * radius - a_sq) / (dq + c)
                    ca, cb = -alpha * (1.0 - beta), beta
                    step_norm = np.linalg.norm(ca * gr + cb * gn)
                d = scale * (ca * gr + cb * gn) / Dd
        model = 0.0
        if valid:
            model = -(g @ d) - 0.5 * (d @ H @ d)
            valid = model > 0
        if not valid:
            invalid_run += 1
            trace.append((cost, None, False, radius))
            if invalid_run >= 5:
                term = 4
                break
            if strategy == 0:
                radius /= decrease
                decrease *= 2
            else:
                mu *= 10.0
                reuse = False
            continue
        invalid_run = 0
        cand = plus(w, d)
        Jc, rc, cc = evaluate(cand)
        x, xc = flat(w), flat(cand)
        if np.linalg.norm(x - xc) <= parameter_tolerance * (np.linalg.norm(x) + parameter_tolerance):
            term = 3
            break
        if abs(cost - cc) <= function_tolerance * cost:
            term = 1
            break
        rho = (cost - cc) / model
        if np.isfinite(cc) and rho > min_relative_decrease:
            w, J, r, cost_prev, cost = cand, Jc, rc, cost, cc
            H, g = J.T @ J, J.T @ r
            if strategy == 0:
                tt = 2.0 * rho - 1.0
                radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - tt ** 3))
                decrease = 2.0
            else:
                if rho < 0.25:
                    radius *= 0.5
                if rho > 0.75:
                    radius = max(radius, 3.0 * step_norm)
                mu = max(1e-8, 2.0 * mu / 10.0)
                reuse = False
            trace.append((cost_prev, cc, True, radius))
            if np.abs(g).max() <= gradient_tolerance:
                term = 2
                break
        else:
            if strategy == 0:
                radius /= decrease
                decrease *= 2
            else:
                radius *= 0.5
                reuse = True
            trace.append((cost, cc, False, radius))
    return w, trace, term


# ---------------------------------------------------------------------------
# selector
# ---------------------------------------------------------------------------
def slerp(a, t, b):
    d = float(np.dot(a, b))
    ad = abs(d)
    if ad >= 1.0 - np.finfo(float).eps:
        s0, s1 = 1.0 - t, t
    else:
        th = np.arccos(ad)
        s0, s1 = np.sin((1 - t) * th) / np.sin(th), np.sin(t * th) / np.sin(th)
    if d < 0:
        s1 = -s1
    return s0 * a + s1 * b


def linear_imu_matrices(Qi, Qj, nr, dImu, accVar, biasVar):
    Nij, Mij = np.zeros((3, 3)), np.zeros((3, 3))
    c11 = c12 = 0.0
    for i in range(nr):
        R = R_of(slerp(Qi, i / nr, Qj))
        jkh = nr - i - 0.5
        Nij += jkh * R
        Mij += R
        c11 += jkh * jkh
        c12 += jkh
    cov = np.zeros((9, 9))
    I = np.eye(3)
    cov[0:3, 0:3] = I * nr * c11 * dImu ** 4 * accVar
    cov[0:3, 3:6] = I * c12 * dImu ** 3 * accVar
    cov[3:6, 0:3] = cov[0:3, 3:6].T
    cov[3:6, 3:6] = I * nr * dImu ** 2 * accVar
    cov[6:9, 6:9] = I * nr * biasVar
    A = -np.eye(9)
    A[0:3, 3:6] = -I * nr * dImu
    A[0:3, 6:9] = Nij * dImu ** 2
    A[3:6, 6:9] = Mij * dImu
    return np.linalg.inv(cov), A, cov


def omega_imu(p):
    H = p.H
    D = 9 * (H + 1)
    Om = np.zeros((D, D))
    for h in range(1, H + 1):
        W, A, _ = linear_imu_matrices(p.horizon_quat[h - 1], p.horizon_quat[h], p.nr_imu, p.delta_imu,
                                      p.acc_var, p.acc_bias_var)
        a, b = 9 * (h - 1), 9 * h
        Om[a:a + 9, a:a + 9] += A.T @ W @ A
        Om[a:a + 9, b:b + 9] += A.T @ W
        Om[b:b + 9, a:a + 9] += (A.T @ W).T
        Om[b:b + 9, b:b + 9] += W
    Om[:9, :9] += np.eye(9)
    return Om


def space_to_plane(cam, P):
    mx, my = P[0] / P[2], P[1] / P[2]
    k1, k2, p1, p2 = cam["k1"], cam["k2"], cam["p1"], cam["p2"]
    rho2 = mx * mx + my * my
    rad = k1 * rho2 + k2 * rho2 * rho2
    dx = mx * rad + 2 * p1 * mx * my + p2 * (rho2 + 2 * mx * mx)
    dy = my * rad + 2 * p2 * mx * my + p1 * (rho2 + 2 * my * my)
    return np.array([cam["fx"] * (mx + dx) + cam["cx"], cam["fy"] * (my + dy) + cam["cy"]])


def feature_C(p, xy):
    """Compact 3H x 3H information block of one feature (None when numVisible == 1)."""
    H = p.H
    Ric = R_of(p.q_ic)
    R1 = R_of(p.horizon_quat[1])
    t1 = p.horizon_pos[1] + R1 @ p.t_ic
    Rwc1 = R1 @ Ric
    if len(p.cloud_depth) == 0:
        d = 1.0
    else:
        d2 = ((p.cloud_xy - xy) ** 2).sum(1)
        d = p.cloud_depth[int(np.argmin(d2))]
    f = np.array([xy[0], xy[1], 1.0])
    f = f / np.linalg.norm(f) * d
    pell = t1 + Rwc1 @ f
    Ch = [np.zeros((3, 3)) for _ in range(H)]
    nvis = 1
    for h in range(2, H + 1):
        Rh = R_of(p.horizon_quat[h])
        th = p.horizon_pos[h] + Rh @ p.t_ic
        Rwch = Rh @ Ric
        u = Rwch.T @ (pell - th)
        u = u / np.linalg.norm(u)
        px = space_to_plane(p.cam, u)
        uu, vv = int(np.round(px[0])), int(np.round(px[1]))
        if not (0 <= uu < p.cam["width"] and 0 <= vv < p.cam["height"]):
            continue
        B = skew(u) @ (Rwch @ Ric).T      # the reference's double q_IC (feature_selector.cpp:304)
        Ch[h - 1] = B.T @ B
        nvis += 1
    if nvis == 1:
        return None
    B = skew(f / np.linalg.norm(f)) @ (Rwc1 @ Ric).T
    Ch[0] = B.T @ B
    W = np.linalg.inv(sum(Ch))
    G = np.vstack(Ch)
    Cm = -G @ W @ G.T
    for h in range(H):
        Cm[3 * h:3 * h + 3, 3 * h:3 * h + 3] += Ch[h]
    return Cm


def pos_index(H):
    return np.concatenate([9 * h + np.arange(3) for h in range(1, H + 1)])


def greedy_select(p, exhaustive=True):
    """Greedy log-det selection on dense matrices with numpy slogdet (every candidate
    scored every round).  Returns ids, values, per-round margins."""
    H = p.H
    D = 9 * (H + 1)
    P = pos_index(H)
    M = omega_imu(p)
    for xy in p.used_xy:
        Cm = feature_C(p, xy)
        if Cm is not None:
            M[np.ix_(P, P)] += Cm
    Cs = [feature_C(p, xy) for xy in p.cand_xy]
    alive = [i for i, c in enumerate(Cs) if c is not None]
    ids, vals, margins = [], [], []
    for _ in range(p.kappa):
        best, bi, second = -1.0, -1, -np.inf
        for i in alive:
            A = M.copy()
            A[np.ix_(P, P)] += p.cand_prob[i] * Cs[i]
            v = np.linalg.slogdet(A)[1]
            if v > best:
                second, best, bi = best, v, i
            elif v > second:
                second = v
        if bi < 0:
            continue
        M[np.ix_(P, P)] += p.cand_prob[bi] * Cs[bi]
        alive.remove(bi)
        ids.append(int(p.cand_id[bi]))
        vals.append(best)
        margins.append(best - second)
    return ids, vals, margins
