"""Row f3 (selector side): HorizonGenerator::imu (utility/horizon_generator.cpp:25-70)."""
import ctypes as C

import numpy as np
import pytest


def _inputs(pkg, seed):
    S = pkg.synth
    rng = np.random.default_rng(seed)
    traj = S.Trajectory(phase=rng.uniform(0, 5))
    t0, t1 = 2.0, 2.1
    f = lambda a: np.ascontiguousarray(a, np.float64)
    return dict(pos0=f(traj.pos(t0)), quat0=f(S.rot_to_quat(traj.rot(t0))), ba0=f(rng.normal(0, 0.02, 3)),
                pos1=f(traj.pos(t1)), quat1=f(S.rot_to_quat(traj.rot(t1))), vel1=f(traj.vel(t1)),
                acc=f(traj.rot(t1).T @ (traj.acc(t1) + np.array([0, 0, 9.80665]))), gyr=f(traj.omega_body(t1)))


def _numpy(pkg, H, x, nr, delta):
    S = pkg.synth
    grav = np.array([0, 0, -9.80665])
    pos, quat = np.zeros((H + 1, 3)), np.zeros((H + 1, 4))
    pos[0], quat[0], pos[1], quat[1] = x["pos0"], x["quat0"], x["pos1"], x["quat1"]
    p, v, q = x["pos1"].copy(), x["vel1"].copy(), x["quat1"].copy()
    qimu = np.concatenate([x["gyr"] * delta / 2, [1.0]])
    a = x["acc"] - x["ba0"]
    for h in range(2, H + 1):
        for _ in range(nr):
            q = S.quat_mul(q, qimu)
            qa = S._eigen_quat_rotate(q, a)
            v = v + (grav + qa) * delta
            p = p + v * delta + 0.5 * grav * delta * delta + 0.5 * qa * delta * delta
        pos[h], quat[h] = p, q
    return pos, quat


def _call(abi, fn, H, x, nr, delta, ctx=None):
    pos, quat = np.zeros((H + 1, 3)), np.zeros((H + 1, 4))
    args = [H] + [abi.dptr(x[k]) for k in ("pos0", "quat0", "ba0", "pos1", "quat1", "vel1", "acc", "gyr")] + \
           [nr, delta, abi.dptr(pos), abi.dptr(quat)]
    rc = fn(*args) if ctx is None else fn(ctx, *args)
    assert not rc
    return pos, quat


@pytest.mark.parametrize("seed,H,nr", [(0, 10, 20), (1, 13, 20), (2, 3, 7), (3, 1, 20)])
def test_oracle_horizon_matches_numpy(pkg, oracle, seed, H, nr):
    x = _inputs(pkg, seed)
    po, qo = _call(pkg.abi, oracle.oracle_horizon_imu, H, x, nr, 0.005)
    pn, qn = _numpy(pkg, H, x, nr, 0.005)
    assert np.abs(po - pn).max() <= 1e-13 * max(np.abs(pn).max(), 1.0) and np.abs(qo - qn).max() <= 1e-13
    if H >= 3:      # the robot really moves along the horizon, roughly along its velocity
        assert np.linalg.norm(po[-1] - po[1]) > 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("seed,H,nr", [(0, 10, 20), (1, 13, 20), (2, 16, 33), (3, 1, 20)])
def test_cuda_horizon_matches_oracle(pkg, oracle, seed, H, nr):
    ctx = pkg.lib.Context(0)
    x = _inputs(pkg, seed)
    pg, qg = _call(pkg.abi, ctx.L.bvio_horizon_imu, H, x, nr, 0.005, ctx=ctx.h)
    po, qo = _call(pkg.abi, oracle.oracle_horizon_imu, H, x, nr, 0.005)
    assert np.abs(pg - po).max() <= 1e-12 * max(np.abs(po).max(), 1.0) and np.abs(qg - qo).max() <= 1e-12
    ctx.close()
