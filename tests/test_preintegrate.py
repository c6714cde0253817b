"""Row f3 (input side of the IMU factors): IntegrationBase::{push_back, propagate, midPointIntegration}
(integration_base.h:30-158).  CPU: the C++ oracle against the numpy restatement in synth.Preintegration.
GPU: bvio_preintegrate against the oracle."""
import ctypes as C

import numpy as np
import pytest


def _samples(seed, n):
    rng = np.random.default_rng(seed)
    t = np.arange(n + 1) * 0.005
    acc = np.stack([0.3 * np.sin(3 * t + 0.1) + 0.1, 0.2 * np.cos(2 * t), 9.81 + 0.5 * np.sin(5 * t)], 1) + rng.normal(0, 0.05, (n + 1, 3))
    gyr = np.stack([0.2 * np.sin(2 * t), 0.1 * np.cos(3 * t + 0.3), 0.5 + 0.1 * np.sin(t)], 1) + rng.normal(0, 0.003, (n + 1, 3))
    dt = np.full(n + 1, 0.005) + rng.uniform(-2e-4, 2e-4, n + 1)
    return dt, acc, gyr, rng.normal(0, 0.02, 3), rng.normal(0, 0.002, 3)


def _oracle(pkg, oracle, dt, acc, gyr, ba, bg):
    abi, S = pkg.abi, pkg.synth
    p = abi.Preint()
    p.delta_q[3] = 1.0
    for i in range(3):
        p.lin_ba[i], p.lin_bg[i] = ba[i], bg[i]
    for i in range(15):
        p.jacobian[i * 15 + i] = 1.0
    for k in range(1, len(dt)):
        oracle.oracle_preint_propagate(C.byref(p), float(dt[k]), abi.dptr(acc[k - 1].copy()), abi.dptr(gyr[k - 1].copy()),
                                       abi.dptr(acc[k].copy()), abi.dptr(gyr[k].copy()), S.ACC_N, S.GYR_N, S.ACC_W, S.GYR_W)
    return np.frombuffer(bytes(p), np.float64).copy()


@pytest.mark.parametrize("seed,n", [(0, 20), (1, 7), (2, 1)])
def test_oracle_preintegration_matches_numpy(pkg, oracle, seed, n):
    S = pkg.synth
    dt, acc, gyr, ba, bg = _samples(seed, n)
    o = _oracle(pkg, oracle, dt, acc, gyr, ba, bg)
    pre = S.Preintegration(acc[0], gyr[0], ba, bg)
    for k in range(1, n + 1):
        pre.push_back(dt[k], acc[k], gyr[k])
    ref = S.pack_preint(pre)
    assert np.abs(o - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1.0)


@pytest.mark.gpu
def test_cuda_preintegration_matches_oracle(pkg, oracle):
    abi, S = pkg.abi, pkg.synth
    ctx = pkg.lib.Context(0)
    cases = [_samples(s, n) for s, n in ((0, 20), (1, 7), (2, 1), (3, 33), (4, 20), (5, 20), (6, 20), (7, 20), (8, 20), (9, 20))]
    segs = (abi.ImuSegment * len(cases))()
    keep = []
    for sg, (dt, acc, gyr, ba, bg) in zip(segs, cases):
        dt, acc, gyr = (np.ascontiguousarray(a, np.float64) for a in (dt, acc, gyr))
        keep.append((dt, acc, gyr))
        sg.n_samples, sg.dt, sg.acc, sg.gyr = len(dt), abi.dptr(dt), abi.dptr(acc), abi.dptr(gyr)
        for i in range(3):
            sg.lin_ba[i], sg.lin_bg[i] = ba[i], bg[i]
    out = (abi.Preint * len(cases))()
    ctx.check(ctx.L.bvio_preintegrate(ctx.h, segs, len(cases), S.ACC_N, S.GYR_N, S.ACC_W, S.GYR_W, out), "bvio_preintegrate")
    for i, c in enumerate(cases):
        g = np.frombuffer(bytes(out[i]), np.float64)
        o = _oracle(pkg, oracle, *c)
        assert np.abs(g[:17] - o[:17]).max() <= 1e-13 * max(np.abs(o[:17]).max(), 1.0), i
        assert np.abs(g[17:242] - o[17:242]).max() <= 1e-12 * np.abs(o[17:242]).max(), i           # jacobian
        assert np.abs(g[242:] - o[242:]).max() <= 1e-12 * np.abs(o[242:]).max(), i                 # covariance
    ctx.close()
