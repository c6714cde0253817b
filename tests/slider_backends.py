"""Oracle backend for the closed-loop slider (tests only: the product backend is slider.GpuBackend)."""
import ctypes as C
import dataclasses

import numpy as np


class OracleBackend:
    def __init__(self, orc, abi):
        self.orc, self.abi = orc, abi
        self.t_call = 0.0

    def optimize(self, w, opts):
        abi = self.abi
        h, o, s = abi.WindowHandle(w), abi.default_opts(**opts), abi.Summary()
        assert self.orc.oracle_optimize(C.byref(h.s), C.byref(o), C.byref(s)) == 0
        return dataclasses.replace(w, para_pose=h.pose, para_speed_bias=h.sb, para_ex_pose=h.ex, para_td=h.td,
                                   inv_depth=h.inv), s.as_dict()

    def triangulate(self, w, init_depth):
        h = self.abi.WindowHandle(w)
        d = np.zeros(w.L)
        assert self.orc.oracle_triangulate(C.byref(h.s), init_depth, self.abi.dptr(d)) == 0
        return d

    def horizon_imu(self, H, pos0, quat0, ba0, pos1, quat1, vel1, acc, gyr, nr_imu, delta_imu):
        f = lambda a: np.ascontiguousarray(a, np.float64)
        pos, quat = np.zeros((H + 1, 3)), np.zeros((H + 1, 4))
        args = [self.abi.dptr(f(a)) for a in (pos0, quat0, ba0, pos1, quat1, vel1, acc, gyr)]
        self.orc.oracle_horizon_imu(H, *args, nr_imu, delta_imu, self.abi.dptr(pos), self.abi.dptr(quat))
        return pos, quat

    def marginalize(self, w, flag, opts=None):
        return self.abi.call_marginalize(self.orc.oracle_marginalize, w, flag, opts=self.abi.default_opts(**(opts or {})))

    def omega_prior(self, w, opts=None):
        h, om = self.abi.WindowHandle(w), np.zeros(81)
        assert self.orc.oracle_window_omega_prior(C.byref(h.s), C.byref(self.abi.default_opts(**(opts or {}))), self.abi.dptr(om)) == 0
        return om.reshape(9, 9)

    def select(self, prob):
        abi = self.abi
        h, ss = abi.SelectHandle(prob), abi.SelectSummary()
        ids = np.zeros(max(prob.kappa, 1), np.int32)
        assert self.orc.oracle_select(C.byref(h.s), abi.iptr(ids), None, C.byref(ss)) == 0
        return ids[:ss.n_selected].copy()
