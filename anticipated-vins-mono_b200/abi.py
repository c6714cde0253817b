"""ctypes mirror of include/bvio.h (struct layouts only, no compute).

Used by the Python host-side tests/bench to call the C-ABI of libbvio.so (the
product) and, in tests only, liboracle.so (the checker) with identical inputs.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Preint(C.Structure):
    _fields_ = [("delta_p", C.c_double * 3), ("delta_q", C.c_double * 4), ("delta_v", C.c_double * 3),
                ("lin_ba", C.c_double * 3), ("lin_bg", C.c_double * 3), ("sum_dt", C.c_double),
                ("jacobian", C.c_double * 225), ("covariance", C.c_double * 225)]


assert C.sizeof(Preint) == 467 * 8


class ImuSegment(C.Structure):
    _fields_ = [("n_samples", C.c_int32), ("dt", c_double_p), ("acc", c_double_p), ("gyr", c_double_p),
                ("lin_ba", C.c_double * 3), ("lin_bg", C.c_double * 3)]


class Prior(C.Structure):
    _fields_ = [("n", C.c_int32), ("nblocks", C.c_int32), ("block_kind", c_int32_p), ("block_frame", c_int32_p),
                ("block_idx", c_int32_p), ("x0", c_double_p), ("lin_jac", c_double_p), ("lin_res", c_double_p)]


class WindowS(C.Structure):
    _fields_ = [("K", C.c_int32), ("para_pose", c_double_p), ("para_speed_bias", c_double_p),
                ("para_ex_pose", c_double_p), ("para_td", c_double_p), ("L", C.c_int32),
                ("inv_depth", c_double_p), ("lm_obs_offset", c_int32_p), ("obs_frame", c_int32_p),
                ("obs_xy", c_double_p), ("obs_vel", c_double_p), ("obs_td", c_double_p), ("obs_row", c_double_p),
                ("preint", C.POINTER(Preint)), ("prior", C.POINTER(Prior)),
                ("n_relo", C.c_int32), ("relo_pose", c_double_p), ("relo_lm", c_int32_p), ("relo_xy", c_double_p)]


class Opts(C.Structure):
    _fields_ = [("max_iters", C.c_int32), ("max_time_s", C.c_double), ("estimate_extrinsic", C.c_int32),
                ("estimate_td", C.c_int32), ("focal_length", C.c_double), ("cauchy_a", C.c_double),
                ("G", C.c_double * 3), ("TR", C.c_double), ("ROW", C.c_double),
                ("function_tolerance", C.c_double), ("gradient_tolerance", C.c_double),
                ("parameter_tolerance", C.c_double), ("initial_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("strategy", C.c_int32), ("jacobi_scaling", C.c_int32)]


class Summary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("num_accepted", C.c_int32), ("num_rejected", C.c_int32),
                ("termination", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("final_radius", C.c_double), ("final_gradient_max", C.c_double), ("device_ms", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class PriorOut(C.Structure):
    _fields_ = [("n", C.c_int32), ("nblocks", C.c_int32), ("block_kind", c_int32_p), ("block_frame", c_int32_p),
                ("block_idx", c_int32_p), ("x0", c_double_p), ("lin_jac", c_double_p), ("lin_res", c_double_p),
                ("cap_n", C.c_int32), ("cap_blocks", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k1", C.c_double), ("k2", C.c_double), ("p1", C.c_double), ("p2", C.c_double),
                ("width", C.c_int32), ("height", C.c_int32)]


class SelectIn(C.Structure):
    _fields_ = [("H", C.c_int32), ("horizon_pos", c_double_p), ("horizon_quat", c_double_p),
                ("q_ic", C.c_double * 4), ("t_ic", C.c_double * 3), ("cam", Camera),
                ("nr_imu", C.c_int32), ("delta_imu", C.c_double), ("acc_var", C.c_double),
                ("acc_bias_var", C.c_double),
                ("N", C.c_int32), ("cand_id", c_int32_p), ("cand_xy", c_double_p), ("cand_prob", c_double_p),
                ("U", C.c_int32), ("used_id", c_int32_p), ("used_xy", c_double_p),
                ("C", C.c_int32), ("cloud_xy", c_double_p), ("cloud_depth", c_double_p),
                ("kappa", C.c_int32),
                ("state_k1_pos", c_double_p), ("state_k1_quat", c_double_p), ("omega_prior", c_double_p)]


class SelectSummary(C.Structure):
    _fields_ = [("n_selected", C.c_int32), ("n_candidates_valid", C.c_int32), ("candidates_scored", C.c_int64),
                ("final_logdet", C.c_double), ("min_margin", C.c_double), ("device_ms", C.c_double),
                ("transport", C.c_int32), ("world", C.c_int32), ("grid", C.c_int32), ("cpw", C.c_int32),
                ("round_score_us", C.c_double), ("round_barrier_us", C.c_double), ("round_exchange_us", C.c_double)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


def dptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def iptr(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int32_p)


def default_opts(**kw) -> Opts:
    """EuRoC values: config/euroc/euroc_config.yaml:54-63, parameters.h:13; Ceres 1.14 defaults."""
    o = Opts()
    o.max_iters = 8
    o.max_time_s = 0.0
    o.estimate_extrinsic = 0
    o.estimate_td = 0
    o.focal_length = 460.0
    o.cauchy_a = 1.0
    o.G[0], o.G[1], o.G[2] = 0.0, 0.0, 9.81007
    o.TR, o.ROW = 0.0, 480.0
    o.function_tolerance = 1e-6
    o.gradient_tolerance = 1e-10
    o.parameter_tolerance = 1e-8
    o.initial_radius = 1e4
    o.min_relative_decrease = 1e-3
    o.strategy = 0
    o.jacobi_scaling = 1
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class WindowHandle:
    """Owns the numpy buffers a bvio_window points into (state arrays are copies: in/out)."""

    def __init__(self, w):
        self.src = w
        self.pose = np.ascontiguousarray(w.para_pose, np.float64).copy()
        self.sb = np.ascontiguousarray(w.para_speed_bias, np.float64).copy()
        self.ex = np.ascontiguousarray(w.para_ex_pose, np.float64).copy()
        self.td = np.ascontiguousarray(w.para_td, np.float64).copy()
        self.inv = np.ascontiguousarray(w.inv_depth, np.float64).copy()
        self.off = np.ascontiguousarray(w.lm_obs_offset, np.int32)
        self.frame = np.ascontiguousarray(w.obs_frame, np.int32)
        self.xy = np.ascontiguousarray(w.obs_xy, np.float64)
        self.pre = np.ascontiguousarray(w.preint, np.float64)
        s = WindowS()
        s.K = w.K
        s.para_pose, s.para_speed_bias = dptr(self.pose), dptr(self.sb)
        s.para_ex_pose, s.para_td = dptr(self.ex), dptr(self.td)
        s.L = len(self.inv)
        s.inv_depth = dptr(self.inv)
        s.lm_obs_offset, s.obs_frame, s.obs_xy = iptr(self.off), iptr(self.frame), dptr(self.xy)
        s.obs_vel = s.obs_td = s.obs_row = None
        if getattr(w, "obs_vel", None) is not None:
            self.vel = np.ascontiguousarray(w.obs_vel, np.float64)
            self.otd = np.ascontiguousarray(w.obs_td, np.float64)
            self.row = np.ascontiguousarray(w.obs_row, np.float64)
            s.obs_vel, s.obs_td, s.obs_row = dptr(self.vel), dptr(self.otd), dptr(self.row)
        s.preint = C.cast(self.pre.ctypes.data, C.POINTER(Preint))
        self.prior_s = None
        if w.prior is not None:
            p = w.prior
            self._pk = np.ascontiguousarray(p["block_kind"], np.int32)
            self._pf = np.ascontiguousarray(p["block_frame"], np.int32)
            self._pi = np.ascontiguousarray(p["block_idx"], np.int32)
            self._px0 = np.ascontiguousarray(p["x0"], np.float64)
            self._pj = np.ascontiguousarray(p["lin_jac"], np.float64)
            self._pr = np.ascontiguousarray(p["lin_res"], np.float64)
            ps = Prior()
            ps.n, ps.nblocks = int(p["n"]), len(self._pk)
            ps.block_kind, ps.block_frame, ps.block_idx = iptr(self._pk), iptr(self._pf), iptr(self._pi)
            ps.x0, ps.lin_jac, ps.lin_res = dptr(self._px0), dptr(self._pj), dptr(self._pr)
            self.prior_s = ps
            s.prior = C.pointer(ps)
        else:
            s.prior = None
        # ABI v2: relocalization matches (estimator.cpp:760-792)
        s.n_relo, s.relo_pose, s.relo_lm, s.relo_xy = 0, None, None, None
        if getattr(w, "relo_pose", None) is not None:
            self.relo_pose = np.ascontiguousarray(w.relo_pose, np.float64).copy()
            self.relo_lm = np.ascontiguousarray(w.relo_lm, np.int32)
            self.relo_xy = np.ascontiguousarray(w.relo_xy, np.float64).reshape(-1, 2)
            s.n_relo, s.relo_pose = len(self.relo_lm), dptr(self.relo_pose)
            s.relo_lm = iptr(self.relo_lm) if s.n_relo else None
            s.relo_xy = dptr(self.relo_xy) if s.n_relo else None
        self.s = s

    def state_vector(self):
        return np.concatenate([self.pose.ravel(), self.sb.ravel(), self.ex, self.inv])


class SelectHandle:
    def __init__(self, p):
        self.src = p
        f64 = lambda a: np.ascontiguousarray(a, np.float64)
        i32 = lambda a: np.ascontiguousarray(a, np.int32)
        self.pos, self.quat = f64(p.horizon_pos), f64(p.horizon_quat)
        self.cid, self.cxy, self.cp = i32(p.cand_id), f64(p.cand_xy), f64(p.cand_prob)
        self.uid, self.uxy = i32(p.used_id), f64(p.used_xy).reshape(-1, 2)
        self.clxy, self.cld = f64(p.cloud_xy).reshape(-1, 2), f64(p.cloud_depth)
        s = SelectIn()
        s.H = p.H
        s.horizon_pos, s.horizon_quat = dptr(self.pos), dptr(self.quat)
        for i in range(4):
            s.q_ic[i] = p.q_ic[i]
        for i in range(3):
            s.t_ic[i] = p.t_ic[i]
        for k in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "width", "height"):
            setattr(s.cam, k, p.cam[k])
        s.nr_imu, s.delta_imu, s.acc_var, s.acc_bias_var = p.nr_imu, p.delta_imu, p.acc_var, p.acc_bias_var
        s.N, s.cand_id, s.cand_xy, s.cand_prob = len(self.cid), iptr(self.cid), dptr(self.cxy), dptr(self.cp)
        s.U = len(self.uid)
        s.used_id = iptr(self.uid) if s.U else None
        s.used_xy = dptr(self.uxy) if s.U else None
        s.C = len(self.cld)
        s.cloud_xy = dptr(self.clxy) if s.C else None
        s.cloud_depth = dptr(self.cld) if s.C else None
        s.kappa = p.kappa
        # ABI v2 optional inputs
        self.k1p = f64(p.state_k1_pos) if getattr(p, "state_k1_pos", None) is not None else None
        self.k1q = f64(p.state_k1_quat) if getattr(p, "state_k1_quat", None) is not None else None
        self.omp = f64(p.omega_prior).reshape(81) if getattr(p, "omega_prior", None) is not None else None
        s.state_k1_pos = dptr(self.k1p) if self.k1p is not None else None
        s.state_k1_quat = dptr(self.k1q) if self.k1q is not None else None
        s.omega_prior = dptr(self.omp) if self.omp is not None else None
        self.s = s


def call_marginalize(lib_fn, w, flag, ctx=None, opts=None):
    """bvio_marginalize / oracle_marginalize on a synth.Window; returns the next prior as the dict synth.Window.prior
    expects (plus J, the n x n linearized_jacobians), or None when the reference would keep the old prior."""
    import time
    h = WindowHandle(w)
    cap_n, cap_b = 15 * w.K + 16, 2 * w.K + 2
    bk, bf, bi = (np.zeros(cap_b, np.int32) for _ in range(3))
    x0, jac, res = np.zeros(9 * cap_b), np.zeros(cap_n * cap_n), np.zeros(cap_n)
    out = PriorOut()
    out.block_kind, out.block_frame, out.block_idx = iptr(bk), iptr(bf), iptr(bi)
    out.x0, out.lin_jac, out.lin_res = dptr(x0), dptr(jac), dptr(res)
    out.cap_n, out.cap_blocks = cap_n, cap_b
    o = opts if opts is not None else default_opts()
    t0 = time.perf_counter()
    rc = lib_fn(C.byref(h.s), C.byref(o), flag, C.byref(out)) if ctx is None else \
        lib_fn(ctx, C.byref(h.s), C.byref(o), flag, C.byref(out))
    call_marginalize.t_call = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"marginalize failed: {rc}")
    n, nb = out.n, out.nblocks
    if n < 0:
        return None
    J = jac[:n * n].reshape(n, n, order="F").copy()
    return dict(n=n, block_kind=bk[:nb].copy(), block_frame=bf[:nb].copy(), block_idx=bi[:nb].copy(),
                x0=x0.copy(), lin_jac=jac[:n * n].copy(), lin_res=res[:n].copy(), J=J)


call_marginalize.t_call = 0.0


class MarginalizeJob:
    """bvio_marginalize_begin ... bvio_marginalize_end around other work (the prior is first needed by the next frame)."""

    def __init__(self, L, ctx, w, flag, opts=None):
        import time
        self.L, self.ctx = L, ctx
        self.h = WindowHandle(w)
        cap_n, cap_b = 15 * w.K + 16, 2 * w.K + 2
        self.bk, self.bf, self.bi = (np.zeros(cap_b, np.int32) for _ in range(3))
        self.x0, self.jac, self.res = np.zeros(9 * cap_b), np.zeros(cap_n * cap_n), np.zeros(cap_n)
        self.out = PriorOut()
        self.out.block_kind, self.out.block_frame, self.out.block_idx = iptr(self.bk), iptr(self.bf), iptr(self.bi)
        self.out.x0, self.out.lin_jac, self.out.lin_res = dptr(self.x0), dptr(self.jac), dptr(self.res)
        self.out.cap_n, self.out.cap_blocks = cap_n, cap_b
        self.o = opts if opts is not None else default_opts()
        self.job = C.c_void_p()
        t0 = time.perf_counter()
        rc = L.bvio_marginalize_begin(ctx, C.byref(self.h.s), C.byref(self.o), flag, C.byref(self.out), C.byref(self.job))
        self.t_begin = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError(f"marginalize_begin failed: {rc}")

    def end(self):
        import time
        t0 = time.perf_counter()
        rc = self.L.bvio_marginalize_end(self.ctx, self.job)
        self.t_end = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError(f"marginalize_end failed: {rc}")
        n, nb = self.out.n, self.out.nblocks
        if n < 0:
            return None
        J = self.jac[:n * n].reshape(n, n, order="F").copy()
        return dict(n=n, block_kind=self.bk[:nb].copy(), block_frame=self.bf[:nb].copy(), block_idx=self.bi[:nb].copy(),
                    x0=self.x0.copy(), lin_jac=self.jac[:n * n].copy(), lin_res=self.res[:n].copy(), J=J)
