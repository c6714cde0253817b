"""ctypes loader for oracle/liboracle.so -- the CHECKER.  Only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import subprocess

import __graft_entry__ as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    abi = g.load_package().abi
    L = C.CDLL(PATH)
    dp, ip, i32, d = abi.c_double_p, abi.c_int32_p, C.c_int32, C.c_double
    L.oracle_projection_factor.argtypes = [dp, dp, dp, dp, dp, d, d, dp, dp, dp, dp, dp]
    L.oracle_projection_factor.restype = None
    L.oracle_projection_td_factor.argtypes = [dp, dp, dp, dp, d, d, d, d, d, d, dp, dp, dp, d, d, d, dp, dp, dp, dp, dp, dp]
    L.oracle_projection_td_factor.restype = None
    L.oracle_imu_factor.argtypes = [C.POINTER(abi.Preint), dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]
    L.oracle_imu_factor.restype = None
    L.oracle_imu_sqrt_info.argtypes = [dp, dp]
    L.oracle_imu_sqrt_info.restype = None
    L.oracle_prior_residual.argtypes = [C.POINTER(abi.Prior), C.POINTER(abi.WindowS), dp, dp]
    L.oracle_prior_residual.restype = None
    L.oracle_preint_propagate.argtypes = [C.POINTER(abi.Preint), d, dp, dp, dp, dp, d, d, d, d]
    L.oracle_preint_propagate.restype = None
    L.oracle_cost.argtypes = [C.POINTER(abi.WindowS), C.POINTER(abi.Opts)]
    L.oracle_cost.restype = d
    L.oracle_linearize.argtypes = [C.POINTER(abi.WindowS), C.POINTER(abi.Opts), dp, dp, dp, dp, dp]
    L.oracle_window_omega_prior.argtypes = [C.POINTER(abi.WindowS), C.POINTER(abi.Opts), dp]
    L.oracle_optimize.argtypes = [C.POINTER(abi.WindowS), C.POINTER(abi.Opts), C.POINTER(abi.Summary)]
    L.oracle_double2vector.argtypes = [dp, i32, dp, dp]
    L.oracle_double2vector.restype = None
    L.oracle_horizon_imu.argtypes = [i32, dp, dp, dp, dp, dp, dp, dp, dp, i32, d, dp, dp]
    L.oracle_horizon_imu.restype = None
    L.oracle_triangulate.argtypes = [C.POINTER(abi.WindowS), d, dp]
    L.oracle_marginalize.argtypes = [C.POINTER(abi.WindowS), C.POINTER(abi.Opts), i32, C.POINTER(abi.PriorOut)]
    L.oracle_omega_imu.argtypes = [C.POINTER(abi.SelectIn), dp]
    L.oracle_omega_imu.restype = None
    L.oracle_linear_imu_matrices.argtypes = [dp, dp, i32, d, d, d, dp, dp, dp]
    L.oracle_linear_imu_matrices.restype = None
    L.oracle_build_delta.argtypes = [C.POINTER(abi.SelectIn), i32, dp, dp, ip, dp]
    L.oracle_build_delta.restype = None
    L.oracle_select.argtypes = [C.POINTER(abi.SelectIn), ip, dp, C.POINTER(abi.SelectSummary)]
    L.oracle_select_k1.argtypes = [C.POINTER(abi.SelectIn), dp, dp, ip, dp, C.POINTER(abi.SelectSummary)]
    L.oracle_logdet.argtypes = [dp, i32]
    L.oracle_logdet.restype = d
    _lib = L
    return L
