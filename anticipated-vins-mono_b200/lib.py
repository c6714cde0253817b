"""Loader + argtypes for csrc/libbvio.so (the product).  Fails loudly when the CUDA
library is missing -- there is no CPU path behind these calls."""
from __future__ import annotations

import ctypes as C
import os

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# BVIO_LIB_PATH: load another build of the same library (kernel A/B measurements); never a different implementation
LIB_PATH = os.environ.get("BVIO_LIB_PATH") or os.path.join(_HERE, "csrc", "libbvio.so")
_lib = None

# every symbol include/bvio.h declares
EXPORTS = [
    "bvio_create", "bvio_destroy", "bvio_last_error", "bvio_abi_version", "bvio_default_opts",
    "bvio_optimize", "bvio_optimize_batch", "bvio_batch_upload", "bvio_batch_solve", "bvio_batch_download",
    "bvio_batch_free", "bvio_batch_solve_timed", "bvio_stream", "bvio_launch_count", "bvio_marginalize", "bvio_select",
    "bvio_nccl_unique_id", "bvio_comm_init", "bvio_select_sharded", "bvio_select_upload", "bvio_select_run",
    "bvio_select_fetch", "bvio_select_free", "bvio_debug_linearize", "bvio_debug_build_delta", "bvio_triangulate",
    "bvio_preintegrate", "bvio_horizon_imu", "bvio_select_upload_mode", "bvio_marginalize_begin", "bvio_marginalize_end", "bvio_window_omega_prior",
]


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, dp, ip = C.c_void_p, C.c_int32, abi.c_double_p, abi.c_int32_p
    L.bvio_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.bvio_destroy.argtypes = [vp]
    L.bvio_destroy.restype = None
    L.bvio_last_error.argtypes = [vp]
    L.bvio_last_error.restype = C.c_char_p
    L.bvio_default_opts.argtypes = [C.POINTER(abi.Opts)]
    L.bvio_default_opts.restype = None
    L.bvio_optimize.argtypes = [vp, C.POINTER(abi.WindowS), C.POINTER(abi.Opts), C.POINTER(abi.Summary)]
    L.bvio_optimize_batch.argtypes = [vp, C.POINTER(abi.WindowS), i32, C.POINTER(abi.Opts), C.POINTER(abi.Summary)]
    L.bvio_batch_upload.argtypes = [vp, C.POINTER(abi.WindowS), i32, C.POINTER(abi.Opts), C.POINTER(vp)]
    L.bvio_batch_solve.argtypes = [vp, vp]
    L.bvio_batch_download.argtypes = [vp, vp, C.POINTER(abi.WindowS), C.POINTER(abi.Summary)]
    L.bvio_batch_solve_timed.argtypes = [vp, vp, dp, ip]
    L.bvio_batch_free.argtypes = [vp, vp]
    L.bvio_batch_free.restype = None
    L.bvio_stream.argtypes = [vp]
    L.bvio_stream.restype = vp
    L.bvio_launch_count.argtypes = [vp]
    L.bvio_launch_count.restype = C.c_int64
    L.bvio_marginalize.argtypes = [vp, C.POINTER(abi.WindowS), C.POINTER(abi.Opts), i32, C.POINTER(abi.PriorOut)]
    L.bvio_marginalize_begin.argtypes = [vp, C.POINTER(abi.WindowS), C.POINTER(abi.Opts), i32, C.POINTER(abi.PriorOut), C.POINTER(vp)]
    L.bvio_marginalize_end.argtypes = [vp, vp]
    L.bvio_window_omega_prior.argtypes = [vp, C.POINTER(abi.WindowS), C.POINTER(abi.Opts), dp]
    L.bvio_select.argtypes = [vp, C.POINTER(abi.SelectIn), ip, dp, C.POINTER(abi.SelectSummary)]
    L.bvio_nccl_unique_id.argtypes = [vp]
    L.bvio_comm_init.argtypes = [vp, vp, i32, i32]
    L.bvio_select_sharded.argtypes = [vp, C.POINTER(abi.SelectIn), ip, dp, C.POINTER(abi.SelectSummary)]
    L.bvio_select_upload.argtypes = [vp, C.POINTER(abi.SelectIn), C.POINTER(vp)]
    L.bvio_select_upload_mode.argtypes = [vp, C.POINTER(abi.SelectIn), C.c_int32, C.POINTER(vp)]
    L.bvio_select_run.argtypes = [vp, vp]
    L.bvio_select_fetch.argtypes = [vp, vp, ip, dp, C.POINTER(abi.SelectSummary)]
    L.bvio_select_free.argtypes = [vp, vp]
    L.bvio_select_free.restype = None
    L.bvio_debug_linearize.argtypes = [vp, C.POINTER(abi.WindowS), C.POINTER(abi.Opts), dp, dp, dp, dp, dp]
    L.bvio_debug_build_delta.argtypes = [vp, C.POINTER(abi.SelectIn), dp, ip, dp]
    L.bvio_triangulate.argtypes = [vp, C.POINTER(abi.WindowS), C.c_double, dp]
    L.bvio_horizon_imu.argtypes = [vp, i32, dp, dp, dp, dp, dp, dp, dp, dp, i32, C.c_double, dp, dp]
    L.bvio_preintegrate.argtypes = [vp, C.POINTER(abi.ImuSegment), i32, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.POINTER(abi.Preint)]
    _lib = L
    return L


class Context:
    """RAII wrapper around bvio_ctx (single caller)."""

    def __init__(self, device: int = 0):
        self.L = load()
        self.h = C.c_void_p()
        rc = self.L.bvio_create(device, C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"bvio_create failed rc={rc} (no CUDA device? there is no CPU fallback)")

    def check(self, rc, what):
        if rc != 0:
            msg = self.L.bvio_last_error(self.h)
            raise RuntimeError(f"{what} failed rc={rc}: {msg.decode() if msg else ''}")

    def close(self):
        if self.h:
            self.L.bvio_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
