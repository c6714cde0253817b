"""ctypes loader for oracle/_ref/libvins_ref.so -- the REFERENCE's own factor sources (compiled from /root/reference by
`make -C oracle ref`) behind the C entry points of oracle/ref_shim/ref_driver.cpp.  Tests only."""
import ctypes as C
import os
import subprocess

import __graft_entry__ as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "oracle", "_ref", "libvins_ref.so")
PATH_BVIO = os.path.join(ROOT, "oracle", "_ref", "libvins_bvio.so")
_lib = None
_lib_bvio = None


def load():
    """None when the library is neither prebuilt nor buildable (no /root/reference)."""
    global _lib
    if _lib is None:
        _lib = _load(PATH, "ref")
    return _lib


def load_bvio():
    """oracle/_ref/libvins_bvio.so: the same reference classes and entry points with the reference-side adapter
    (adapters/vins) linked behind Estimator::optimization() / FeatureSelector::select() -> libbvio.so.  Loading it needs
    libbvio.so (and therefore the CUDA runtime), but no GPU until bvio_glue_enable() is called."""
    global _lib_bvio
    if _lib_bvio is None:
        L = _load(PATH_BVIO, "ref_bvio")
        if L is not None:
            abi = g.load_package().abi
            L.bvio_glue_enable.argtypes = [C.c_int32]
            L.bvio_glue_attach.argtypes = [C.c_void_p]
            L.bvio_glue_attach.restype = None
            L.bvio_glue_disable.restype = None
            L.bvio_glue_counts.argtypes = [abi.c_int32_p]
            L.bvio_glue_counts.restype = None
            L.bvio_glue_last_summary.argtypes = [C.POINTER(abi.Summary), C.POINTER(abi.SelectSummary)]
            L.bvio_glue_last_summary.restype = None
            L.bvio_glue_launches.restype = C.c_longlong
            L.bvio_glue_capture_created.argtypes = [C.c_int32]
            L.bvio_glue_capture_created.restype = None
        _lib_bvio = L
    return _lib_bvio


def _load(path, target):
    if not os.path.exists(path) and os.path.isdir("/root/reference/vins_estimator/src/factor"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), target])
    if not os.path.exists(path):
        return None
    abi = g.load_package().abi
    L = C.CDLL(path)
    dp, i32, d = abi.c_double_p, C.c_int32, C.c_double
    L.ref_projection_factor.argtypes = [dp, dp, dp, dp, dp, d, d, dp, dp, dp, dp, dp]
    L.ref_projection_factor.restype = None
    L.ref_projection_td_factor.argtypes = [dp, dp, dp, dp, d, d, d, d, d, d, dp, dp, dp, d, d, d, dp, dp, dp, dp, dp, dp]
    L.ref_projection_td_factor.restype = None
    L.ref_imu_factor.argtypes = [C.POINTER(abi.Preint), dp, dp, dp, dp, dp, dp, dp, dp, dp, dp]
    L.ref_imu_factor.restype = None
    L.ref_preint_propagate.argtypes = [C.POINTER(abi.Preint), d, dp, dp, dp, dp, d, d, d, d]
    L.ref_preint_propagate.restype = None
    L.ref_preintegrate.argtypes = [i32, dp, dp, dp, dp, dp, d, d, d, d, dp, dp, C.POINTER(abi.Preint)]
    L.ref_preintegrate.restype = None
    L.ref_pose_plus.argtypes = [dp, dp, dp]
    L.ref_pose_plus.restype = None
    L.ref_logdet.argtypes = [dp, i32]
    L.ref_logdet.restype = d
    L.ref_projection_block_corrected.argtypes = [dp, dp, dp, dp, dp, d, d, d, dp, dp, dp, dp, dp]
    L.ref_projection_block_corrected.restype = None
    L.ref_prior_eval.argtypes = [C.POINTER(abi.Prior), C.POINTER(abi.WindowS), dp, dp]
    L.ref_marginalize.argtypes = [C.POINTER(abi.WindowS), C.POINTER(abi.Opts), i32, C.POINTER(abi.PriorOut)]
    L.ref_horizon_length.restype = i32
    L.ref_horizon_imu.argtypes = [i32, dp, dp, dp, dp, dp, dp, dp, dp, i32, d, dp, dp]
    L.ref_horizon_imu.restype = None
    L.ref_horizon_gt_open.argtypes = [C.c_char_p]
    L.ref_horizon_gt_open.restype = C.c_void_p
    L.ref_horizon_gt.argtypes = [C.c_void_p, d, dp, dp, d, dp, dp]
    L.ref_horizon_gt.restype = None
    L.ref_horizon_gt_close.argtypes = [C.c_void_p]
    L.ref_horizon_gt_close.restype = None
    ip, vp = abi.c_int32_p, C.c_void_p
    L.ref_window_size.restype = i32
    L.ref_fm_create.argtypes = [d, d]
    L.ref_fm_create.restype = vp
    L.ref_fm_destroy.argtypes = [vp]
    L.ref_fm_destroy.restype = None
    L.ref_fm_add_frame.argtypes = [vp, i32, i32, ip, dp, d]
    L.ref_fm_last_track_num.argtypes = [vp]
    L.ref_fm_set_poses.argtypes = [vp, i32, dp, dp]
    L.ref_fm_set_poses.restype = None
    L.ref_fm_triangulate.argtypes = [vp]
    L.ref_fm_triangulate.restype = None
    L.ref_fm_feature_count.argtypes = [vp]
    L.ref_fm_get_depth_vector.argtypes = [vp, dp]
    L.ref_fm_get_depth_vector.restype = None
    L.ref_fm_set_depth.argtypes = [vp, i32, dp]
    L.ref_fm_set_depth.restype = None
    L.ref_fm_remove_failures.argtypes = [vp]
    L.ref_fm_remove_failures.restype = None
    L.ref_fm_remove_back_shift_depth.argtypes = [vp, dp, dp]
    L.ref_fm_remove_back_shift_depth.restype = None
    L.ref_fm_remove_front.argtypes = [vp, i32]
    L.ref_fm_remove_front.restype = None
    L.ref_fm_dump.argtypes = [vp, i32, ip, ip, ip, dp]
    L.ref_sel_create.argtypes = [C.POINTER(abi.Camera), dp, dp, d, d, i32, i32]
    L.ref_sel_create.restype = vp
    L.ref_sel_create_gt.argtypes = [C.POINTER(abi.Camera), dp, dp, d, d, i32, i32, C.c_char_p]
    L.ref_sel_create_gt.restype = vp
    L.ref_sel_destroy.argtypes = [vp]
    L.ref_sel_destroy.restype = None
    L.ref_sel_set_backend.argtypes = [vp, dp, dp, dp, i32, ip, ip, ip, dp, dp, ip]
    L.ref_sel_set_backend.restype = None
    L.ref_sel_select.argtypes = [vp, i32, C.c_uint, C.c_uint, dp, dp, dp, dp, dp, dp, i32, i32, ip, dp, dp, ip, ip, ip, ip]
    W, O = C.POINTER(abi.WindowS), C.POINTER(abi.Opts)
    L.ref_estimator_optimization.argtypes = [W, O, i32, W, dp, dp, dp, dp, dp, ip, ip, dp, dp, dp, dp, dp, dp, dp, dp, C.POINTER(abi.PriorOut)]
    L.ref_estimator_last_normal.argtypes = [dp, dp, i32]
    L.ref_estimator_set_relo.argtypes = [i32, ip, dp, dp, i32]
    L.ref_estimator_set_relo.restype = None
    L.ref_estimator_get_relo.argtypes = [dp, dp, dp]
    L.SOLVE_CB = C.CFUNCTYPE(None)
    L.ref_set_solve_callback.argtypes = [C.c_void_p]
    L.ref_set_solve_callback.restype = None
    L.ref_live_dims.argtypes = [ip, ip, ip]
    L.ref_live_get_state.argtypes = [dp]
    L.ref_live_get_state.restype = None
    L.ref_live_set_state.argtypes = [dp]
    L.ref_live_set_state.restype = None
    L.ref_live_plus.argtypes = [dp, dp, dp]
    L.ref_live_plus.restype = None
    L.ref_live_evaluate.argtypes = [dp, dp, dp]
    PP = C.POINTER(abi.Preint)
    L.ref_estimator_slide.argtypes = [i32, dp, dp, dp, ip, dp, dp, dp, dp, dp, d, d, d, d, i32, ip, ip, ip, dp, dp, dp, dp, PP, ip,
                                      i32, ip, ip, ip, dp]
    L.ref_estimator_process_imu.argtypes = [dp, dp, dp, i32, dp, dp, dp, dp, d, d, d, d, dp, dp, dp, PP]
    L.ref_estimator_process_imu.restype = None
    L.ref_est_create.argtypes = [dp, dp, d, i32, d, d, d, d, d, d]
    L.ref_est_create.restype = vp
    L.ref_est_release.argtypes = [vp]
    L.ref_est_release.restype = None
    L.ref_est_set_state.argtypes = [vp, i32, dp, dp]
    L.ref_est_set_state.restype = None
    L.ref_est_set_bias.argtypes = [vp, i32, dp, dp]
    L.ref_est_set_bias.restype = None
    L.ref_est_get_states.argtypes = [vp, dp, dp]
    L.ref_est_get_states.restype = None
    L.ref_est_set_nonlinear.argtypes = [vp]
    L.ref_est_set_nonlinear.restype = None
    L.ref_est_frame_count.argtypes = [vp]
    L.ref_est_prior_size.argtypes = [vp]
    L.ref_est_process_imu.argtypes = [vp, i32, dp, dp, dp]
    L.ref_est_process_imu.restype = None
    L.ref_est_process_image.argtypes = [vp, d, i32, ip, dp]
    L.ref_est_dump_features.argtypes = [vp, i32, ip, ip, ip, dp]
    return L
