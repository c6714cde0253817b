"""(De)serialisation of synthetic windows / selector problems for the golden fixtures."""
import numpy as np

import __graft_entry__ as g

_W = ("para_pose", "para_speed_bias", "para_ex_pose", "para_td", "inv_depth", "lm_obs_offset", "obs_frame", "obs_xy", "preint")
_P = ("block_kind", "block_frame", "block_idx", "x0", "lin_jac", "lin_res")
_S = ("horizon_pos", "horizon_quat", "q_ic", "t_ic", "cand_id", "cand_xy", "cand_prob", "used_id", "used_xy", "cloud_xy", "cloud_depth")
_CAM = ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "width", "height")


def window_to_dict(w):
    d = {"in_" + k: np.asarray(getattr(w, k)) for k in _W}
    d["in_K"] = np.array([w.K], np.int32)
    if getattr(w, "obs_vel", None) is not None:
        for k in ("obs_vel", "obs_td", "obs_row"):
            d["in_" + k] = np.asarray(getattr(w, k))
    d["in_has_prior"] = np.array([w.prior is not None], np.int32)
    if w.prior is not None:
        d["in_prior_n"] = np.array([w.prior["n"]], np.int32)
        for k in _P:
            d["in_prior_" + k] = np.asarray(w.prior[k])
    return d


def window_from_dict(d):
    synth = g.load_package().synth
    prior = None
    if int(d["in_has_prior"][0]):
        prior = {k: d["in_prior_" + k] for k in _P}
        prior["n"] = int(d["in_prior_n"][0])
    kw = {k: d["in_" + k] for k in _W}
    for k in ("obs_vel", "obs_td", "obs_row"):
        if "in_" + k in d:
            kw[k] = d["in_" + k]
    return synth.Window(K=int(d["in_K"][0]), prior=prior, **kw)


def select_to_dict(p):
    d = {"in_" + k: np.asarray(getattr(p, k)) for k in _S}
    d["in_cam"] = np.array([p.cam[k] for k in _CAM], np.float64)
    d["in_scalars"] = np.array([p.H, p.nr_imu, p.delta_imu, p.acc_var, p.acc_bias_var, p.kappa], np.float64)
    return d


def select_from_dict(d):
    synth = g.load_package().synth
    cam = {k: (int(v) if k in ("width", "height") else float(v)) for k, v in zip(_CAM, d["in_cam"])}
    H, nr, dimu, av, abv, kappa = d["in_scalars"]
    kw = {k: d["in_" + k] for k in _S}
    return synth.SelectProblem(H=int(H), cam=cam, nr_imu=int(nr), delta_imu=float(dimu), acc_var=float(av),
                               acc_bias_var=float(abv), kappa=int(kappa), **kw)
