"""Row f2: FeatureManager::triangulate (feature_manager.cpp:202-257) -- oracle vs numpy's SVD and vs the true depth on
the CPU; bvio_triangulate vs the oracle on the GPU."""
import ctypes as C

import numpy as np
import pytest


def _tracks(pkg, w):
    sl = pkg.slider
    return [sl.Track(lid=l, start=int(w.obs_frame[w.lm_obs_offset[l]]),
                     xy=[w.obs_xy[k] for k in range(w.lm_obs_offset[l], w.lm_obs_offset[l + 1])]) for l in range(w.L)]


def _ric(pkg):
    U_, _, Vt_ = np.linalg.svd(pkg.synth.EUROC_RIC)
    return U_ @ Vt_


@pytest.mark.parametrize("seed,noise", [(0, False), (1, True), (2, True)])
def test_oracle_triangulate_matches_numpy_svd(pkg, oracle, seed, noise):
    abi, S = pkg.abi, pkg.synth
    w = S.make_window(seed=seed, K=11, L=80, noise=noise, perturb=noise)
    h = abi.WindowHandle(w)
    d = np.zeros(w.L)
    assert oracle.oracle_triangulate(C.byref(h.s), 5.0, abi.dptr(d)) == 0
    ref = np.array([pkg.slider.triangulate(tr, w.para_pose, _ric(pkg), S.EUROC_TIC) for tr in _tracks(pkg, w)])
    assert np.abs(d - ref).max() <= 1e-8 * np.abs(ref).max()
    if not noise:       # exact poses and observations: the DLT returns the true depth
        assert np.abs(d - 1.0 / w.gt_inv_depth).max() <= 1e-7 * (1.0 / w.gt_inv_depth).max()


def test_oracle_triangulate_falls_back_to_init_depth(pkg, oracle):
    """a landmark seen twice from the same place has no parallax: depth < 0.1 (or undefined) -> INIT_DEPTH"""
    import dataclasses
    abi, S = pkg.abi, pkg.synth
    w = S.make_window(seed=3, K=4, L=6, track_min=2, track_max=2, noise=False, perturb=False)
    pose = w.para_pose.copy()
    pose[:] = pose[0]                        # all frames coincide
    xy = w.obs_xy.copy()
    xy[1::2] = xy[0::2] + 1e-3               # inconsistent second observation: triangulates behind / at infinity
    w2 = dataclasses.replace(w, para_pose=pose, obs_xy=xy)
    h = abi.WindowHandle(w2)
    d = np.zeros(w.L)
    assert oracle.oracle_triangulate(C.byref(h.s), 5.0, abi.dptr(d)) == 0
    assert np.isfinite(d).all() and ((d == 5.0) | (d >= 0.1)).all() and (d == 5.0).any()


@pytest.mark.gpu
@pytest.mark.parametrize("seed,K,L", [(0, 11, 150), (1, 11, 1500), (2, 2, 20), (3, 15, 64)])
def test_cuda_triangulate_matches_oracle(pkg, oracle, seed, K, L):
    abi, S = pkg.abi, pkg.synth
    ctx = pkg.lib.Context(0)
    kw = dict(track_min=2, track_max=2) if K == 2 else {}
    w = S.make_window(seed=seed, K=K, L=L, **kw)
    h = abi.WindowHandle(w)
    dg, do = np.zeros(w.L), np.zeros(w.L)
    ctx.check(ctx.L.bvio_triangulate(ctx.h, C.byref(h.s), 5.0, abi.dptr(dg)), "bvio_triangulate")
    assert oracle.oracle_triangulate(C.byref(h.s), 5.0, abi.dptr(do)) == 0
    fb = do == 5.0
    assert ((dg == 5.0) == fb).all()
    assert np.abs(dg - do)[~fb].max() <= 1e-9 * np.abs(do).max()
    ctx.close()
