"""Closed-loop sliding-window driver: host logic on the CPU oracle backend, and -- on a GPU --
the same session through libbvio.so, frame by frame against the oracle."""
import numpy as np
import pytest

from slider_backends import OracleBackend


def _run(sim, backend, frames):
    lats = []
    for _ in range(frames):
        lat = sim.step(backend)
        if lat is not None:
            lats.append(lat)
    return lats


def test_slider_bookkeeping_on_oracle(pkg, oracle):
    """30 frames: the window fills, every later frame optimises, marginalises (prior feeds the next window),
    selects new features up to the budget and slides; the estimate stays on the ground-truth trajectory."""
    sl = pkg.slider
    sim = sl.SlidingWindowSim(seed=3, max_feats=60, max_cand=80, opts=dict(max_iters=8))
    lats = _run(sim, OracleBackend(oracle, pkg.abi), 30)
    assert len(lats) == 30 - (sim.K - 1)
    assert len(sim.pose) == sim.K - 1 and len(sim.preint) == sim.K - 1        # slid, waiting for the next frame
    assert sim.prior is not None and sim.prior["n"] >= 15
    # prior blocks refer to frames of the slid window (addr_shift applied), frame-major
    assert sim.prior["block_frame"].max() <= sim.K - 2
    for lat in lats[2:]:
        assert lat["L"] >= 20 and lat["iterations"] >= 1
    # features that are parameters obey the reference's filter; observations are consecutive frames
    for tr in sim.tracks.values():
        assert 0 <= tr.start <= sim.K - 1 and tr.start + len(tr.xy) <= sim.K - 1 + 1
    errs = np.array([h[1] for h in sim.history])
    assert errs[-5:].max() < 0.25, errs
    # the per-frame budget: tracked + newly selected never exceeds max_feats
    newest = len(sim.pose)          # index the newest frame had before the slide
    n_new = sum(1 for tr in sim.tracks.values() if tr.alive)
    assert n_new <= 60


def test_slider_second_new_branch_on_oracle(pkg, oracle):
    """keyframes="parallax" at 25 Hz: the reference's addFeatureCheckParallax rule alternates MARGIN_OLD and
    MARGIN_SECOND_NEW (estimator.cpp:134-137), so slideWindowNew / removeFront and the flag-1 marginalization run in the
    loop, each prior feeding the next window."""
    sl = pkg.slider
    sim = sl.SlidingWindowSim(seed=5, max_feats=100, max_cand=150, opts=dict(max_iters=8), keyframes="parallax",
                              frame_dt=0.04)
    be = OracleBackend(oracle, pkg.abi)
    flags, prior_n, sum_dt = [], [], []
    for _ in range(44):
        lat = sim.step(be)
        if lat is not None:
            flags.append(lat["flag"])
            prior_n.append(sim.prior["n"] if sim.prior is not None else 0)
            sum_dt.append(sim.preint[-1][16])
            assert len(sim.pose) == sim.K - 1 == len(sim.preint) == len(sim.pre_obj) == len(sim.imu_buf)
            for tr in sim.tracks.values():
                assert 0 <= tr.start and tr.start + len(tr.xy) <= sim.K - 1 and len(tr.xy) >= 1
    flags = np.array(flags)
    assert (flags == 0).sum() >= 10 and (flags == 1).sum() >= 10
    assert sim.sum_of_back == (flags == 0).sum() and sim.sum_of_front == (flags == 1).sum()
    # MARGIN_SECOND_NEW drops Pose[WINDOW_SIZE - 1] from the prior (75 -> 69); the next MARGIN_OLD restores 75
    assert set(prior_n[2:]) == {69, 75}
    assert all(n == 69 for n, f in zip(prior_n[2:], flags[2:]) if f == 1)
    # after a SECOND_NEW slide the newest interval spans two camera periods (IMU samples appended, estimator.cpp:1046-1057)
    for f, sd in zip(flags, sum_dt):
        assert abs(sd - (0.08 if f == 1 else 0.04)) < 1e-9
    errs = np.array([h[1] for h in sim.history])
    assert errs.max() < 0.15, errs


def test_slide_new_remove_front_and_imu_merge(pkg):
    """FeatureManager::removeFront (feature_manager.cpp:330-351) and the preintegration merge of slideWindow's
    MARGIN_SECOND_NEW branch, on a hand-built window."""
    sl, S = pkg.slider, pkg.synth
    sim = sl.SlidingWindowSim(seed=1, keyframes="parallax")
    while len(sim.pose) < sim.K:                          # fill the window without solving
        sim._ingest()
    K, WS = sim.K, sl.WINDOW_SIZE
    xy = lambda n: [np.array([0.01 * k, 0.0]) for k in range(n)]
    sim.tracks = {
        1: sl.Track(lid=1, start=WS, xy=xy(1)),            # born in the newest frame: start moves to WS-1
        2: sl.Track(lid=2, start=WS - 1, xy=xy(2)),        # seen in WS-1 and WS: loses the WS-1 observation
        3: sl.Track(lid=3, start=WS - 1, xy=xy(1)),        # seen only in WS-1: erased
        4: sl.Track(lid=4, start=3, xy=xy(WS - 3 + 1)),    # long track through both frames: one observation fewer
        5: sl.Track(lid=5, start=2, xy=xy(4)),             # ended before WS-1: untouched
        6: sl.Track(lid=6, start=4, xy=xy(WS - 1 - 4 + 1)),  # ends exactly at WS-1: loses its last observation
    }
    newest_pose, newest_sb = sim.pose[WS].copy(), sim.sb[WS].copy()
    both = sim.imu_buf[WS - 1] + sim.imu_buf[WS]
    old = sim.pre_obj[WS - 1]
    ref = S.Preintegration(old.linearized_acc, old.linearized_gyr, old.lin_ba, old.lin_bg)
    for dt, a, g in both:                                  # one integration over the concatenated samples
        ref.push_back(dt, a, g)
    sim._slide_new()
    assert np.array_equal(sim.preint[-1], S.pack_preint(ref))
    assert len(sim.pose) == K - 1 and (sim.pose[-1] == newest_pose).all() and (sim.sb[-1] == newest_sb).all()
    assert len(sim.imu_buf[-1]) == len(both) and abs(sim.preint[-1][16] - sum(b[0] for b in both)) < 1e-12
    t = sim.tracks
    assert t[1].start == WS - 1 and len(t[1].xy) == 1
    assert t[2].start == WS - 1 and len(t[2].xy) == 1 and t[2].xy[0][0] == 0.01
    assert 3 not in t
    assert t[4].start == 3 and len(t[4].xy) == WS - 3 and t[4].xy[-1][0] == 0.01 * (WS - 3) and t[4].xy[-2][0] == 0.01 * (WS - 5)
    assert len(t[5].xy) == 4 and len(t[6].xy) == WS - 1 - 4
    assert sim.sum_of_front == 1


def test_regauge_keeps_frame0_yaw_and_position(pkg):
    sl, S = pkg.slider, pkg.synth
    rng = np.random.default_rng(0)
    w = S.make_window(seed=1, K=5, L=10)
    pose, sb = w.para_pose.copy(), w.para_speed_bias.copy()
    before = pose[0].copy()
    # apply a global yaw + translation (the gauge freedom of VIO) to the "solution"
    Rz = sl._yaw_R(17.0)
    sol = pose.copy()
    for i in range(len(sol)):
        sol[i, :3] = Rz @ pose[i, :3] + np.array([0.3, -0.2, 0.1])
        sol[i, 3:] = S.rot_to_quat(Rz @ S.quat_to_rot(pose[i, 3:]))
    sbs = sb.copy()
    sbs[:, :3] = sb[:, :3] @ Rz.T
    sl.regauge(before, sol, sbs)
    assert np.allclose(sol[:, :3], pose[:, :3], atol=1e-9)
    assert np.allclose(sbs[:, :3], sb[:, :3], atol=1e-9)
    for i in range(len(sol)):
        assert np.allclose(S.quat_to_rot(sol[i, 3:]), S.quat_to_rot(pose[i, 3:]), atol=1e-9)


def test_triangulate_recovers_depth(pkg):
    sl, S = pkg.slider, pkg.synth
    w = S.make_window(seed=2, K=8, L=30, noise=False, perturb=False)
    U_, _, Vt_ = np.linalg.svd(S.EUROC_RIC)
    ric = U_ @ Vt_
    for l in range(w.L):
        o0, o1 = w.lm_obs_offset[l], w.lm_obs_offset[l + 1]
        tr = sl.Track(lid=l, start=int(w.obs_frame[o0]), xy=[w.obs_xy[k] for k in range(o0, o1)])
        d = sl.triangulate(tr, w.gt_pose, ric, S.EUROC_TIC)
        assert abs(d - 1.0 / w.gt_inv_depth[l]) < 1e-6 * d


class DualBackend(OracleBackend):
    """Runs the CUDA library and the oracle on the SAME inputs at every call of a closed-loop session, checks parity,
    and hands the oracle's result back to the driver (so the session itself is deterministic)."""

    def __init__(self, orc, abi, gpu):
        super().__init__(orc, abi)
        self.gpu, self.n = gpu, {"optimize": 0, "marginalize": 0, "select": 0, "triangulate": 0}

    def triangulate(self, w, init_depth):
        dg, do = self.gpu.triangulate(w, init_depth), super().triangulate(w, init_depth)
        fb = do == init_depth
        assert ((dg == init_depth) == fb).all()
        assert np.abs(dg - do)[~fb].max(initial=0.0) <= 1e-9 * np.abs(do).max()
        self.n["triangulate"] += 1
        return do

    def optimize(self, w, opts):
        wg, sg = self.gpu.optimize(w.copy(), opts)
        wo, so = super().optimize(w.copy(), opts)
        assert (sg["iterations"], sg["num_accepted"], sg["num_rejected"], sg["termination"]) == \
               (so["iterations"], so["num_accepted"], so["num_rejected"], so["termination"]), (sg, so)
        for a, b in ((wg.para_pose, wo.para_pose), (wg.para_speed_bias, wo.para_speed_bias), (wg.inv_depth, wo.inv_depth)):
            assert np.linalg.norm(a - b) <= 1e-6 * np.linalg.norm(b), (self.n, sg, so)
        assert abs(sg["final_cost"] - so["final_cost"]) <= 1e-8 * so["final_cost"]
        self.n["optimize"] += 1
        return wo, so

    def marginalize(self, w, flag, opts=None):
        pg, po = self.gpu.marginalize(w, flag, opts), super().marginalize(w, flag, opts)
        assert (pg is None) == (po is None)
        if po is not None:
            assert pg["n"] == po["n"] and (pg["block_kind"] == po["block_kind"]).all()
            assert (pg["block_frame"] == po["block_frame"]).all() and (pg["block_idx"] == po["block_idx"]).all()
            assert np.allclose(pg["x0"], po["x0"], rtol=0, atol=1e-12)
            Hg, Ho = pg["J"].T @ pg["J"], po["J"].T @ po["J"]          # marginalization_factor.cpp:295-296
            gg, go = pg["J"].T @ pg["lin_res"], po["J"].T @ po["lin_res"]
            # same tolerances as tests/test_gpu_marg.py: cond(A_mm) ~ 1e8 (1e11 without any prior on frame 0) and
            # different elimination orders (analytic depth elimination vs one big pseudo-inverse)
            htol = 1e-7 if w.prior is not None else 2e-5
            assert np.abs(Hg - Ho).max() <= htol * np.abs(Ho).max(), self.n
            assert np.abs(gg - go).max() <= 5e-5 * max(np.abs(go).max(), 1.0), self.n
        self.n["marginalize"] += 1
        return po

    def select(self, prob):
        ig, io = self.gpu.select(prob), super().select(prob)
        assert len(ig) == len(io) and (ig == io).all(), (ig, io)
        self.n["select"] += 1
        return io


@pytest.mark.gpu
def test_closed_loop_session_gpu_vs_oracle_on_identical_inputs(pkg, oracle):
    """BASELINE configs[4] in miniature: 40 consecutive frames; every optimize (reference budget, prior = previous
    marginalization), marginalize and select call is issued to libbvio.so AND to the oracle on the same host buffers."""
    sl = pkg.slider
    ctx = pkg.lib.Context(0)
    sim = sl.SlidingWindowSim(seed=5, max_feats=100, max_cand=150, opts=dict(max_iters=8))
    dual = DualBackend(oracle, pkg.abi, sl.GpuBackend(ctx, pkg.abi))
    for f in range(40):
        sim.step(dual)
    assert dual.n["optimize"] == 30 and dual.n["marginalize"] == 30 and dual.n["select"] >= 10, dual.n
    assert dual.n["triangulate"] >= 10, dual.n
    assert sim.prior["n"] == 75
    ctx.close()


@pytest.mark.gpu
def test_closed_loop_session_with_non_keyframes_gpu_vs_oracle(pkg, oracle):
    """As above at 25 Hz with the reference's keyframe rule: MARGIN_OLD and MARGIN_SECOND_NEW alternate, so the device
    marginalization runs flag 1 on chained priors and optimize sees merged (two-period) preintegrations."""
    sl = pkg.slider
    ctx = pkg.lib.Context(0)
    sim = sl.SlidingWindowSim(seed=5, max_feats=100, max_cand=150, opts=dict(max_iters=8), keyframes="parallax",
                              frame_dt=0.04)
    dual = DualBackend(oracle, pkg.abi, sl.GpuBackend(ctx, pkg.abi))
    for f in range(36):
        sim.step(dual)
    assert dual.n["optimize"] == 26 and sim.sum_of_front >= 8 and sim.sum_of_back >= 8, (dual.n, sim.sum_of_front)
    assert sim.prior["n"] in (69, 75)
    ctx.close()


def test_slide_transfers_depth_to_the_new_anchor(pkg):
    """FeatureManager::removeBackShiftDepth (feature_manager.cpp:275-311): when frame 0 is dropped, a landmark anchored
    there keeps its 3-D position -- its depth is re-expressed in the camera of its next observation."""
    sl, S = pkg.slider, pkg.synth
    sim = sl.SlidingWindowSim(seed=1)
    w = S.make_window(seed=3, K=sim.K, L=30, noise=False, perturb=False)
    sim.pose, sim.sb = w.gt_pose.copy(), w.gt_speed_bias.copy()
    sim.preint = [np.zeros(S.PREINT_DOUBLES) for _ in range(sim.K)]
    world = {}
    for l in range(w.L):
        o0, o1 = int(w.lm_obs_offset[l]), int(w.lm_obs_offset[l + 1])
        tr = sl.Track(lid=l, start=int(w.obs_frame[o0]), xy=[w.obs_xy[k].copy() for k in range(o0, o1)],
                      depth=1.0 / w.gt_inv_depth[l])
        sim.tracks[l] = tr
        Rc, tc = sim._cam_pose(sim.pose[tr.start])
        world[l] = Rc @ (np.array([tr.xy[0][0], tr.xy[0][1], 1.0]) * tr.depth) + tc
    starts = {l: tr.start for l, tr in sim.tracks.items()}
    nobs = {l: len(tr.xy) for l, tr in sim.tracks.items()}
    pose_before = sim.pose.copy()
    sim._slide()
    assert len(sim.pose) == sim.K - 1 and np.array_equal(sim.pose, pose_before[1:])
    moved = 0
    for l, tr in sim.tracks.items():
        if starts[l] == 0:
            assert tr.start == 0 and len(tr.xy) == nobs[l] - 1
            Rc, tc = sim._cam_pose(sim.pose[0])                      # the old frame 1
            pw = Rc @ (np.array([tr.xy[0][0], tr.xy[0][1], 1.0]) * tr.depth) + tc
            assert np.linalg.norm(pw - world[l]) < 1e-9 * max(1.0, np.linalg.norm(world[l]))
            moved += 1
        else:
            assert tr.start == starts[l] - 1 and len(tr.xy) == nobs[l]
    assert moved > 0
    # landmarks seen only in frame 0 and one more frame lose their track (fewer than 2 observations left)
    assert all(len(tr.xy) >= 2 for tr in sim.tracks.values())


def test_regauge_uses_full_rotation_near_pitch_singularity(pkg):
    """estimator.cpp:541-546: within 1 degree of +-90 deg pitch the yaw difference is meaningless; the reference
    falls back to rot_diff = Rs[0] * R00^T."""
    sl, S = pkg.slider, pkg.synth
    R0 = S.euler_zyx(0.4, np.deg2rad(89.7), 0.1)
    before = np.concatenate([[1.0, 2.0, 3.0], S.rot_to_quat(R0)])
    Rg = S.euler_zyx(0.3, 0.02, -0.01)                 # an arbitrary rotation the solver drifted by
    pose = np.zeros((2, 7))
    pose[0, :3], pose[0, 3:] = [0.5, 0.5, 0.5], S.rot_to_quat(Rg @ R0)
    pose[1, :3], pose[1, 3:] = [1.5, 0.5, 0.5], S.rot_to_quat(Rg @ R0 @ S.euler_zyx(0.1, 0, 0))
    sb = np.zeros((2, 9))
    sl.regauge(before, pose, sb)
    assert np.allclose(S.quat_to_rot(pose[0, 3:]), R0, atol=1e-9) and np.allclose(pose[0, :3], [1, 2, 3])
    assert np.allclose(pose[1, :3], np.array([1, 2, 3]) + Rg.T @ np.array([1.0, 0, 0]), atol=1e-9)
