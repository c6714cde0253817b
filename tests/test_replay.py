"""Row f4: ROS-free replay of the estimator node's two input topics (wire codec, image_t conversion, measurement
pairing, per-frame IMU segments) -- host logic, CPU only."""
import numpy as np
import pytest


def _imu(rp, k, t0=100.0, rate=200.0, rng=None):
    t = t0 + k / rate
    sec = int(t)
    return rp.ImuMsg(k, sec, int(round((t - sec) * 1e9)), "imu", np.array([0, 0, 0, 1.0]), rng.normal(0, 0.1, 3),
                     rng.normal(0, 1, 3) + [0, 0, 9.8])


def _feat(rp, seq, t, rng, n=30):
    sec = int(t)
    ids = np.sort(rng.choice(5000, n, replace=False))
    xy = rng.uniform(-0.6, 0.6, (n, 2))
    return rp.pointcloud_from_features(seq, sec, int(round((t - sec) * 1e9)), ids, xy, rng.uniform(0, 480, (n, 2)),
                                       rng.normal(0, 0.2, (n, 2)), rng.uniform(0.05, 1, n)), ids, xy


def test_wire_round_trip_and_image_conversion(pkg):
    rp = pkg.replay
    rng = np.random.default_rng(0)
    msg, ids, xy = _feat(rp, 7, 100.25, rng)
    raw = rp.encode_pointcloud(msg)
    # header: seq, sec, nsec, len("world"); then the point count
    assert raw[:12] == np.array([7, 100, 250000000], "<u4").tobytes() and raw[12:21] == b"\x05\x00\x00\x00world"
    assert np.frombuffer(raw[21:25], "<u4")[0] == 30 and len(raw) == 25 + 30 * 12 + 4 + 6 * (4 + 4 + 30 * 4)
    back = rp.decode_pointcloud(raw)
    assert back.seq == 7 and back.frame_id == "world" and np.array_equal(back.points, msg.points)
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(back.channels, msg.channels))
    image = rp.image_from_pointcloud(back)
    assert sorted(image) == ids.tolist()
    for fid, x in zip(ids, xy):
        (cam, v), = image[int(fid)]
        assert cam == 0 and v[2] == 1.0 and v[0] == float(np.float32(x[0])) and v[1] == float(np.float32(x[1]))
        assert 0.05 <= v[rp.CHANNELS.index("prob") + 2] <= 1.0
    with pytest.raises(ValueError):
        rp.decode_pointcloud(raw[:-3])
    with pytest.raises(ValueError):
        rp.decode_pointcloud(raw + b"\x00")
    bad = rp.PointCloudMsg(0, 0, 0, "w", np.array([[0.1, 0.2, 0.5]], np.float32), msg.channels)
    with pytest.raises(AssertionError):
        rp.image_from_pointcloud(bad)
    # ids survive the float32 channel up to 2^24 (the tracker's ids are small consecutive integers)
    big = rp.pointcloud_from_features(0, 0, 0, [16777215], [[0, 0]], [[0, 0]], [[0, 0]], [1.0])
    assert list(rp.image_from_pointcloud(big)) == [16777215]
    imu = _imu(rp, 3, rng=rng)
    b2 = rp.decode_imu(rp.encode_imu(imu))
    assert len(rp.encode_imu(imu)) == 12 + 4 + 3 + 8 * (4 + 9 + 3 + 9 + 3 + 9)
    assert np.array_equal(b2.linear_acceleration, imu.linear_acceleration) and b2.stamp == imu.stamp


def test_measurement_pairing_and_imu_segments(pkg, tmp_path):
    """200 Hz IMU, 10 Hz features offset by 1.3 ms from the IMU grid: every frame gets the IMU samples up to its stamp
    plus one interpolated sample at the image time; the segments tile the timeline (sum dt == frame period) and the
    sample after each image is used twice."""
    rp = pkg.replay
    rng = np.random.default_rng(1)
    imus = [_imu(rp, k, rng=rng) for k in range(130)]                 # 100.000 .. 100.645
    feats = [_feat(rp, i, 100.0513 + 0.1 * i, rng)[0] for i in range(7)]   # the last one is beyond the IMU data
    early = _feat(rp, 99, 99.9, rng)[0]                               # older than every IMU message: thrown away
    path = str(tmp_path / "dump.bin")
    recs = [("feature", rp.encode_pointcloud(early))]
    fi = 0
    for m in imus:                                                    # arrival order: by stamp
        while fi < len(feats) and feats[fi].stamp <= m.stamp:
            recs.append(("feature", rp.encode_pointcloud(feats[fi])))
            fi += 1
        recs.append(("imu", rp.encode_imu(m)))
    recs += [("feature", rp.encode_pointcloud(f)) for f in feats[fi:]]
    rp.write_dump(path, recs)
    imu_buf, feat_buf, clock, segs = [], [], rp.ImuClock(), []
    for topic, m in rp.read_dump(path):                               # node main loop: callbacks fill the buffers
        (imu_buf if topic == "imu" else feat_buf).append(m)
        for ms, img in rp.get_measurements(imu_buf, feat_buf):
            dt, acc, gyr = rp.imu_segment(clock, ms, img.stamp)
            segs.append((img, ms, dt, acc, gyr))
    assert [s[0].seq for s in segs] == [0, 1, 2, 3, 4, 5]             # `early` thrown, frame 6 still waiting for IMU
    assert len(feat_buf) == 1 and feat_buf[0].seq == 6
    for k, (img, ms, dt, acc, gyr) in enumerate(segs):
        assert ms[-1].stamp > img.stamp >= ms[-2].stamp and len(dt) == len(ms)
        if k:
            assert abs(dt.sum() - 0.1) < 1e-9 and ms[0] is segs[k - 1][1][-1]     # reused sample
            assert abs(dt[0] - (ms[0].stamp - segs[k - 1][0].stamp)) < 1e-12
        # last sample: linear interpolation at the image time
        t0, t1 = ms[-2].stamp, ms[-1].stamp
        w2 = (img.stamp - t0) / (t1 - t0)
        assert np.allclose(acc[-1], (1 - w2) * ms[-2].linear_acceleration + w2 * ms[-1].linear_acceleration, atol=1e-9)
        assert np.allclose(gyr[-1], (1 - w2) * ms[-2].angular_velocity + w2 * ms[-1].angular_velocity, atol=1e-9)
        assert np.array_equal(acc[-2], ms[-2].linear_acceleration)
    assert segs[0][2][0] == 0.0                                       # first ever sample: current_time initialised to it


def test_replayed_segments_feed_the_preintegration_oracle(pkg, oracle):
    """The (dt, acc, gyr) arrays of a replayed frame are the inputs of IntegrationBase::push_back: integrate them with the
    oracle and with the numpy class, starting from the reference's acc_0 / gyr_0 carry-over (estimator.cpp:88-93, 117-118)."""
    import ctypes as C
    rp, abi, S = pkg.replay, pkg.abi, pkg.synth
    rng = np.random.default_rng(2)
    imus = [_imu(rp, k, rng=rng) for k in range(60)]
    feats = [_feat(rp, i, 100.0513 + 0.1 * i, rng)[0] for i in range(2)]
    imu_buf, feat_buf, clock = list(imus), list(feats), rp.ImuClock()
    (ms0, img0), (ms1, img1) = rp.get_measurements(imu_buf, feat_buf)
    _, acc0, gyr0 = rp.imu_segment(clock, ms0, img0.stamp)
    dt, acc, gyr = rp.imu_segment(clock, ms1, img1.stamp)
    ba, bg = np.zeros(3), np.zeros(3)
    pre = S.Preintegration(acc0[-1], gyr0[-1], ba, bg)                # acc_0 / gyr_0 = last processIMU sample of frame 0
    c = abi.Preint()
    c.delta_q[3] = 1.0
    for i in range(15):
        c.jacobian[i * 15 + i] = 1.0
    pa, pg = acc0[-1].copy(), gyr0[-1].copy()
    for k in range(len(dt)):
        pre.push_back(dt[k], acc[k], gyr[k])
        oracle.oracle_preint_propagate(C.byref(c), float(dt[k]), abi.dptr(pa), abi.dptr(pg), abi.dptr(acc[k].copy()),
                                       abi.dptr(gyr[k].copy()), S.ACC_N, S.GYR_N, S.ACC_W, S.GYR_W)
        pa, pg = acc[k].copy(), gyr[k].copy()
    got = np.frombuffer(bytes(c), dtype=np.float64)
    assert abs(got[16] - 0.1) < 1e-9 and np.allclose(got[:17], S.pack_preint(pre)[:17], rtol=1e-12, atol=1e-14)


def _record_session(pkg, path, seed, frames=26, imu_rate=200, frame_dt=0.1):
    """An independent front-end stand-in: 200 Hz IMU messages and, per camera frame, one feature message with every
    visible landmark (tracker-style ids: consecutive integers in order of first detection).  Written as a dump file."""
    S, sl, rp = pkg.synth, pkg.slider, pkg.replay
    rng = np.random.default_rng(seed)
    traj = S.Trajectory(phase=rng.uniform(0, 5))
    world = sl.World(rng, n=9000)
    U_, _, Vt_ = np.linalg.svd(S.EUROC_RIC)
    ric, tic = U_ @ Vt_, S.EUROC_TIC.copy()
    ba, bg, g = rng.normal(0, 0.02, 3), rng.normal(0, 0.002, 3), np.array([0, 0, S.G_NORM])
    t_ros = 1000.0                                              # ROS time of trajectory time 1.0
    n_per = int(round(frame_dt * imu_rate))
    recs, tracker_id, next_id, init = [], {}, 1, {}
    prev_xy = {}

    def stamp(t):
        sec = int(t)
        return sec, int(round((t - sec) * 1e9))
    k_imu = 0
    for f in range(frames):
        tf = 1.0 + f * frame_dt + 0.0013                        # camera 1.3 ms off the IMU grid
        while 1.0 + k_imu / imu_rate <= tf + 1.5 / imu_rate:   # IMU messages up to just past the image
            t = 1.0 + k_imu / imu_rate
            R = traj.rot(t)
            acc = R.T @ (traj.acc(t) + g) + ba + rng.normal(0, S.ACC_N, 3)
            gyr = traj.omega_body(t) + bg + rng.normal(0, S.GYR_N, 3)
            recs.append(("imu", rp.encode_imu(rp.ImuMsg(k_imu, *stamp(t_ros + t - 1.0), "imu", np.array([0, 0, 0, 1.0]), gyr, acc))))
            k_imu += 1
        R, p = traj.rot(tf), traj.pos(tf)
        ids, xy, _ = world.observe(S.EUROC_CAM, R @ ric, p + R @ tic)
        fid = []
        for i in ids.tolist():
            if i not in tracker_id:
                tracker_id[i] = next_id
                next_id += 1
            fid.append(tracker_id[i])
        order = np.argsort(fid)
        fid, xy, lid = np.array(fid)[order], xy[order] + rng.normal(0, 0.5 / 460.0, (len(order), 2)), ids[order]
        vel = np.array([(xy[j] - prev_xy[l]) / frame_dt if l in prev_xy else (0.0, 0.0) for j, l in enumerate(lid.tolist())]).reshape(-1, 2)
        prev_xy = {l: xy[j] for j, l in enumerate(lid.tolist())}
        recs.append(("feature", rp.encode_pointcloud(rp.pointcloud_from_features(f, *stamp(t_ros + tf - 1.0), fid, xy, np.zeros((len(fid), 2)),
                                                                                 vel, world.score[lid]))))
        if f < 11:                                              # what the initializer would deliver
            init[f] = (np.concatenate([p + rng.normal(0, 0.02, 3), S.rot_to_quat(R)]),
                       np.concatenate([traj.vel(tf) + rng.normal(0, 0.05, 3), ba + rng.normal(0, 0.01, 3), bg + rng.normal(0, 0.001, 3)]))
    rp.write_dump(path, recs)
    gt = lambda t: (np.concatenate([traj.pos(t - t_ros + 1.0), S.rot_to_quat(traj.rot(t - t_ros + 1.0))]), np.zeros(9))
    return dict(ric=ric, tic=tic, init=init, gt=gt, n_frames=frames, n_ids=next_id - 1)


def test_recorded_session_replays_through_the_backend(pkg, oracle, tmp_path):
    """Dump -> decode -> measurement pairing -> window bookkeeping -> optimize / marginalize / select (oracle backend
    here; slider.GpuBackend in the product): the estimate follows the recorded trajectory, the budget holds, only ids
    newer than everything seen before are ever selected."""
    from slider_backends import OracleBackend
    sl, rp, S = pkg.slider, pkg.replay, pkg.synth
    path = str(tmp_path / "session.bvio")
    rec = _record_session(pkg, path, seed=4)
    ses = sl.ReplaySession(S.EUROC_CAM, rec["ric"], rec["tic"], rec["init"], max_feats=70, H=10, opts=dict(max_iters=8),
                           keyframes="parallax", gt=rec["gt"])
    be = OracleBackend(oracle, pkg.abi)
    lats, n_new = [], 0
    for topic, msg in rp.read_dump(path):
        before, known = ses.last_feature_id, set(ses.tracks)
        for lat in ses.feed(topic, msg, be):
            if lat is not None:
                lats.append(lat)
            started = set(ses.tracks) - known
            assert all(i > before for i in started), (before, started)       # only never-seen-before ids are selectable
            n_new += len(started)
            before, known = ses.last_feature_id, set(ses.tracks)
    assert n_new > 70
    assert ses.frame == rec["n_frames"] - 1 or ses.frame == rec["n_frames"]      # the last image may still wait for IMU
    assert len(lats) >= ses.frame - 10 - 1 and all(l["iterations"] >= 1 and l["L"] >= 15 for l in lats[2:])
    errs = np.array([h[1] for h in ses.history])
    assert errs[-5:].max() < 0.3, errs
    assert ses.prior is not None and ses.prior["n"] in (69, 75)
    assert sum(1 for tr in ses.tracks.values() if tr.alive) <= 70
    assert abs(ses.frame_dt - 0.1) < 1e-6 and ses.n_imu in (20, 21, 22)
