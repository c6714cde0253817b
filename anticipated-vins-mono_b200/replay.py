"""ROS-free replay of recorded front-end traffic, host side.

Lets a dump of the two topics the estimator node subscribes to -- the feature tracker's `sensor_msgs/PointCloud` and the
`sensor_msgs/Imu` stream -- drive the C-ABI without ROS:

  * `decode_pointcloud` / `encode_pointcloud`, `decode_imu` / `encode_imu`: the ROS 1 wire serialization of the two
    message types (little-endian; `time` = uint32 sec + uint32 nsec; `string` and arrays are uint32-length prefixed),
    i.e. the payload bytes a rosbag record or a TCPROS frame carries;
  * `image_from_pointcloud`: the feature message -> `image_t` conversion of the estimator node
    (vins_estimator/src/estimator_node.cpp:303-321) for the layout the tracker publishes
    (feature_tracker/src/feature_tracker_ros.cpp:75-115): points = normalized-plane (x, y, 1) as float32, channels =
    id, pixel u, pixel v, velocity x, velocity y, probability; `feature_id = int(id + 0.5) / NUM_OF_CAM`;
  * `get_measurements`: the pairing of every feature message with the IMU messages up to its (td-shifted) stamp plus
    the first one after it, which is reused by the next frame (estimator_node.cpp:86-140);
  * `imu_segment`: the per-frame IMU loop of `process()` (estimator_node.cpp:232-275): dt from consecutive stamps, and
    the virtual sample linearly interpolated at the image time from the last two messages.  Its output is what
    `Estimator::processIMU` receives, i.e. the (dt, acc, gyr) arrays of `bvio_preintegrate`.

Pure Python / numpy: this is I/O glue, not arithmetic of the hot path.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

NUM_OF_CAM = 1            # vins_estimator/src/parameters.h:14
CHANNELS = ("id", "u", "v", "velocity_x", "velocity_y", "prob")


def stamp_to_sec(sec, nsec):
    """ros::Time::toSec()"""
    return float(sec) + 1e-9 * float(nsec)


@dataclasses.dataclass
class PointCloudMsg:
    seq: int
    sec: int
    nsec: int
    frame_id: str
    points: np.ndarray                 # [n,3] float32
    channels: list                     # [(name, float32 [n])]

    @property
    def stamp(self):
        return stamp_to_sec(self.sec, self.nsec)


@dataclasses.dataclass
class ImuMsg:
    seq: int
    sec: int
    nsec: int
    frame_id: str
    orientation: np.ndarray            # [4] x y z w
    angular_velocity: np.ndarray       # [3]
    linear_acceleration: np.ndarray    # [3]
    covariances: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(27))

    @property
    def stamp(self):
        return stamp_to_sec(self.sec, self.nsec)


def _str(b):
    return struct.pack("<I", len(b)) + b


class _Reader:
    def __init__(self, buf):
        self.b, self.o = memoryview(buf), 0

    def take(self, fmt):
        v = struct.unpack_from(fmt, self.b, self.o)
        self.o += struct.calcsize(fmt)
        return v

    def string(self):
        (n,) = self.take("<I")
        s = bytes(self.b[self.o:self.o + n])
        if len(s) != n:
            raise ValueError("truncated string")
        self.o += n
        return s.decode()

    def array(self, dtype, n):
        nbytes = np.dtype(dtype).itemsize * n
        if self.o + nbytes > len(self.b):
            raise ValueError("truncated array")
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.o).copy()
        self.o += nbytes
        return a


def encode_pointcloud(m: PointCloudMsg) -> bytes:
    out = [struct.pack("<III", m.seq, m.sec, m.nsec), _str(m.frame_id.encode())]
    pts = np.ascontiguousarray(m.points, dtype="<f4").reshape(-1, 3)
    out += [struct.pack("<I", len(pts)), pts.tobytes(), struct.pack("<I", len(m.channels))]
    for name, values in m.channels:
        v = np.ascontiguousarray(values, dtype="<f4")
        out += [_str(name.encode()), struct.pack("<I", len(v)), v.tobytes()]
    return b"".join(out)


def decode_pointcloud(buf) -> PointCloudMsg:
    r = _Reader(buf)
    seq, sec, nsec = r.take("<III")
    frame_id = r.string()
    (n,) = r.take("<I")
    pts = r.array("<f4", 3 * n).reshape(n, 3)
    (nc,) = r.take("<I")
    chans = []
    for _ in range(nc):
        name = r.string()
        (k,) = r.take("<I")
        chans.append((name, r.array("<f4", k)))
    if r.o != len(r.b):
        raise ValueError("trailing bytes after sensor_msgs/PointCloud")
    return PointCloudMsg(seq, sec, nsec, frame_id, pts, chans)


def encode_imu(m: ImuMsg) -> bytes:
    c = np.asarray(m.covariances, "<f8").reshape(27)
    return b"".join([struct.pack("<III", m.seq, m.sec, m.nsec), _str(m.frame_id.encode()),
                     np.asarray(m.orientation, "<f8").tobytes(), c[:9].tobytes(),
                     np.asarray(m.angular_velocity, "<f8").tobytes(), c[9:18].tobytes(),
                     np.asarray(m.linear_acceleration, "<f8").tobytes(), c[18:].tobytes()])


def decode_imu(buf) -> ImuMsg:
    r = _Reader(buf)
    seq, sec, nsec = r.take("<III")
    frame_id = r.string()
    q, c0 = r.array("<f8", 4), r.array("<f8", 9)
    w, c1 = r.array("<f8", 3), r.array("<f8", 9)
    a, c2 = r.array("<f8", 3), r.array("<f8", 9)
    if r.o != len(r.b):
        raise ValueError("trailing bytes after sensor_msgs/Imu")
    return ImuMsg(seq, sec, nsec, frame_id, q, w, a, np.concatenate([c0, c1, c2]))


def pointcloud_from_features(seq, sec, nsec, ids, xy, uv, vel, prob) -> PointCloudMsg:
    """What the tracker publishes (feature_tracker_ros.cpp:75-115); everything is narrowed to float32 on the wire."""
    n = len(ids)
    pts = np.column_stack([np.asarray(xy, np.float32).reshape(n, 2), np.ones(n, np.float32)])
    uv, vel = np.asarray(uv, np.float32).reshape(n, 2), np.asarray(vel, np.float32).reshape(n, 2)
    chans = [("", np.asarray(ids, np.float32)), ("", uv[:, 0]), ("", uv[:, 1]), ("", vel[:, 0]), ("", vel[:, 1]),
             ("", np.asarray(prob, np.float32))]
    return PointCloudMsg(seq, sec, nsec, "world", pts, chans)


def image_from_pointcloud(m: PointCloudMsg):
    """estimator_node.cpp:303-321 -> {feature_id: [(camera_id, [x, y, z, u, v, vx, vy, prob])]}, float32 widened to
    double; asserts z == 1 like the node."""
    if len(m.channels) < 6:
        raise ValueError("feature message needs 6 channels: " + ", ".join(CHANNELS))
    ch = [np.asarray(c[1], np.float32) for c in m.channels[:6]]
    image = {}
    for i in range(len(m.points)):
        v = int(float(ch[0][i]) + 0.5)                         # `int v = values[i] + 0.5`: float32 promoted to double
        fid, cam = v // NUM_OF_CAM, v % NUM_OF_CAM
        x, y, z = (float(t) for t in m.points[i])
        if z != 1.0:
            raise AssertionError("ROS_ASSERT(z == 1)")
        image.setdefault(fid, []).append((cam, np.array([x, y, z, float(ch[1][i]), float(ch[2][i]), float(ch[3][i]),
                                                         float(ch[4][i]), float(ch[5][i])])))
    return image


def get_measurements(imu_buf: list, feature_buf: list, td=0.0):
    """estimator_node.cpp:86-140.  Consumes from the two lists (oldest first) exactly like the node consumes its queues
    and returns [(imu messages, feature message)]; messages that cannot be paired yet stay in the lists."""
    out = []
    while True:
        if not imu_buf or not feature_buf:
            return out
        if not imu_buf[-1].stamp > feature_buf[0].stamp + td:
            return out                                       # wait for imu
        if not imu_buf[0].stamp < feature_buf[0].stamp + td:
            feature_buf.pop(0)                               # throw img
            continue
        img = feature_buf.pop(0)
        imus = []
        while imu_buf[0].stamp < img.stamp + td:
            imus.append(imu_buf.pop(0))
        imus.append(imu_buf[0])                              # used twice: interpolation now, integration next time
        out.append((imus, img))


class ImuClock:
    """`current_time` of the node's process() loop (estimator_node.cpp:37, 237-240)."""
    def __init__(self):
        self.current_time = -1.0


def imu_segment(clock: ImuClock, imus, img_stamp, td=0.0):
    """estimator_node.cpp:232-275 -> (dt [n], acc [n,3], gyr [n,3]) as handed to Estimator::processIMU."""
    dts, accs, gyrs = [], [], []
    a, g = np.zeros(3), np.zeros(3)
    img_t = img_stamp + td
    for m in imus:
        t = m.stamp
        if t <= img_t:
            if clock.current_time < 0:
                clock.current_time = t
            dt = t - clock.current_time
            assert dt >= 0
            clock.current_time = t
            a, g = np.array(m.linear_acceleration, float), np.array(m.angular_velocity, float)
            dts.append(dt)
        else:
            dt_1, dt_2 = img_t - clock.current_time, t - img_t
            clock.current_time = img_t
            assert dt_1 >= 0 and dt_2 >= 0 and dt_1 + dt_2 > 0
            w1, w2 = dt_2 / (dt_1 + dt_2), dt_1 / (dt_1 + dt_2)
            a = w1 * a + w2 * np.array(m.linear_acceleration, float)
            g = w1 * g + w2 * np.array(m.angular_velocity, float)
            dts.append(dt_1)
        accs.append(a.copy())
        gyrs.append(g.copy())
    return np.array(dts), np.array(accs).reshape(-1, 3), np.array(gyrs).reshape(-1, 3)


def write_dump(path, records):
    """records: [(topic 'imu' | 'feature', payload bytes)] in arrival order -> one length-prefixed file."""
    with open(path, "wb") as f:
        f.write(b"BVIODUMP1\n")
        for topic, payload in records:
            t = {"imu": 0, "feature": 1}[topic]
            f.write(struct.pack("<BI", t, len(payload)))
            f.write(payload)


def read_dump(path):
    out = []
    with open(path, "rb") as f:
        if f.readline() != b"BVIODUMP1\n":
            raise ValueError("not a BVIODUMP1 file")
        while True:
            h = f.read(5)
            if not h:
                return out
            t, n = struct.unpack("<BI", h)
            payload = f.read(n)
            if len(payload) != n:
                raise ValueError("truncated dump")
            out.append(("imu", decode_imu(payload)) if t == 0 else ("feature", decode_pointcloud(payload)))
