"""CPU tests of the wire formats (lane_slam_b200/wire.py): ROS 1 serialisation of the messages either side of the path and the
rosbag 2.0 reader / writer -- byte layouts written out by hand from the .msg definitions, round trips, bz2 chunks."""
import os
import struct

import numpy as np

import realset
from lane_slam_b200 import wire


def test_segment_layout_is_the_msg_definition():
    """duckietown_msgs/Segment: uint8 color, Vector2D[2] pixels_normalized (float32 x, y), Vector2D normal, Point[2] points (float64)."""
    a = np.zeros(1, wire.SEGMENT_DTYPE)
    a["color"] = 2
    a["pixels_normalized"] = [[[0.25, 0.5], [0.75, 1.0]]]
    a["normal"] = [[-1.0, 0.5]]
    a["points"] = [[[1.5, -2.5, 0.0], [3.5, 4.5, 0.0]]]
    h = wire.Header(7, 100, 200, "duck")
    msg = wire.serialize_segment_list(h, a)
    want = struct.pack("<III", 7, 100, 200) + struct.pack("<I", 4) + b"duck" + struct.pack("<I", 1) + \
        struct.pack("<B", 2) + struct.pack("<ffff", 0.25, 0.5, 0.75, 1.0) + struct.pack("<ff", -1.0, 0.5) + \
        struct.pack("<dddddd", 1.5, -2.5, 0.0, 3.5, 4.5, 0.0)
    assert msg == want and len(want) == 16 + 4 + 4 + 73
    h2, a2 = wire.deserialize_segment_list(msg)
    assert h2 == h and a2.tobytes() == a.tobytes()


def test_message_round_trips():
    h = wire.Header(1, 2, 3, "/camera")
    ci = wire.CompressedImage(h, "jpeg", np.asarray(realset.jpeg(0)))
    b = wire.serialize_compressed_image(ci)
    assert b[:12] == struct.pack("<III", 1, 2, 3) and b[12:16] == struct.pack("<I", 7) and b[16:23] == b"/camera"
    ci2 = wire.deserialize_compressed_image(b)
    assert ci2.header == h and ci2.format == "jpeg" and np.array_equal(ci2.data, ci.data)
    w = wire.deserialize_wheels_cmd(wire.serialize_wheels_cmd(wire.WheelsCmdStamped(h, 0.25, -0.5)))
    assert (w.vel_left, w.vel_right) == (0.25, -0.5)
    lp = wire.deserialize_lane_pose(wire.serialize_lane_pose(wire.LanePose(h, 0.125, 0.0, -0.5, 0.0, 0, True)))
    assert (lp.d, lp.phi, lp.status, lp.in_lane) == (0.125, -0.5, 0, True)
    ai = wire.deserialize_anti_instagram_transform(wire.serialize_anti_instagram_transform(wire.AntiInstagramTransform(h, [1, 2, 3, 4, 5, 6])))
    assert ai.s.tolist() == [1, 2, 3, 4, 5, 6]


def test_rosbag_write_read(tmp_path):
    msgs = []
    for i in range(6):
        h = wire.Header(i, 10 + i, 5000 * i, "cam")
        msgs.append(("/duck/camera_node/image/compressed", "sensor_msgs/CompressedImage", (10 + i, 5000 * i),
                     wire.serialize_compressed_image(wire.CompressedImage(h, "jpeg", np.asarray(realset.jpeg(i))))))
        msgs.append(("/duck/wheels_driver_node/wheels_cmd", "duckietown_msgs/WheelsCmdStamped", (10 + i, 5000 * i + 1),
                     wire.serialize_wheels_cmd(wire.WheelsCmdStamped(h, 0.3, 0.3 + 0.01 * i))))
    for comp in ("none", "bz2"):
        p = str(tmp_path / ("log_%s.bag" % comp))
        wire.write_bag(p, msgs, compression=comp, chunk_messages=5)
        raw = open(p, "rb").read()
        assert raw.startswith(b"#ROSBAG V2.0\n") and raw[13 + 4096:13 + 4096 + 4] != b""
        got = list(wire.read_bag(p))
        assert got == msgs
        imgs = [wire.deserialize_compressed_image(m[3]) for m in wire.read_bag(p, topics={"/duck/camera_node/image/compressed"})]
        blob, off = wire.jpeg_blob(imgs)
        assert len(off) == 7 and np.array_equal(blob[off[2]:off[3]], np.asarray(realset.jpeg(2)))
    # bag header: op 3 with index_pos / conn_count / chunk_count, padded to 4096 bytes
    hl, = struct.unpack_from("<I", raw, 13)
    f = wire._fields(raw[17:17 + hl])
    assert f["op"] == b"\x03" and struct.unpack("<I", f["conn_count"])[0] == 2 and struct.unpack("<I", f["chunk_count"])[0] >= 2
    assert struct.unpack("<Q", f["index_pos"])[0] > 4096
