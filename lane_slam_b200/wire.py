"""Wire formats either side of the path, without ROS: ROS 1 message (de)serialisation of the messages
the line path consumes and produces, and a minimal rosbag 2.0 reader / writer, so that a recorded log can be replayed through
``FrontEnd.process_jpeg`` and its results stored the way the reference's nodes publish them.

Message layouts (ROS 1 serialisation: little endian, no padding, ``string`` / ``T[]`` = uint32 length + payload,
``time`` = uint32 secs + uint32 nsecs) follow the reference's definitions:
  sensor_msgs/CompressedImage   Header header, string format, uint8[] data           (consumed: line_detector_node.py:151-158)
  duckietown_msgs/SegmentList   Header header, Segment[] segments                     src/duckietown_msgs/msg/SegmentList.msg:1-2
  duckietown_msgs/Segment       uint8 color, Vector2D[2] pixels_normalized, Vector2D normal, geometry_msgs/Point[2] points
                                                                                      src/duckietown_msgs/msg/Segment.msg:1-8
  duckietown_msgs/Vector2D      float32 x, float32 y                                  src/duckietown_msgs/msg/Vector2D.msg:1-2
  duckietown_msgs/WheelsCmdStamped  Header header, float32 vel_left, float32 vel_right   (odometry.py:108-120)
  duckietown_msgs/LanePose      Header header, float32 d, sigma_d, phi, sigma_phi, int32 status, bool in_lane
  duckietown_msgs/AntiInstagramTransform  Header header, float64[6] s                 (line_detector_node.py:112-114)
"""
import bz2
import struct
from collections import namedtuple

import numpy as np

Header = namedtuple("Header", ["seq", "secs", "nsecs", "frame_id"])
CompressedImage = namedtuple("CompressedImage", ["header", "format", "data"])
WheelsCmdStamped = namedtuple("WheelsCmdStamped", ["header", "vel_left", "vel_right"])
LanePose = namedtuple("LanePose", ["header", "d", "sigma_d", "phi", "sigma_phi", "status", "in_lane"])
AntiInstagramTransform = namedtuple("AntiInstagramTransform", ["header", "s"])

# one Segment on the wire: 73 bytes, packed
SEGMENT_DTYPE = np.dtype([("color", "u1"), ("pixels_normalized", "<f4", (2, 2)), ("normal", "<f4", (2,)), ("points", "<f8", (2, 3))])
assert SEGMENT_DTYPE.itemsize == 73

TYPES = {
    "sensor_msgs/CompressedImage": "8f7a12909da2c9d3332d540a0977563f",
    "duckietown_msgs/SegmentList": "",          # md5 sums of the duckietown messages depend on the workspace; readers here ignore them
    "duckietown_msgs/WheelsCmdStamped": "",
    "duckietown_msgs/LanePose": "",
    "duckietown_msgs/AntiInstagramTransform": "",
}


# ---- primitives -----------------------------------------------------------------------------------------------------------
def _pack_string(s):
    b = s.encode() if isinstance(s, str) else bytes(s)
    return struct.pack("<I", len(b)) + b


def pack_header(h):
    h = h if h is not None else Header(0, 0, 0, "")
    return struct.pack("<III", h.seq, h.secs, h.nsecs) + _pack_string(h.frame_id)


def unpack_header(buf, o=0):
    seq, secs, nsecs, n = struct.unpack_from("<IIII", buf, o)
    o += 16
    return Header(seq, secs, nsecs, bytes(buf[o:o + n]).decode()), o + n


# ---- messages ---------------------------------------------------------------------------------------------------------------
def serialize_compressed_image(m):
    data = np.asarray(m.data, np.uint8).tobytes() if not isinstance(m.data, (bytes, bytearray)) else bytes(m.data)
    return pack_header(m.header) + _pack_string(m.format) + struct.pack("<I", len(data)) + data


def deserialize_compressed_image(buf):
    h, o = unpack_header(buf)
    n, = struct.unpack_from("<I", buf, o); o += 4
    fmt = bytes(buf[o:o + n]).decode(); o += n
    n, = struct.unpack_from("<I", buf, o); o += 4
    return CompressedImage(h, fmt, np.frombuffer(buf, np.uint8, n, o))


def serialize_wheels_cmd(m):
    return pack_header(m.header) + struct.pack("<ff", m.vel_left, m.vel_right)


def deserialize_wheels_cmd(buf):
    h, o = unpack_header(buf)
    vl, vr = struct.unpack_from("<ff", buf, o)
    return WheelsCmdStamped(h, vl, vr)


def serialize_lane_pose(m):
    return pack_header(m.header) + struct.pack("<ffffiB", m.d, m.sigma_d, m.phi, m.sigma_phi, m.status, 1 if m.in_lane else 0)


def deserialize_lane_pose(buf):
    h, o = unpack_header(buf)
    d, sd, phi, sp, status, il = struct.unpack_from("<ffffiB", buf, o)
    return LanePose(h, d, sd, phi, sp, status, bool(il))


def serialize_anti_instagram_transform(m):
    return pack_header(m.header) + np.asarray(m.s, "<f8").reshape(6).tobytes()


def deserialize_anti_instagram_transform(buf):
    h, o = unpack_header(buf)
    return AntiInstagramTransform(h, np.frombuffer(buf, "<f8", 6, o).copy())


def segments_array(batch, f, stage="detector"):
    """Frame f of a SegmentBatch as a structured array of wire Segments, as the three nodes publish them:
    stage="detector": line_detector_node -- color, pixels_normalized, normal (points zero)         line_detector_node.py:251-265
    stage="ground":   ground_projection_node -- color, points (z = 0)                              ground_projection_node.py:59-64
    stage="sanity":   line_sanity_node -- the ground segments that pass the filter                  line_sanity_node.py:52-70"""
    s = batch.frame_slice(f)
    if stage == "sanity":
        idx = np.nonzero(batch.keep[s])[0] + s.start
    else:
        idx = np.arange(s.start, s.stop)
    a = np.zeros(len(idx), SEGMENT_DTYPE)
    a["color"] = batch.color[idx]
    if stage == "detector":
        a["pixels_normalized"] = batch.pixels_normalized[idx].reshape(-1, 2, 2)
        a["normal"] = batch.normal_f32[idx]
    else:
        a["points"][:, :, :2] = batch.ground[idx].reshape(-1, 2, 2)
    return a


def serialize_segment_list(header, segments):
    """segments: structured array (SEGMENT_DTYPE) -> duckietown_msgs/SegmentList bytes."""
    seg = np.ascontiguousarray(segments, SEGMENT_DTYPE)
    return pack_header(header) + struct.pack("<I", len(seg)) + seg.tobytes()


def deserialize_segment_list(buf):
    h, o = unpack_header(buf)
    n, = struct.unpack_from("<I", buf, o)
    return h, np.frombuffer(buf, SEGMENT_DTYPE, n, o + 4).copy()


def serialize_batch(batch, stage="detector", headers=None):
    """One SegmentList message per frame of a SegmentBatch."""
    return [serialize_segment_list(headers[f] if headers else None, segments_array(batch, f, stage)) for f in range(batch.n_frames)]


def jpeg_blob(messages):
    """CompressedImage messages -> (blob uint8, offsets int64 [n+1]) for FrontEnd.process_jpeg."""
    datas = [np.asarray(m.data, np.uint8) for m in messages]
    off = np.concatenate([[0], np.cumsum([len(d) for d in datas])]).astype(np.int64)
    return (np.concatenate(datas) if datas else np.zeros(0, np.uint8)), off


# ---- rosbag 2.0 -----------------------------------------------------------------------------------------------------------
_MAGIC = b"#ROSBAG V2.0\n"
_OP_MSG, _OP_BAG, _OP_INDEX, _OP_CHUNK, _OP_CHUNK_INFO, _OP_CONN = 0x02, 0x03, 0x04, 0x05, 0x06, 0x07


def _fields(buf):
    out, o = {}, 0
    while o < len(buf):
        n, = struct.unpack_from("<I", buf, o); o += 4
        k, _, v = bytes(buf[o:o + n]).partition(b"=")
        out[k.decode()] = v
        o += n
    return out


def _pack_fields(d):
    out = b""
    for k, v in d.items():
        f = k.encode() + b"=" + v
        out += struct.pack("<I", len(f)) + f
    return out


def _record(fields, data):
    h = _pack_fields(fields)
    return struct.pack("<I", len(h)) + h + struct.pack("<I", len(data)) + data


def _iter_records(buf, o=0, end=None):
    end = len(buf) if end is None else end
    while o + 8 <= end:
        hl, = struct.unpack_from("<I", buf, o)
        h = _fields(buf[o + 4:o + 4 + hl])
        o += 4 + hl
        dl, = struct.unpack_from("<I", buf, o)
        yield h, buf[o + 4:o + 4 + dl]
        o += 4 + dl


def read_bag(path, topics=None):
    """Yield (topic, datatype, (secs, nsecs), message bytes) in file order.  Chunks may be uncompressed or bz2."""
    buf = memoryview(open(path, "rb").read())
    if bytes(buf[:len(_MAGIC)]) != _MAGIC:
        raise ValueError("not a rosbag 2.0 file")
    conns = {}

    def handle(h, data):
        op = h["op"][0]
        if op == _OP_CONN:
            c = _fields(data)
            conns[struct.unpack("<I", h["conn"])[0]] = (h["topic"].decode(), c.get("type", b"").decode())
        elif op == _OP_MSG:
            topic, typ = conns.get(struct.unpack("<I", h["conn"])[0], ("?", "?"))
            if topics is None or topic in topics:
                secs, nsecs = struct.unpack("<II", h["time"])
                return topic, typ, (secs, nsecs), bytes(data)
        return None

    for h, data in _iter_records(buf, len(_MAGIC)):
        op = h["op"][0]
        if op == _OP_CHUNK:
            comp = h.get("compression", b"none")
            if comp == b"bz2":
                data = memoryview(bz2.decompress(bytes(data)))
            elif comp != b"none":
                raise ValueError("rosbag chunk compression %r is not supported (none / bz2)" % comp)
            for h2, d2 in _iter_records(data):
                r = handle(h2, d2)
                if r:
                    yield r
        else:
            r = handle(h, data)
            if r:
                yield r


def write_bag(path, messages, compression="none", chunk_messages=256):
    """messages: iterable of (topic, datatype, (secs, nsecs), bytes).  Writes chunks, connection records, per-chunk index
    records and chunk infos, and the bag header (index_pos, conn_count, chunk_count)."""
    conn_ids, conn_recs = {}, {}
    chunks = []          # (position, start, end, {conn: count})
    out = bytearray(_MAGIC)
    out += b"\0" * 4096  # bag header record, rewritten at the end
    pending, index = [], {}

    def conn_record(cid, topic, typ):
        data = _pack_fields({"topic": topic.encode(), "type": typ.encode(), "md5sum": TYPES.get(typ, "").encode(), "message_definition": b""})
        return _record({"op": bytes([_OP_CONN]), "conn": struct.pack("<I", cid), "topic": topic.encode()}, data)

    def flush():
        if not pending:
            return
        body = bytearray()
        idx = {}
        times = []
        for cid, t, rec, is_msg in pending:
            if is_msg:
                idx.setdefault(cid, []).append((t, len(body)))
                times.append(t)
            body += rec
        raw = bytes(body)
        data = bz2.compress(raw) if compression == "bz2" else raw
        pos = len(out)
        out.extend(_record({"op": bytes([_OP_CHUNK]), "compression": compression.encode(), "size": struct.pack("<I", len(raw))}, data))
        for cid, ents in idx.items():
            d = b"".join(struct.pack("<III", t[0], t[1], off) for t, off in ents)
            out.extend(_record({"op": bytes([_OP_INDEX]), "ver": struct.pack("<I", 1), "conn": struct.pack("<I", cid),
                                "count": struct.pack("<I", len(ents))}, d))
        chunks.append((pos, min(times), max(times), {c: len(e) for c, e in idx.items()}))
        pending.clear()

    for topic, typ, t, data in messages:
        if topic not in conn_ids:
            cid = len(conn_ids)
            conn_ids[topic] = cid
            conn_recs[cid] = conn_record(cid, topic, typ)
            pending.append((cid, t, conn_recs[cid], False))
        cid = conn_ids[topic]
        rec = _record({"op": bytes([_OP_MSG]), "conn": struct.pack("<I", cid), "time": struct.pack("<II", t[0], t[1])}, data)
        pending.append((cid, t, rec, True))
        if len(pending) >= chunk_messages:
            flush()
    flush()
    index_pos = len(out)
    for cid in sorted(conn_recs):
        out.extend(conn_recs[cid])
    for pos, t0, t1, counts in chunks:
        d = b"".join(struct.pack("<II", c, n) for c, n in counts.items())
        out.extend(_record({"op": bytes([_OP_CHUNK_INFO]), "ver": struct.pack("<I", 1), "chunk_pos": struct.pack("<Q", pos),
                            "start_time": struct.pack("<II", *t0), "end_time": struct.pack("<II", *t1),
                            "count": struct.pack("<I", len(counts))}, d))
    h = _pack_fields({"op": bytes([_OP_BAG]), "index_pos": struct.pack("<Q", index_pos), "conn_count": struct.pack("<I", len(conn_recs)),
                      "chunk_count": struct.pack("<I", len(chunks))})
    head = struct.pack("<I", len(h)) + h
    pad = 4096 - len(head) - 4
    out[len(_MAGIC):len(_MAGIC) + 4096] = head + struct.pack("<I", pad) + b" " * pad
    open(path, "wb").write(bytes(out))
