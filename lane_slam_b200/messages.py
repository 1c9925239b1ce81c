"""Plain-Python mirrors of the five ROS messages on the path (src/duckietown_msgs/msg/Segment.msg:1-8,
SegmentList.msg:1-2, Vector2D.msg:1-2; geometry_msgs/Point) so a SegmentBatch can be handed to code
written against the reference's message fields without ROS."""
import numpy as np


class Vector2D(object):
    __slots__ = ("x", "y")

    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = np.float32(x), np.float32(y)   # float32 on the wire


class Point(object):
    __slots__ = ("x", "y", "z")

    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)


class Segment(object):
    WHITE, YELLOW, RED = 0, 1, 2
    __slots__ = ("color", "pixels_normalized", "normal", "points")

    def __init__(self):
        self.color = 0
        self.pixels_normalized = [Vector2D(), Vector2D()]
        self.normal = Vector2D()
        self.points = [Point(), Point()]


class SegmentList(object):
    __slots__ = ("header", "segments")

    def __init__(self, header=None):
        self.header = header
        self.segments = []


def segment_lists_from_batch(batch, stage="detector", headers=None):
    """SegmentBatch -> one SegmentList per frame, as published by
    stage="detector": line_detector_node (color, pixels_normalized, normal)        line_detector_node.py:251-265
    stage="ground":   ground_projection_node (color, points only)                   ground_projection_node.py:59-64
    stage="sanity":   line_sanity_node (ground segments that pass the filter)       line_sanity_node.py:52-70"""
    out = []
    for f in range(batch.n_frames):
        sl = SegmentList(headers[f] if headers else None)
        s = batch.frame_slice(f)
        for i in range(s.start, s.stop):
            if stage == "sanity" and not batch.keep[i]:
                continue
            seg = Segment()
            seg.color = int(batch.color[i])
            if stage == "detector":
                p = batch.pixels_normalized[i]
                seg.pixels_normalized = [Vector2D(p[0], p[1]), Vector2D(p[2], p[3])]
                seg.normal = Vector2D(batch.normal_f32[i][0], batch.normal_f32[i][1])
            else:
                g = batch.ground[i]
                seg.points = [Point(g[0], g[1], 0.0), Point(g[2], g[3], 0.0)]
            sl.segments.append(seg)
        out.append(sl)
    return out
