"""CPU test of FrontEnd.hough_lines (the ctypes wrapper of lsf_hough_batch) against a stand-in library function that fills the
lsf_segments buffers the way the C entry does: argument order, buffer shapes, the grow-and-retry on LSF_E_CAPACITY, the SegmentBatch
that comes back.  (The C entry itself needs a GPU: tests/test_gpu_parity.py::test_hough_detector_isolated.)"""
import ctypes as C

import numpy as np

import lane_slam_b200 as L
from lane_slam_b200 import _lib as lib_mod, frontend


class FakeLib(object):
    def __init__(self, per_task):
        self.per_task, self.calls = per_task, []

    def lsf_hough_batch(self, ctx, th, ml, mg, ground, seg_ref):
        seg = seg_ref._obj
        n = 2
        S = sum(self.per_task)
        self.calls.append((th, ml, mg, ground, seg.capacity))
        seg.n_frames, seg.n_segments = n, S
        if S > seg.capacity:
            return lib_mod.LSF_E_CAPACITY

        def arr(ptr, shape, dtype):
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), shape=shape)
        arr(seg.counts, (n, 3), np.int32)[:] = np.array(self.per_task, np.int32).reshape(n, 3)
        arr(seg.frame_offset, (n + 1,), np.int32)[:] = [0, sum(self.per_task[:3]), S]
        arr(seg.color, (S,), np.uint8)[:] = np.repeat(np.tile([0, 1, 2], n), self.per_task)
        arr(seg.lines_px, (S, 4), np.float32)[:] = np.arange(S * 4, dtype=np.float32).reshape(S, 4)
        arr(seg.normals, (S, 2), np.float64)[:] = 0.5
        arr(seg.centers, (S, 2), np.float32)[:] = 1.5
        arr(seg.pixels_normalized, (S, 4), np.float32)[:] = 0.25
        arr(seg.normal_f32, (S, 2), np.float32)[:] = 0.5
        if ground:
            arr(seg.ground, (S, 4), np.float64)[:] = 2.0
            arr(seg.keep, (S,), np.uint8)[:] = 1
        return 0

    def lsf_last_error(self, ctx):
        return b"fake"


def _front_end(per_task):
    fe = object.__new__(frontend.FrontEnd)
    fe._lib, fe._ctx, fe._last_n = FakeLib(per_task), None, 2
    return fe


def test_hough_lines_wrapper_shapes_and_values():
    fe = _front_end([3, 0, 2, 1, 4, 0])
    b = fe.hough_lines(20, 3, 1, ground=True)
    assert fe._lib.calls == [(20, 3, 1, 1, 1024)]
    assert b.n_frames == 2 and b.n_segments == 10 and b.counts.tolist() == [[3, 0, 2], [1, 4, 0]] and b.frame_offset.tolist() == [0, 5, 10]
    f1 = b.frame(1)
    assert f1["color"].tolist() == [0, 1, 1, 1, 1] and f1["lines_px"].shape == (5, 4) and f1["lines_px"][0, 0] == 20.0
    assert f1["ground"].shape == (5, 4) and f1["keep"].all() and f1["desc"].shape == (5, 32) and b.match_idx is None


def test_hough_lines_wrapper_grows_and_retries():
    fe = _front_end([700, 0, 0, 0, 800, 0])          # 1500 rows > the starting capacity max(1024, 512 * n)
    b = fe.hough_lines(2, 3, 1, ground=False)
    caps = [c[4] for c in fe._lib.calls]
    assert len(caps) == 2 and caps[0] == 1024 and caps[1] >= 1500
    assert b.n_segments == 1500 and b.lines_px.shape == (1500, 4) and not b.keep.any()
