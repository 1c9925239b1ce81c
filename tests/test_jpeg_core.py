"""CPU test of the GPU JPEG decoder's algorithm: lane_slam_b200/csrc/jpeg_core.cuh (the arithmetic the CUDA kernels run) is
compiled for the host and driven by oracle/csrc/jpeg_parallel_check.cpp exactly the way k_jpeg.cu orchestrates it --
stuffing removal, self-synchronising parallel Huffman decode simulated thread by thread, prefix sums, DC prediction, islow IDCT,
fancy upsampling, colour conversion -- and must reproduce cv2.imdecode bit for bit."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest

import realset
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("jpc") / "libjpc.so")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "oracle", "csrc", "jpeg_parallel_check.cpp")])
    lib = C.CDLL(so)

    def decode(data, shape, sub_bytes):
        b = np.ascontiguousarray(data, np.uint8)
        out = np.zeros(shape, np.uint8)
        r, n = C.c_int(), C.c_int()
        rc = lib.jpc_decode(b.ctypes.data_as(C.c_void_p), C.c_size_t(len(b)), out.ctypes.data_as(C.c_void_p), int(sub_bytes), C.byref(r), C.byref(n))
        return rc, out, r.value, n.value
    return decode


def test_parallel_decode_equals_cv2_on_real_frames(sim):
    for i in range(0, realset.count(), 2):
        img = realset.image(i)
        for sub in (48, 152):
            rc, out, rounds, nsub = sim(realset.jpeg(i), img.shape, sub)
            assert rc == 0 and np.array_equal(out, img), (i, sub)
            assert rounds < nsub          # the self-synchronisation converges long before the sequential worst case


def test_parallel_decode_samplings_sizes_qualities(sim):
    for (H, W) in [(480, 640), (123, 161), (97, 203)]:
        im = synth.frame(5, H, W)
        for q in (40, 95):
            for ss in (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
                       cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440):
                ok, enc = cv2.imencode('.jpg', im, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss])
                ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
                rc, out, rounds, nsub = sim(enc, ref.shape, 64)
                assert rc == 0 and np.array_equal(out, ref), (H, W, q, ss)
    ok, enc = cv2.imencode('.jpg', cv2.cvtColor(synth.frame(1), cv2.COLOR_BGR2GRAY))
    ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    rc, out, _, _ = sim(enc, ref.shape, 64)
    assert rc == 0 and np.array_equal(out, ref)
    ok, enc = cv2.imencode('.jpg', synth.frame(1), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert sim(enc, (480, 640, 3), 64)[0] == -2                 # outside the scope: reported, not mis-decoded
    ok, enc = cv2.imencode('.jpg', synth.frame(1), [cv2.IMWRITE_JPEG_RST_INTERVAL, 4])
    assert sim(enc, (480, 640, 3), 64)[0] == -2
