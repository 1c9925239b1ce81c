"""Helper of tests/test_jpeg_core.py: decodes a few JPEG files with the host-thread build of k_jpeg_huff (oracle/jpeg_huff_emu.py) and
compares with cv2.imdecode.  Run in a process of its own (under ThreadSanitizer the runtime must be preloaded).
usage: jpeg_huff_emu_run.py <lib.so> [one|big|huff]  (one / three 640x480 frames more; huff: one more, Huffman pass only)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2
import numpy as np

cv2.setNumThreads(1)          # OpenCV's own worker threads are not instrumented: under ThreadSanitizer they only add noise

import realset
from oracle import synth

lib = C.CDLL(sys.argv[1])


def check(enc, huff_only=False):
    enc = np.ascontiguousarray(enc, np.uint8)
    ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    out = np.zeros(ref.shape, np.uint8)
    rc = lib.jhe_decode(enc.ctypes.data_as(C.c_void_p), C.c_size_t(len(enc)), out.ctypes.data_as(C.c_void_p), 0xFF | (0x100 if huff_only else 0))
    assert rc == 0 and (huff_only or np.array_equal(out, ref)), (len(enc), rc)


small = cv2.resize(realset.image(1), (160, 120))
P = cv2
cases = [cv2.imencode('.jpg', small, [P.IMWRITE_JPEG_QUALITY, 90])[1].ravel(),
         cv2.imencode('.jpg', synth.frame(2, 123, 161), [P.IMWRITE_JPEG_QUALITY, 75, P.IMWRITE_JPEG_SAMPLING_FACTOR, P.IMWRITE_JPEG_SAMPLING_FACTOR_444])[1].ravel(),
         cv2.imencode('.jpg', cv2.cvtColor(small, cv2.COLOR_BGR2GRAY), [P.IMWRITE_JPEG_QUALITY, 95])[1].ravel(),
         cv2.imencode('.jpg', small, [P.IMWRITE_JPEG_QUALITY, 100, P.IMWRITE_JPEG_OPTIMIZE, 1])[1].ravel()]
if len(sys.argv) > 2 and sys.argv[2] == "huff":
    check(np.asarray(realset.jpeg(7)), huff_only=True)          # a real 640x480 camera frame through the Huffman pass only
elif len(sys.argv) > 2:
    cases += [np.asarray(realset.jpeg(7))]                       # a real 640x480 camera frame
    if sys.argv[2] == "big":
        cases += [np.asarray(realset.jpeg(0)), cv2.imencode('.jpg', synth.frame(1), [P.IMWRITE_JPEG_QUALITY, 90])[1].ravel()]
for e in cases:
    check(e)
print("emulated k_jpeg kernels ok:", len(cases), "files")
