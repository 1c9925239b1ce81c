"""Helper of tests/test_jpeg_core.py: the REAL source of the JPEG kernels on host threads (oracle/jpeg_huff_emu.py), built with
AddressSanitizer, fed files whose entropy-coded bytes were corrupted (headers intact): whatever the Huffman rounds make of the
garbage, no kernel may read or write outside the buffers the library allocates for it.  usage: <lib.so> <seed> <count>"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2
import numpy as np

import realset
from oracle import synth

cv2.setNumThreads(1)
lib = C.CDLL(sys.argv[1])
rng = np.random.default_rng(int(sys.argv[2]))
small = cv2.resize(realset.image(1), (160, 120))
P = cv2
bases = [cv2.imencode('.jpg', small, [P.IMWRITE_JPEG_QUALITY, 90])[1].ravel(),
         cv2.imencode('.jpg', synth.frame(2, 123, 161), [P.IMWRITE_JPEG_QUALITY, 75, P.IMWRITE_JPEG_SAMPLING_FACTOR, P.IMWRITE_JPEG_SAMPLING_FACTOR_444])[1].ravel()]
shapes = [(120, 160, 3), (123, 161, 3)]
codes = {}
for it in range(int(sys.argv[3])):
    k = it % 2
    b = bases[k].copy()
    for _ in range(rng.integers(1, 30)):
        b[rng.integers(650, len(b) - 2)] = rng.integers(0, 256)          # entropy-coded bytes only
    out = np.zeros(shapes[k], np.uint8)
    # the first two files go through all kernels, the rest through the Huffman pass only (the other kernels' accesses do not
    # depend on the data; spawning their thousands of threads under the sanitizer is what takes the time)
    rc = lib.jhe_decode(b.ctypes.data_as(C.c_void_p), C.c_size_t(len(b)), out.ctypes.data_as(C.c_void_p), 0xFF | (0 if it < 2 else 0x100))
    codes[rc] = codes.get(rc, 0) + 1
print("kernel fuzz ok", codes)
