"""Helper of tests/test_hough_core.py: runs the host-thread build of the k_hough.cu kernels (oracle/hough_emu.py) on bit-planes made
from the oracle's colour masks / edges and compares every output row with the oracle's LineDetectorHSV.  usage: <lib.so> [small]"""
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
from oracle import hough_emu, reference_glue as rg, synth

lib = C.CDLL(sys.argv[1])
small = len(sys.argv) > 2
PB = {k: int(hough_emu._enum(k)) for k in ("PB_EDGE", "PB_BW0", "PB_EC0", "PB_COUNT")}


def pack(img, wp, garbage):
    """[h][w] bytes -> [h][wp] words, bit x & 31 of word x >> 5; padding bits of the last word set when garbage."""
    h, w = img.shape
    bits = np.zeros((h, wp * 32), np.uint8)
    bits[:, :w] = img > 0
    if garbage:
        bits[:, w:] = 1
    return np.packbits(bits.reshape(h, wp, 32), axis=2, bitorder="little").view(np.uint32).reshape(h, wp)


def check(frames, isz, cut, conf):
    n = len(frames)
    det = rg.LineDetectorHSV(dict(conf))
    h, w = isz[0] - cut, isz[1]
    wp = (w + 31) // 32
    planes = np.zeros((n, PB["PB_COUNT"], h, wp), np.uint32)
    want = []
    for f, fr in enumerate(frames):
        im = fr if isz == fr.shape[:2] else cv2.resize(fr, (isz[1], isz[0]), interpolation=cv2.INTER_NEAREST)
        im = np.ascontiguousarray(im[cut:])
        det.setImage(im)
        planes[f, PB["PB_EDGE"]] = pack(det.edges, wp, True)
        for ci, c in enumerate(rg.COLORS):
            bw = rg.color_mask_cv(det.hsv, det.cfg, c)
            planes[f, PB["PB_BW0"] + ci] = pack(bw, wp, True)
            planes[f, PB["PB_EC0"] + ci] = pack(cv2.bitwise_and(bw, det.edges), wp, True)
            want.append(det.detectLines(c))
    cap, ml = 60000, 8192
    counts = np.zeros(n * 3, np.int32); color = np.zeros(cap, np.uint8); lines = np.zeros((cap, 4), np.float32)
    normals = np.zeros((cap, 2)); centers = np.zeros((cap, 2), np.float32); pixn = np.zeros((cap, 4), np.float32); nf = np.zeros((cap, 2), np.float32)
    vp = C.c_void_p
    S = lib.hke_run(planes.ctypes.data_as(vp), n, h, w, conf["hough_threshold"], conf["hough_min_line_length"], conf["hough_max_line_gap"], cut, isz[0],
                    isz[1], ml, cap, counts.ctypes.data_as(vp), color.ctypes.data_as(vp), lines.ctypes.data_as(vp), normals.ctypes.data_as(vp),
                    centers.ctypes.data_as(vp), pixn.ctypes.data_as(vp), nf.ctypes.data_as(vp))
    assert S >= 0, S
    lo = 0
    for t, d in enumerate(want):
        cnt = len(d.lines)
        assert counts[t] == cnt, (t, counts[t], cnt)
        s = slice(lo, lo + cnt); lo += cnt
        if not cnt:
            continue
        ln = np.asarray(d.lines, np.int32)
        assert np.array_equal(color[s], np.full(cnt, t % 3, np.uint8))
        assert np.array_equal(lines[s], ln.astype(np.float32)), ("lines", t)
        assert np.array_equal(normals[s], np.asarray(d.normals, np.float64)), ("normals", t)
        assert np.array_equal(centers[s].astype(np.float64), np.asarray(d.centers, np.float64)), ("centers", t)
        px = (ln + np.array((0, cut, 0, cut))) * np.array((1. / isz[1], 1. / isz[0], 1. / isz[1], 1. / isz[0]))
        assert np.array_equal(pixn[s], px.astype(np.float32)) and np.array_equal(nf[s], np.asarray(d.normals, np.float64).astype(np.float32)), ("wire", t)
    assert lo == S
    return S


base = dict(rg.DEFAULT_DETECTOR_CONFIG)
total = 0
real = [realset.image(i) for i in (0, 4)]
total += check(real, (120, 160), 40, dict(base, hough_threshold=2, hough_min_line_length=3, hough_max_line_gap=1))
total += check([synth.frame(3, 123, 161)], (123, 161), 0, dict(base, hough_threshold=5, hough_min_line_length=4, hough_max_line_gap=2))
total += check([np.zeros((480, 640, 3), np.uint8)], (120, 160), 40, base)
if not small:
    total += check(real + [synth.frame(1)], (120, 160), 40, dict(base, hough_threshold=20, hough_min_line_length=3, hough_max_line_gap=1))
    total += check([synth.frame(0)], (480, 640), 0, dict(base, hough_threshold=20, hough_min_line_length=3, hough_max_line_gap=1))
print("emulated k_hough kernels ok:", total, "lines")
