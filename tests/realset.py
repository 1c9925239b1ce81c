"""Loader of tests/golden/real_images.npz (real Duckietown frames + the reference detector's outputs on them,
made by tests/golden/make_golden_real.py).  Decoding follows duckietown_utils/jpg.py:21-31 (cv2.imdecode)."""
import os
import zlib

import cv2
import numpy as np
import yaml

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_images.npz")
COLORS = ("white", "yellow", "red")
_cache = {}


def load():
    if "g" not in _cache:
        _cache["g"] = np.load(PATH)
    return _cache["g"]


def jpeg(i):
    g = load()
    o = g["jpeg_offsets"]
    return g["jpeg_bytes"][o[i]:o[i + 1]]


def count():
    return len(load()["names"])


def image(i):
    """Decoded BGR frame i; the CRC pins the decode to what the reference saw in the authoring container."""
    if ("img", i) not in _cache:
        img = cv2.imdecode(jpeg(i), cv2.IMREAD_COLOR)
        assert np.uint32(zlib.crc32(img.tobytes())) == load()["decoded_crc"][i], "cv2.imdecode differs from the golden decode"
        _cache[("img", i)] = img
    return _cache[("img", i)]


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def golden(prefix):
    """-> dict(lines, normals, centers: concatenated white/yellow/red; counts; area / edges packed or their CRCs)."""
    g = load()
    out = dict(counts=[len(g["%s_%s_lines" % (prefix, c)]) for c in COLORS])
    for k in ("lines", "normals", "centers"):
        out[k] = np.concatenate([g["%s_%s_%s" % (prefix, c, k)] for c in COLORS])
    for c in COLORS:
        for k in ("area", "area_crc"):
            name = "%s_%s_%s" % (prefix, c, k)
            if name in g:
                out[c + "_" + k] = g[name]
    for k in ("edges", "edges_crc"):
        if "%s_%s" % (prefix, k) in g:
            out[k] = g["%s_%s" % (prefix, k)]
    return out


def yaml_sets():
    g = load()
    return {str(n): yaml.safe_load(str(g["yaml_%s_conf" % n])) for n in g["yaml_names"]}, [int(i) for i in g["yaml_frames"]]
