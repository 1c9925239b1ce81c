"""Generate tests/golden/reference_detections.npz by importing the REFERENCE's own Python files from
/root/reference (LineDetectorLSD, Configurable) and running them, unmodified, against cv2 4.13 on the
synthetic frames of oracle/synth.py.  Run in the authoring container only (the GPU box has no /root/reference).

Shims (the reference targets Python 2 / OpenCV 3):
  * duckietown_utils is loaded as a bare package holding only parameters.py (its __init__ needs ROS modules);
  * cv2.createLineSegmentDetector(_refine=...) -> OpenCV 4 names the kwarg `refine`; forwarded positionally.
The node-level arithmetic (normalisation, ground projection, sanity) is not importable without ROS; the golden
file therefore pins the detector plugin (lines / normals / centers / area) which is where all the image
arithmetic lives, plus cv2.undistortPoints for the default calibration.
"""
import importlib.util
import os
import sys
import types

import cv2
import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/src"


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def reference_detector():
    du = types.ModuleType("duckietown_utils"); du.__path__ = []; sys.modules["duckietown_utils"] = du
    load("duckietown_utils.parameters", REF + "/duckietown/include/duckietown_utils/parameters.py")
    pkg = types.ModuleType("line_detector"); pkg.__path__ = [REF + "/line_detector/include/line_detector"]
    sys.modules["line_detector"] = pkg
    load("line_detector.line_detector_interface", REF + "/line_detector/include/line_detector/line_detector_interface.py")
    lsd = load("line_detector.line_detector_lsd", REF + "/line_detector/include/line_detector/line_detector_lsd.py")
    orig = cv2.createLineSegmentDetector
    cv2.createLineSegmentDetector = lambda _refine=1, **k: orig(_refine)
    cfg = yaml.safe_load(open(REF + "/duckietown/config/baseline/line_detector/line_detector_node/default.yaml"))
    return lsd.LineDetectorLSD(**cfg["detector"][1]), cfg


def main():
    from oracle import synth
    det, cfg = reference_detector()
    out = {}
    cases = []
    # (seed, H, W, img_size, top_cutoff): the reference default (160x120 cut 40) and a native 320x240 frame
    for seed in range(6):
        cases.append((seed, 480, 640, (120, 160), 40))
    for seed in range(6, 9):
        cases.append((seed, 240, 320, (240, 320), 0))
    out["cases"] = np.array([[c[0], c[1], c[2], c[3][0], c[3][1], c[4]] for c in cases], np.int32)
    for i, (seed, H, W, isz, cut) in enumerate(cases):
        img = synth.frame(seed, H, W)
        # line_detector_node.py:163-175 (identity colour transform)
        if isz != (H, W):
            img = cv2.resize(img, (isz[1], isz[0]), interpolation=cv2.INTER_NEAREST)
        img = img[cut:, :, :]
        det.setImage(img)
        for c in ("white", "yellow", "red"):
            d = det.detectLines(c)
            n = len(d.lines)
            out["%d_%s_lines" % (i, c)] = np.asarray(d.lines, np.float32).reshape(n, 4)
            out["%d_%s_normals" % (i, c)] = np.asarray(d.normals, np.float64).reshape(n, 2)
            out["%d_%s_centers" % (i, c)] = np.asarray(d.centers, np.float32).reshape(n, 2)
            out["%d_%s_area" % (i, c)] = np.packbits(d.area > 0)
        out["%d_edges" % i] = np.packbits(det.edges > 0)
        out["%d_hsv_sum" % i] = det.hsv.astype(np.int64).sum(axis=(0, 1))
    # undistortPoints known answers for the default calibration
    ci = yaml.safe_load(open(REF + "/duckietown/config/baseline/calibration/camera_intrinsic/default.yaml"))
    K = np.array(ci["camera_matrix"]["data"], float).reshape(3, 3); D = np.array(ci["distortion_coefficients"]["data"], float)
    R = np.array(ci["rectification_matrix"]["data"], float).reshape(3, 3); P = np.array(ci["projection_matrix"]["data"], float).reshape(3, 4)
    rng = np.random.default_rng(0)
    uv = np.stack([rng.uniform(0, 639, 64), rng.uniform(0, 479, 64)], -1)
    out["undist_uv"] = uv
    out["undist_out"] = cv2.undistortPoints(uv.reshape(-1, 1, 2), K, D, R=R, P=P).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "reference_detections.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_detections.npz"), "cv2", cv2.__version__)


if __name__ == "__main__":
    main()
