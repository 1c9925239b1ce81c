"""Generate tests/golden/hsv_detector.npz: the REFERENCE's own LineDetectorHSV (line_detector1.py, the detector eight of the ten
shipped YAML files select) imported from /root/reference and run unmodified against cv2 4.13 on synthetic and real frames, with
the threshold sets of the shipped YAML files.  Authoring container only (the GPU box has no /root/reference)."""
import os
import sys

import cv2
import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from make_golden import load, REF  # noqa: E402  (same shims: duckietown_utils.parameters only)
import make_golden  # noqa: E402


def reference_class():
    make_golden.reference_detector()            # installs the package shims
    return load("line_detector.line_detector1", REF + "/line_detector/include/line_detector/line_detector1.py").LineDetectorHSV


def main():
    from oracle import synth
    import realset
    cls = reference_class()
    out = {}
    cases = []
    ydir = REF + "/duckietown/config/baseline/line_detector/line_detector_node/"
    for name in ("universal", "bad_lighting", "226-night", "myrtle", "charles", "guy", "oreo"):    # every shipped YAML that selects this class
        cfg = yaml.safe_load(open(ydir + name + ".yaml"))
        assert cfg["detector"][0] == "line_detector.LineDetectorHSV"
        conf = cfg["detector"][1]["configuration"]
        det = cls(**cfg["detector"][1])
        frames = [("synth%d" % s, synth.frame(s)) for s in (0, 3)] + [("real%d" % i, realset.image(i)) for i in (0, 5, 11)]
        if name in ("charles", "guy", "oreo"):
            frames = frames[1:4]
        for tag, img in frames:
            for (isz, cut) in (((120, 160), 40), ((480, 640), 0)):
                im = img if isz == img.shape[:2] else cv2.resize(img, (isz[1], isz[0]), interpolation=cv2.INTER_NEAREST)
                im = im[cut:, :, :]
                det.setImage(im)
                key = "%s/%s/%dx%d" % (name, tag, isz[0], isz[1])
                cases.append(key)
                for c in ("white", "yellow", "red"):
                    d = det.detectLines(c)
                    n = len(d.lines)
                    out[key + "/" + c + "/lines"] = np.asarray(d.lines, np.int32).reshape(n, 4)
                    out[key + "/" + c + "/normals"] = np.asarray(d.normals, np.float64).reshape(n, 2)
                    out[key + "/" + c + "/centers"] = np.asarray(d.centers, np.float64).reshape(n, 2)
        out["cfg/" + name] = np.array([repr(conf)])
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "hsv_detector.npz"), **out)
    print("wrote hsv_detector.npz:", len(cases), "cases,", sum(len(v) for k, v in out.items() if k.endswith("/lines")), "lines")


if __name__ == "__main__":
    main()
