"""Generate tests/golden/lane_filter_votes.npz by importing the REFERENCE's own LaneFilterHistogram
(/root/reference/src/lane_filter/include/lane_filter/lane_filter.py) and running its unmodified
generate_measurement_likelihood on the ground segments the oracle produces for synthetic frames.
Run in the authoring container only (the GPU box has no /root/reference).

Shims: duckietown_utils is loaded as a bare package holding only parameters.py; duckietown_msgs.msg.Segment is a
three-constant stand-in (the filter reads .color, .points[k].x/.y and the WHITE / YELLOW constants only).
"""
import importlib.util
import os
import sys
import types

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


class Pt(object):
    def __init__(self, x, y):
        self.x, self.y, self.z = x, y, 0.0


class Segment(object):
    WHITE, YELLOW, RED = 0, 1, 2

    def __init__(self, color, g):
        self.color = int(color)
        self.points = [Pt(float(g[0]), float(g[1])), Pt(float(g[2]), float(g[3]))]


def reference_filter():
    du = types.ModuleType("duckietown_utils"); du.__path__ = []; sys.modules["duckietown_utils"] = du
    load("duckietown_utils.parameters", REF + "/duckietown/include/duckietown_utils/parameters.py")
    msgs = types.ModuleType("duckietown_msgs"); msgs.__path__ = []; sys.modules["duckietown_msgs"] = msgs
    msg = types.ModuleType("duckietown_msgs.msg"); msg.Segment = Segment; sys.modules["duckietown_msgs.msg"] = msg
    pkg = types.ModuleType("lane_filter"); pkg.__path__ = [REF + "/lane_filter/include/lane_filter"]
    sys.modules["lane_filter"] = pkg
    load("lane_filter.lane_filter_interface", REF + "/lane_filter/include/lane_filter/lane_filter_interface.py")
    lf = load("lane_filter.lane_filter", REF + "/lane_filter/include/lane_filter/lane_filter.py")
    cfg = yaml.safe_load(open(REF + "/duckietown/config/baseline/lane_filter/lane_filter_node/default.yaml"))
    return lf.LaneFilterHistogram(cfg["filter"][1]["configuration"])


def main():
    from oracle import cmodel as cm, reference_glue as rg, synth
    flt = reference_filter()
    cfg = rg.check_configuration(dict(rg.DEFAULT_DETECTOR_CONFIG))
    out = {}
    seeds = [0, 1, 2, 7, 23, 40]
    for k, seed in enumerate(seeds):
        img = synth.frame(seed)
        o = cm.front_end_frame(img, cfg, (480, 640), 0, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY)
        segs = [Segment(c, g) for c, g in zip(o["color"], o["ground"])]
        ml = flt.generate_measurement_likelihood(segs)
        out["ground_%d" % k] = np.asarray(o["ground"], np.float64)
        out["color_%d" % k] = np.asarray(o["color"], np.uint8)
        out["likelihood_%d" % k] = np.zeros(flt.d.shape) if ml is None else ml
    out["seeds"] = np.array(seeds)
    # the whole filter, as LaneFilterNode.processSegments drives it (lane_filter_node.py:53-65), over a 24-frame sequence
    flt.initialize()
    frames = synth.sequence(24, base_seed=60)
    rng = np.random.default_rng(4)
    dvw = np.stack([rng.uniform(0.05, 0.15, 24), rng.uniform(0.0, 0.4, 24), rng.uniform(-1.5, 1.5, 24)], axis=1)
    est = np.zeros((24, 3))
    out["filter_belief0"] = flt.belief.copy()
    for t in range(24):
        o = cm.front_end_frame(frames[t], cfg, (480, 640), 0, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY)
        flt.predict(dt=float(dvw[t, 0]), v=float(dvw[t, 1]), w=float(dvw[t, 2]))
        flt.update([Segment(c, g) for c, g in zip(o["color"], o["ground"])])
        d_max, phi_max = flt.getEstimate()
        est[t] = [d_max, phi_max, flt.getMax()]
        if t in (0, 11, 23):
            out["filter_belief_%d" % t] = flt.belief.copy()
    out["filter_dvw"] = dvw
    out["filter_estimates"] = est
    np.savez_compressed(os.path.join(HERE, "lane_filter_votes.npz"), **out)
    print("wrote lane_filter_votes.npz:", {k: int(out["likelihood_%d" % k].astype(bool).sum()) for k in range(len(seeds))})


if __name__ == "__main__":
    main()
