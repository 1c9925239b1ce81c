"""Developer tool: LSD candidate trace of one frame on the GPU (LSF_TRACE_LSD=1) and on the oracle (ORC_LSD_TRACE=1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import synth, reference_glue as rg, cmodel as cm
seed, H, W, ih, iw, cut, dense = [int(x) for x in sys.argv[1:8]]
which = sys.argv[8]
img = synth.frame(seed, H, W, dense=bool(dense))
cfg = rg.check_configuration(dict(rg.DEFAULT_DETECTOR_CONFIG))
if which == "oracle":
    o = cm.front_end_frame(img, cfg, (ih, iw), cut, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY)
    print(o["counts"])
else:
    import lane_slam_b200 as L
    fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(ih, iw), top_cutoff=cut, src_size=(H, W), max_batch=1)
    b = fe.process(img)
    print(b.counts.tolist())
