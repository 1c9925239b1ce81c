"""Developer tool: run N frames of the bench workload twice (for ncu / timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lane_slam_b200 as L
from oracle import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
frames = synth.sequence(n, 0)
import torch
dev = torch.from_numpy(frames).cuda()
fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(480, 640), top_cutoff=0, src_size=(480, 640), max_batch=n,
                max_segments_per_frame=256)
st = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE | L.STAGE_MATCH_PREV
for i in range(reps):
    b = fe.process(dev, stages=st, k=2)
    print(i, b.n_segments, ["%s=%.3f" % x for x in fe.timings()])
