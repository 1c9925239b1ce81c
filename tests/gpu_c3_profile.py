"""Developer tool: the C3 workload (dense 1080p) for timing / profiling.  usage: gpu_c3_profile.py [batch] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lane_slam_b200 as L
from oracle import synth
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h, w = 1080, 1920
pool = np.stack([synth.frame(s, h, w, dense=True) for s in range(8)])
frames = torch.from_numpy(pool[np.arange(nb) % 8].copy()).cuda()
cam, Hg = L.scaled_calibration(w, h)
fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(h, w), top_cutoff=0, camera=cam, homography=Hg, src_size=(h, w),
                max_batch=nb, max_segments_per_frame=4096, chunk_frames=-1)
st = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE
for i in range(reps):
    b = fe.process(frames, stages=st)
    print(i, b.n_segments, ["%s=%.3f" % x for x in fe.timings()], flush=True)
