"""Developer tool: JPEG-input batches (timing / profiling).  usage: gpu_jpeg_profile.py [frames] [reps] [quality]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch
import lane_slam_b200 as L
from oracle import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
q = int(sys.argv[3]) if len(sys.argv) > 3 else 90
frames = synth.sequence(n, 0)
enc = [cv2.imencode('.jpg', f, [cv2.IMWRITE_JPEG_QUALITY, q])[1].ravel() for f in frames]
off = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
blob = torch.from_numpy(np.concatenate(enc)).pin_memory().numpy()
print("jpeg bytes per frame %.0f" % (off[-1] / n))
fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(480, 640), top_cutoff=0, src_size=(480, 640), max_batch=n,
                max_segments_per_frame=256, pinned=True)
st = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE | L.STAGE_MATCH_PREV
for i in range(reps):
    fe.reset_sequence()
    t0 = time.perf_counter()
    b = fe.process_jpeg(blob, off, stages=st, k=2)
    dt = time.perf_counter() - t0
    print(i, b.n_segments, "wall %.2f ms" % (dt * 1e3), ["%s=%.3f" % x for x in fe.timings()], flush=True)
