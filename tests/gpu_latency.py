"""Developer tool: single-frame latency breakdown."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lane_slam_b200 as L
from oracle import synth
frames = torch.from_numpy(synth.sequence(64, 0)).pin_memory().numpy()
fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(480, 640), top_cutoff=0, src_size=(480, 640), max_batch=1,
                max_segments_per_frame=1024, pinned=True)
st = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE | L.STAGE_MATCH_PREV
ts = []; acc = {}
for i in range(264):
    t0 = time.perf_counter()
    b = fe.process(frames[i % 64:i % 64 + 1], stages=st, k=2)
    ts.append(time.perf_counter() - t0)
    if i >= 64:
        for k, v in fe.timings(): acc[k] = acc.get(k, 0) + v / 200
ts = np.array(ts[64:]) * 1e3
print("p50 %.3f p95 %.3f ms; gpu stages (mean ms): %s; sum %.3f" % (np.percentile(ts, 50), np.percentile(ts, 95), {k: round(v, 3) for k, v in acc.items()}, sum(acc.values())))
