"""Developer tool: throughput of the bench workload for several chunk sizes (device and host input)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lane_slam_b200 as L
from oracle import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
frames = synth.sequence(n, 0)
pinned = torch.from_numpy(frames).pin_memory()
dev = pinned.cuda()
fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(480, 640), top_cutoff=0, src_size=(480, 640), max_batch=n,
                max_segments_per_frame=256, pinned=True)
st = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE | L.STAGE_MATCH_PREV
for chunk in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "-1,500,250,125,63").split(",")]:
    fe.set_chunk_frames(chunk)
    for name, src in (("device", dev), ("host", pinned.numpy())):
        for _ in range(2):
            fe.process(src, stages=st, k=2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            b = fe.process(src, stages=st, k=2)
        dt = (time.perf_counter() - t0) / 3
        print("chunk %4d %-6s %.2f ms  %.0f fps  S=%d  %s" % (chunk, name, dt * 1e3, n / dt, b.n_segments,
              ["%s=%.2f" % (k[:12], v) for k, v in fe.timings()]), flush=True)
