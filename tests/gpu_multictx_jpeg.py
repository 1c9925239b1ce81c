"""Developer tool: JPEG-input throughput with T contexts in flight (one host thread each)."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch
import lane_slam_b200 as L
from oracle import synth
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
gw = int(sys.argv[4]) if len(sys.argv) > 4 else 0
frames = synth.sequence(n, 0)
enc = [cv2.imencode('.jpg', f, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].ravel() for f in frames]
off = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
blobs = [torch.from_numpy(np.concatenate(enc)).pin_memory().numpy() for _ in range(T)]
st = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE | L.STAGE_MATCH_PREV
fes = [L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(480, 640), top_cutoff=0, src_size=(480, 640), max_batch=n,
                  max_segments_per_frame=256, pinned=True, grow_warps_per_sm=gw) for _ in range(T)]
for fe, bl in zip(fes, blobs):
    for _ in range(3):
        fe.reset_sequence(); b = fe.process_jpeg(bl, off, stages=st, k=2)
ref = b.n_segments
bar = threading.Barrier(T + 1)
def work(fe, bl):
    bar.wait()
    for _ in range(steps):
        fe.reset_sequence(); b = fe.process_jpeg(bl, off, stages=st, k=2)
        assert b.n_segments == ref
    bar.wait()
ths = [threading.Thread(target=work, args=(fe, bl)) for fe, bl in zip(fes, blobs)]
for t in ths: t.start()
torch.cuda.synchronize()
bar.wait(); t0 = time.perf_counter()
bar.wait(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
for t in ths: t.join()
print("jpeg: threads %d grow_warps %d: %.2f ms per %d-frame batch, %.0f frames/s" % (T, gw, 1e3 * dt / (T * steps), n, T * steps * n / dt))
