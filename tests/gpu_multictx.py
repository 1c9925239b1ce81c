"""Developer tool: throughput with T contexts driven from T host threads on one GPU (batches in flight overlap:
the latency-bound LSD search of one batch runs under the issue-bound dense kernels of another).
usage: gpu_multictx.py [threads] [frames] [steps]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import lane_slam_b200 as L
from oracle import synth
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
frames = torch.from_numpy(synth.sequence(n, 0)).cuda()
st = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE | L.STAGE_MATCH_PREV
fes = [L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(480, 640), top_cutoff=0, src_size=(480, 640), max_batch=n,
                  max_segments_per_frame=256, pinned=True) for _ in range(T)]
for fe in fes:
    for _ in range(3):
        fe.reset_sequence(); b = fe.process(frames, stages=st, k=2)
ref = b.n_segments
bar = threading.Barrier(T + 1)
def work(fe):
    bar.wait()
    for _ in range(steps):
        fe.reset_sequence(); b = fe.process(frames, stages=st, k=2)
        assert b.n_segments == ref
    bar.wait()
ths = [threading.Thread(target=work, args=(fe,)) for fe in fes]
for t in ths: t.start()
torch.cuda.synchronize()
bar.wait(); t0 = time.perf_counter()
bar.wait(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
for t in ths: t.join()
print("threads %d chunk %s grow_per_sm %s: %.2f ms per %d-frame batch, %.0f frames/s" % (
    T, os.environ.get("LSF_CHUNK_FRAMES", "auto"), os.environ.get("LSF_GROW_PER_SM", "-"), 1e3 * dt / (T * steps), n, T * steps * n / dt))
