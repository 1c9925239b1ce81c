"""Developer tool: K13 at configuration C4 (2 000 x 100 000), both tie orders."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lane_slam_b200 as L
from oracle import synth
q, m, src = synth.descriptor_sets(2000, 100000, seed=0)
dq, dm = torch.from_numpy(q).cuda(), torch.from_numpy(m).cuda()
di = torch.empty((2000, 2), dtype=torch.int32, device="cuda"); dd = torch.empty_like(di)
for order in (L.TIES_REFERENCE, L.TIES_INDEX):
    fe = L.FrontEnd(max_batch=1, tie_order=order)
    for _ in range(5):
        fe.knn_device(dq.data_ptr(), 2000, dm.data_ptr(), 100000, 2, di.data_ptr(), dd.data_ptr(), max_dist=128)
    ts = []
    for _ in range(20):
        fe.knn_device(dq.data_ptr(), 2000, dm.data_ptr(), 100000, 2, di.data_ptr(), dd.data_ptr(), max_dist=128)
        ts.append(dict(fe.timings())["knn"])
    print("tie_order", order, "knn kernel ms: median %.3f min %.3f" % (np.median(ts), min(ts)))
    fe.close()
