"""Developer tool: the other BASELINE.json configurations on the GPU (timings only; parity is in test_gpu_parity.py).
  C3  batch of dense 1920x1080 frames: detection + descriptors
  C4  Hamming kNN 2 000 x 100 000 x 256 bit, k = 2"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lane_slam_b200 as L
from oracle import synth

which = sys.argv[1] if len(sys.argv) > 1 else "c3,c4"
if "c3" in which:
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    base = np.stack([synth.frame(s, 1080, 1920, dense=True) for s in range(4)])
    frames = np.concatenate([base] * (nb // 4))
    dev = torch.from_numpy(frames).cuda()
    cam, Hg = L.scaled_calibration(1920, 1080)
    fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(1080, 1920), top_cutoff=0, camera=cam, homography=Hg,
                    src_size=(1080, 1920), max_batch=nb, max_segments_per_frame=16384, max_segments_per_color=8192)
    st = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE
    for chunk in (-1, 0):
        fe.set_chunk_frames(chunk)
        for _ in range(2):
            b = fe.process(dev, stages=st)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        b = fe.process(dev, stages=st)
        dt = time.perf_counter() - t0
        print("C3 %d x 1080p dense chunk=%d: %.1f ms  %.0f fps  %d segments/frame  %s" % (
            nb, chunk, dt * 1e3, nb / dt, b.n_segments // nb, ["%s=%.2f" % (k[:14], v) for k, v in fe.timings()]), flush=True)
    fe.close()
if "c4" in which:
    rng = np.random.default_rng(0)
    M, Q = 100000, 2000
    m = rng.integers(0, 256, (M, 32), dtype=np.uint8)
    rows = rng.integers(0, M, Q)
    q = m[rows].copy()
    flips = rng.random((Q, 256)) < 0.12
    q ^= np.packbits(flips, axis=1)
    fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(120, 160), top_cutoff=40, max_batch=1)
    dq, dm = torch.from_numpy(q).cuda(), torch.from_numpy(m).cuda()
    di = torch.empty((Q, 2), dtype=torch.int32, device="cuda"); dd = torch.empty((Q, 2), dtype=torch.int32, device="cuda")
    for _ in range(3):
        fe.knn_device(dq.data_ptr(), Q, dm.data_ptr(), M, 2, di.data_ptr(), dd.data_ptr())
    ts = []
    for _ in range(10):
        fe.knn_device(dq.data_ptr(), Q, dm.data_ptr(), M, 2, di.data_ptr(), dd.data_ptr())
        ts.append(dict(fe.timings())["knn"])
    ok = bool((di[:, 0].cpu().numpy() == rows).all())
    t = float(np.median(ts))
    print("C4 kNN %d x %d k=2: %.3f ms  (%.2f G pairs/s, %.1f G popc32/s)  nearest==planted: %s" % (Q, M, t, Q * M / t / 1e6, Q * M * 8 / t / 1e6, ok))
