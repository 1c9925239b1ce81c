#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 line front end.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores

Workload (BASELINE.json configs[1]): a 1 000-frame synthetic 640x480 sequence through the whole front end:
colour masks + Canny + 3x LSD -> normals -> LBD descriptors -> ground projection + line_sanity ->
frame-to-frame Hamming association (k = 2).  One "step" = one pass over the 1 000-frame batch.
  value : frames/s with the frames already resident in HBM (device pointer handed to the C ABI)
  e2e   : frames/s through the same C-ABI call with HOST (pinned) frames: H2D of the frames and D2H of
          the segment lists inside the timed region
Multi-GPU (torchrun, one rank per GPU): every rank runs the same 1 000-frame shard size (weak scaling),
then the ranks all-gather their kept-segment lists and descriptors with NCCL (the path's only exchange step).
Inputs are 921.6 MB per step per GPU (> 126 MB L2), so no L2 flush is needed between iterations.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 480, 640
N_FRAMES = 1000
K_NN = 2
METRIC = "frames/sec (640x480 front end)"
WORKLOAD = "1000-frame synthetic 640x480 sequence: LSD + LBD descriptors + line_sanity + frame-to-frame association"

# ALGORITHMIC bytes per frame of the dense kernels (N = 307 200 pixels; DESIGN.md "Roofline")
N_PIX = H * W
ALGO_BYTES = {
    "color_canny": 4.0 * N_PIX,          # K2+K3: read 3N BGR, write N (labels + NMS class)
    "hysteresis_dilate": 4.0 * N_PIX,    # K4 (read N, write N) + K5 (read N, write N)
    "lsd_pre": 8.68 * N_PIX,             # K6 (read N, write 3*0.64N) + K7 (read 3*0.64N, write 2 B x 3*0.64N)
    "gray_sobel": 5.0 * N_PIX,           # K11a: read N gray, write 4N (dx, dy int16)
}


# DRAM traffic per frame (dram__bytes_read.sum + dram__bytes_write.sum) of the dense kernels, from the ncu --set full
# capture profiles/r01f_ncu_full_296frames.csv (296 frames per launch, divided by 296)
TRAFFIC_SOURCE = "ncu --set full, profiles/r01f_ncu_full_296frames.csv, scaled from 296 to %d frames"
TRAFFIC_BYTES = {
    "color_canny": (272.836e6 + 151.673e6) / 296,
    "hysteresis_dilate": (56.880e6 + 28.300e6) / 296,
    "lsd_pre": (34.587e6 + 8.536e6 + 42.403e6 + 16.717e6) / 296,
    "gray_sobel": (92.231e6 + 306.249e6) / 296,
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def bind_to_gpu_numa_node(torch, local):
    """Multi-rank runs: keep this rank (and so its pinned frame buffers, first-touched below) on the CPUs NVML reports as
    local to its GPU, so that eight ranks do not pull their host->device copies across the socket interconnect.
    Best effort: any failure, or an empty intersection with the CPUs the container allows, leaves the affinity alone."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        local_cpus = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1}
        allowed = set(os.sched_getaffinity(0)) & local_cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        print("[bench] rank on cuda:%d: %d of %d allowed CPUs are local to the GPU -> affinity %s" % (
            local, len(allowed), len(os.sched_getaffinity(0)) if not allowed else len(allowed), "set" if allowed else "unchanged"),
            file=sys.stderr, flush=True)
    except Exception as e:
        print("[bench] NUMA binding skipped: %r" % (e,), file=sys.stderr, flush=True)


def make_frames(n, base_seed, start=0, h=H, w=W, dense=False):
    from oracle import synth  # input generator only (test/bench infrastructure)
    return synth.sequence(n, base_seed=base_seed, H=h, W=w, dense=dense, start=start)


EXTRAS_DEADLINE_S = 300          # extra configurations (c5 runs collectives): give up after this many seconds at N > 1
JPEG_QUALITY = 90


def encode_jpeg(frames):
    """The camera delivers JPEG (sensor_msgs/CompressedImage, decoded by duckietown_utils/jpg.py:21-31 with cv2.imdecode): the
    synthetic frames are encoded once, outside every timed region (cv2.imencode, quality 90, 4:2:0).  Returns (blob uint8,
    offsets int64 [n+1])."""
    import cv2
    enc = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, JPEG_QUALITY])[1].ravel() for f in frames]
    off = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
    return np.concatenate(enc), off


def decode_jpeg(blob, off):
    import cv2
    return np.stack([cv2.imdecode(blob[off[i]:off[i + 1]], cv2.IMREAD_COLOR) for i in range(len(off) - 1)])


def workload_config(frames_per_gpu):
    """The `config` object of the JSON line: identical for the GPU arm and the reference (CPU) arm -- same generator, same
    seeds (rank r replays synth.sequence(base_seed = r * frames_per_gpu)), same detector configuration and stages."""
    return {"workload": WORKLOAD, "frames_per_step_per_gpu": frames_per_gpu, "img_size": [H, W], "top_cutoff": 0, "k": K_NN,
            "detector": "line_detector_node/default.yaml thresholds", "stages": "detect+ground+sanity+describe+match_prev",
            "input": "synth.sequence(base_seed = rank * frames_per_step_per_gpu) as JPEG files (cv2.imencode quality %d, 4:2:0): "
                     "e2e / reference arm start from the JPEG bytes (decode = cv2.imdecode semantics), `value` from the decoded BGR "
                     "frames resident in HBM" % JPEG_QUALITY, "match_radius": 128,
            "l2": "inputs 921.6 MB per step exceed the 126 MB L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle port: restated glue + cv2 4.13, C LBD, BFMatcher)
# ------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import cv2
    cv2.setNumThreads(1)
    from oracle import cmodel as cm, reference_glue as rg
    start, count = args
    blob, off = encode_jpeg(make_frames(count, 0, start=start))     # frames [start, start + count) of rank 0's sequence, as JPEG
    det = rg.LineDetectorLSD(dict(rg.DEFAULT_DETECTOR_CONFIG))
    gp = rg.GroundProjection()
    prev = None
    t0 = time.perf_counter()
    nseg = 0
    for f in range(count):
        frame = cv2.imdecode(blob[off[f]:off[f + 1]], cv2.IMREAD_COLOR)      # jpg.py:21-31, line_detector_node.py:155
        r = rg.front_end_frame(frame, det, gp, (H, W), 0)
        gray = cv2.cvtColor(r["image"], cv2.COLOR_BGR2GRAY)
        blur = cv2.GaussianBlur(gray, (5, 5), 1)
        dx = cv2.Sobel(blur, cv2.CV_16S, 1, 0, ksize=3)
        dy = cv2.Sobel(blur, cv2.CV_16S, 0, 1, ksize=3)
        desc = cm.lbd(r["lines_px"], dx, dy)[2] if len(r["lines_px"]) else np.zeros((0, 32), np.uint8)
        if prev is not None and len(prev) and len(desc):
            rg.knn_hamming_bf(desc, prev, min(K_NN, len(prev)))     # cv2.BFMatcher: all the matching work, vectorised
        prev = desc
        nseg += len(desc)
    return time.perf_counter() - t0, count, nseg


def cpu_baseline(sample_frames, cores):
    """Frames/s of the CPU path with `cores` single-threaded workers, on `sample_frames` frames of the workload."""
    import multiprocessing as mp
    per = max(1, sample_frames // cores)
    # contiguous pieces spread over the 1000-frame sequence the GPU arm replays (same seeds, same arrays)
    jobs = [((i * N_FRAMES) // cores, per) for i in range(cores)]
    ctx = mp.get_context("spawn")     # never fork a process that may hold CUDA / OpenMP state
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    return total / busy, total, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, len(os.sched_getaffinity(0)))
    sample = max(cores * 16, 128)          # about 10 core-seconds of CPU work per step
    vals = []
    for i in range(args.warmup + args.steps):
        fps, total, wall = cpu_baseline(sample, cores)
        if i >= args.warmup:
            vals.append((fps, total, wall))
    fps = float(np.mean([v[0] for v in vals]))
    total = vals[0][1]
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / fps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/f64", "data": "synthetic",
        "config": workload_config(args.frames),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d frames per step = %d contiguous pieces of the same 1000-frame sequence (same seeds), one "
                                   "single-threaded cv2 worker per core (cv2.imdecode of the JPEG frame + restated reference glue + cv2 4.13 "
                                   "LSD/Canny, C LBD, cv2.BFMatcher)" % (total, cores)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def parity_gate(L, b, frames_np, n_check=64):
    """BASELINE.md 3 / SURVEY 8d: the parity gate on the BENCHMARKED data.  n_check frames spread over the step are redone by the
    CPU oracle (oracle/: test infrastructure, used here only as the checker, outside every timed region) and compared with
    what the timed GPU step returned: segment counts per colour, endpoints, ground points, sanity keep mask, every bit of every
    descriptor and the frame-to-frame match indices given identical descriptors."""
    from oracle import cmodel as cm, reference_glue as rg
    cfg = rg.check_configuration(dict(rg.DEFAULT_DETECTOR_CONFIG))
    n = len(frames_np)
    idx = sorted(set(np.linspace(1, n - 1, n_check).round().astype(int).tolist()))
    st = dict(frames=len(idx), exact_frames=0, count_mismatch_frames=0, max_endpoint_px=0.0, max_ground_m=0.0, desc_bits_bad=0,
              desc_bits=0, match_rows_bad=0, match_rows=0, segments=0)
    for f in idx:
        o = cm.front_end_frame(frames_np[f], cfg, (H, W), 0, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY, descriptors=True)
        g = b.frame(f)
        ok = g["counts"] == o["counts"]
        if not ok:
            st["count_mismatch_frames"] += 1
            continue
        S = len(o["lines_px"])
        st["segments"] += S
        if S:
            st["max_endpoint_px"] = max(st["max_endpoint_px"], float(np.abs(g["lines_px"] - o["lines_px"]).max()))
            fin = np.isfinite(o["ground"]) & np.isfinite(g["ground"])
            if fin.any():
                st["max_ground_m"] = max(st["max_ground_m"], float(np.abs(g["ground"][fin] - o["ground"][fin]).max()))
            bad_bits = int(np.unpackbits(g["desc"] ^ o["desc32"]).sum())
            st["desc_bits_bad"] += bad_bits; st["desc_bits"] += S * 256
            ok = ok and np.array_equal(g["keep"], o["keep"]) and bad_bits == 0 and st["max_endpoint_px"] <= 0.5 and st["max_ground_m"] <= 1e-4
            prev = b.frame(f - 1)["desc"]
            if len(prev):
                oi, od = cm.knn_mihasher(o["desc32"], prev, K_NN)       # the reference's own result order (Mihasher), radius 128
                sl = b.frame_slice(f)
                bad_rows = int((b.match_idx[sl] != oi).any(axis=1).sum() + 0)
                st["match_rows_bad"] += bad_rows; st["match_rows"] += S
                ok = ok and bad_rows == 0 and np.array_equal(b.match_dist[sl], od)
        st["exact_frames"] += bool(ok)
    st["pass_fraction"] = st["exact_frames"] / max(1, st["frames"])
    st["gate"] = ">= 0.95 of frames: counts / keep / descriptor bits / match indices exact, endpoints <= 0.5 px, ground <= 1e-4 m"
    st["checker"] = "oracle/csrc/lane_oracle.c (pinned to the reference class, cv2 4.13 and the reference's compiled line_descriptor)"
    return st


def run_gpu(args):
    import torch
    import lane_slam_b200 as L
    from lane_slam_b200 import odometry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        bind_to_gpu_numa_node(torch, local)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.frames
    t_start = time.perf_counter()

    def log(msg):
        if rank == 0 and args.verbose:
            print("[bench %.1fs] %s" % (time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)

    raw_np = make_frames(n, base_seed=rank * n)               # this rank's shard of the log ...
    blob_np, off = encode_jpeg(raw_np)                        # ... as the JPEG files the camera delivers
    frames_np = decode_jpeg(blob_np, off)                     # = cv2.imdecode of them: "the frames" of every arm
    del raw_np
    log("frames generated, %.1f KB of JPEG per frame" % (off[-1] / n / 1e3))
    pinned = torch.from_numpy(frames_np).pin_memory()
    dev = pinned.cuda(non_blocking=False)
    T = max(1, args.inflight)                                 # contexts (= batches) in flight on this GPU, one host thread each
    fes = [L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(H, W), top_cutoff=0, src_size=(H, W), max_batch=n,
                      device=local, max_segments_per_frame=256, pinned=True, chunk_frames=(-1 if T > 1 else 0)) for _ in range(T)]
    fe = fes[0]
    blobs = [torch.from_numpy(blob_np).pin_memory().numpy() for _ in range(T)]
    stages = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE | L.STAGE_MATCH_PREV
    # the path's one exchange step lives in liblsf.so: one ncclAllGather on the ctx's exchange stream, overlapped with the next step
    for c in range(T):
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if world > 1:
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(odometry.nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
        fes[c].exchange_init(rank=rank, world=world, unique_id=bytes(uid.cpu().numpy()) if world > 1 else None)
    inflight = [0] * T
    kept_seen = [0]
    last = [None] * T

    def run_batch(c, kind):
        f = fes[c]
        f.reset_sequence()
        if kind == "jpeg":
            b = f.process_jpeg(blobs[c], off, stages=stages, k=K_NN)
        else:
            b = f.process(dev if kind == "dev" else pinned.numpy(), stages=stages, k=K_NN)
        last[c] = b
        return b

    def exchange(c):
        if world > 1:
            f = fes[c]
            if inflight[c]:                                   # the exchange of this context's previous step ran under this step's kernels
                _, ntot, _ = f.exchange_wait(); inflight[c] -= 1; kept_seen[0] = ntot
            f.allgather_start(frame_base=rank * n); inflight[c] += 1

    def step(c, kind):
        b = run_batch(c, kind)
        exchange(c)
        return b

    def drain(c):
        while inflight[c]:
            _, ntot, _ = fes[c].exchange_wait(); inflight[c] -= 1; kept_seen[0] = ntot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    exts = [torch.cuda.ExternalStream(f.stream(), device=torch.device("cuda", local)) for f in fes]   # the streams liblsf launches on

    def timed(kind, steps, ncontexts=T, stream_ahead=False):
        """K steps bracketed by barrier + synchronize, timed on the device with CUDA events recorded on the streams the kernels are
        launched on (the lsf ctx streams).  With `ncontexts` > 1 the steps are dealt round robin to that many contexts, each driven
        by its own host thread, so that consecutive batches overlap on the GPU (the latency-bound LSD search of one under the
        issue-bound kernels of another; for JPEG input also the host-side header parsing, the H2D copy and the decode).  The
        last exchange of every context is finished before its closing event.  stream_ahead (raw host frames, one context):
        fe.prefetch() stages step i+1 while step i computes.  Returns (seconds, per-stage ms of context 0, d2h bytes, last batch)."""
        import threading
        stage_ms = {}
        d2h = [0]
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(ncontexts)]
        errors = []
        start = threading.Barrier(ncontexts + 1)

        def work(c):
            try:
                torch.cuda.set_device(local)
                start.wait()
                for i in range(c, steps, ncontexts):
                    if stream_ahead and i + 1 < steps:
                        fes[c].prefetch(pinned.numpy())
                    try:
                        b = run_batch(c, kind)
                    except Exception as e:      # the exchange below still runs (previous records): the ranks' collectives stay matched
                        errors.append(repr(e))
                        b = None
                    exchange(c)
                    if c == 0 and b is not None:
                        for name, ms in fes[0].timings():
                            stage_ms[name] = stage_ms.get(name, 0.0) + ms
                        S = b.n_segments
                        d2h[0] = S * (1 + 16 + 16 + 8 + 16 + 8 + 32 + 1 + 32 + 8 * K_NN) + (4 * n + 1) * 4
                drain(c)
                e1[c].record(exts[c])
            except Exception as e:      # surfaced below
                errors.append(repr(e))
                try:
                    start.abort()
                except Exception:
                    pass
        ths = [threading.Thread(target=work, args=(c,)) for c in range(ncontexts)]
        for t in ths:
            t.start()
        barrier()
        e0.record(exts[0])
        if stream_ahead:
            fes[0].prefetch(pinned.numpy())
        start.wait()
        for t in ths:
            t.join()
        failed = 1 if errors else 0
        if world > 1:                                        # every rank takes the same branch below
            ft = torch.tensor([failed], dtype=torch.int32, device="cuda")
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
            failed = int(ft.item())
        barrier()
        if failed:
            raise RuntimeError("bench step failed (%s input, %d context(s)): %s" % (kind, ncontexts, errors or "on another rank"))
        dt = max(e0.elapsed_time(e) for e in e1) * 1e-3
        return dt, stage_ms, d2h[0], last[0]

    def timed_secondary(kind, steps, **kw):
        """A measurement that explains the headline but is not the headline: a failure is reported as null, not fatal."""
        try:
            return timed(kind, steps, **kw)
        except RuntimeError as e:
            print("[bench] %s" % e, file=sys.stderr, flush=True)
            return float("nan"), {}, 0, None

    log("contexts ready")
    # warm-up (>= 3 per context and input kind)
    for c in range(T):
        for i in range(max(3, args.warmup)):
            b = step(c, "dev")
        step(c, "jpeg"); step(c, "jpeg")
        drain(c)
    step(0, "raw"); drain(0)
    log("warm-up done: S=%d" % b.n_segments)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sum(f.launch_count() for f in fes)
    dt_dev, _, _, b = timed("dev", args.steps)
    launches = sum(f.launch_count() for f in fes) - l0
    parity = parity_gate(L, b, frames_np, args.parity_frames) if (rank == 0 and args.parity_frames > 0) else None
    seg_per_step, kept_per_step = int(b.n_segments), int(b.keep.sum())
    dt_e2e, _, d2h_bytes, bj = timed("jpeg", args.steps)
    jpeg_same = bool(bj.n_segments == b.n_segments and np.array_equal(bj.lines_px, b.lines_px) and np.array_equal(bj.desc, b.desc))
    fe.set_chunk_frames(0)                                                         # one batch at a time: chunk pipeline inside the batch
    step(0, "dev"); step(0, "jpeg"); drain(0)
    dt_dev1, _, _, _ = timed_secondary("dev", args.steps, ncontexts=1)
    dt_jpeg1, jpeg_stage_ms, _, _ = timed_secondary("jpeg", args.steps, ncontexts=1)
    dt_raw, _, _, _ = timed_secondary("raw", args.steps, ncontexts=1, stream_ahead=True)   # round-1 e2e: raw BGR host frames, staged one step ahead
    # per-kernel times for the roofline: same steps on ONE stream (chunk pipeline off) so that the library's
    # CUDA events bracket each kernel; not part of `value` / `e2e`
    fe.set_chunk_frames(-1)
    step(0, "dev")
    _, stage_ms, _, _ = timed("dev", args.steps, ncontexts=1)
    fe.set_chunk_frames(0)
    clocks = sampler.stop() if rank == 0 else None
    # p50 latency of one frame through the same call (batch = 1, host frame in, segment list out), rank 0 only
    lat = None
    if rank == 0:
        fe1 = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(H, W), top_cutoff=0, src_size=(H, W), max_batch=1,
                         device=local, max_segments_per_frame=1024, pinned=True)
        one = pinned.numpy()
        ts = []
        for i in range(230):
            t0 = time.perf_counter()
            fe1.process(one[i % n:i % n + 1], stages=stages, k=K_NN)
            ts.append(time.perf_counter() - t0)
        fe1.close()
        ts = np.array(ts[30:]) * 1e3
        lat = {"batch1_p50_ms": float(np.percentile(ts, 50)), "batch1_p95_ms": float(np.percentile(ts, 95)), "frames": int(len(ts)),
               "how": "host wall clock around FrontEnd.process of one pinned host frame (all stages, k=2), after 30 warm-up frames"}
    extra = {}
    for f in fes[1:]:
        f.close()
    if world > 1:
        t = torch.tensor([dt_dev, dt_e2e, dt_dev1, dt_jpeg1, dt_raw], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_dev, dt_e2e, dt_dev1, dt_jpeg1, dt_raw = [float(x) for x in t]
    finish_lock = threading.Lock()
    finished = [False]

    def finish():
        """Rank 0: build and print the one JSON line (once)."""
        with finish_lock:
            if finished[0]:
                return
            finished[0] = True
            finish_locked()

    def finish_locked():
        total_frames = n * world * args.steps
        value = total_frames / dt_dev
        e2e = total_frames / dt_e2e

        def rate(dt):
            return total_frames / dt if np.isfinite(dt) and dt > 0 else None

        def ms_step(dt):
            return 1e3 * dt / args.steps if np.isfinite(dt) and dt > 0 else None
        peak, peak_src = peaks()
        kernels = []
        tot_ms = sum(v for k, v in stage_ms.items() if k not in ("h2d", "d2h"))
        for name, ms in sorted(stage_ms.items(), key=lambda kv: -kv[1]):
            if name in ("h2d", "d2h"):
                continue
            per = ms / args.steps
            ent = {"kernel": name, "ms_per_step": per, "share": ms / tot_ms}
            if name in ALGO_BYTES:
                gbs = ALGO_BYTES[name] * n / (per * 1e-3) / 1e9
                ent.update(bound="hbm", algorithmic_bytes_per_frame=ALGO_BYTES[name], achieved_gbs=gbs, frac=gbs / peak,
                           dram_traffic_bytes_per_launch=TRAFFIC_BYTES[name] * n)
            else:
                ent.update(bound="latency")
            kernels.append(ent)
        dense = [k for k in kernels if k["bound"] == "hbm"]
        dom = max(dense, key=lambda k: k["ms_per_step"])
        roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac"], "traffic": dom["dram_traffic_bytes_per_launch"],
                    "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_frame"] * n, "peak_source": peak_src,
                    "traffic_source": TRAFFIC_SOURCE % n,
                    "note": "dominant HBM-bound kernel; the LSD search (lsd_core) is latency-bound, see 'kernels'"}
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": 1e3 * dt_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f64", "data": "synthetic",
            "config": workload_config(n),
            "pipeline": {"contexts_in_flight": T,
                         "how": "%d contexts per GPU, one host thread each, consecutive steps dealt round robin (batches overlap on the GPU); "
                                "with several contexts every batch runs on one stream (the chunk pipeline is for one batch at a time)" % T,
                         "one_batch_at_a_time": {"value": rate(dt_dev1), "ms_per_step": ms_step(dt_dev1),
                                                 "e2e_jpeg_value": rate(dt_jpeg1), "e2e_jpeg_ms_per_step": ms_step(dt_jpeg1)},
                         "segments_per_step": seg_per_step, "kept_per_step": kept_per_step,
                         "exchange": None if world == 1 else "lsf_allgather_segments: one ncclAllGather of fixed-capacity slots on the ctx's exchange "
                                                            "stream, started after step i and finished under step i+1 (last one inside the timed region); "
                                                            "%d records gathered per step" % kept_seen[0]},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(off[-1]) + n * (16 + 384) + 8192, "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": 1e3 * dt_e2e / args.steps,
                    "how": "the frames as JPEG files in pinned host memory through FrontEnd.process_jpeg (lsf_front_end_batch_jpeg): every step "
                           "parses the headers, copies the compressed bytes to the device, decodes them on the GPU (bit-identical to "
                           "cv2.imdecode), runs the whole front end and copies the segment lists back, all inside the timed region",
                    "jpeg_bytes_per_frame": float(off[-1]) / n, "jpeg_quality": JPEG_QUALITY,
                    "same_result_as_value_arm": jpeg_same,
                    "decode_stage_ms_per_step": {k: v / args.steps for k, v in jpeg_stage_ms.items() if k.startswith("jpeg")},
                    "raw_bgr": {"value": rate(dt_raw), "ms_per_step": ms_step(dt_raw), "h2d_bytes_per_step": int(n * H * W * 3),
                                "how": "round-1 definition: decoded BGR frames in pinned host memory, H2D staged one step ahead (PCIe-bound)"}},
            "gpu_launches": int(launches),
            "latency": dict(lat, batch1000_ms_per_frame_amortised=1e3 * dt_dev / args.steps / n,
                            batch1000_step_ms=1e3 * dt_dev / args.steps) if lat else None,
            "roofline": roofline, "kernels": kernels, "clocks": clocks, "parity": parity, "extra_configs": extra,
        }
        if world == 1 and not args.no_cpu:
            # the CPU leg runs in a fresh interpreter (no CUDA state), bounded in time
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                                     capture_output=True, text=True, timeout=600)
                ref = json.loads(out.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
            except Exception as e:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                        "sample": "failed: %r" % (e,)}
        emit(line)

    if world > 1 and args.extras:
        # The headline is measured; the extra configurations below still run collectives.  Should one of them never return
        # (a rank that failed leaves the others waiting), the line is printed without it and every rank leaves.
        def watchdog():
            time.sleep(EXTRAS_DEADLINE_S)
            if finished[0]:
                return
            if rank == 0:
                extra.setdefault("c5", {"error": "not finished after %d s, abandoned" % EXTRAS_DEADLINE_S})
                try:
                    finish()
                finally:
                    os._exit(0)
            time.sleep(20)
            os._exit(0)
        threading.Thread(target=watchdog, daemon=True).start()
    if "c5" in args.extras:
        try:
            extra["c5"] = bench_c5(torch, dist, L, fe, dev, n, rank, world, local, args, log)
        except Exception as e:                    # the headline stands on its own
            print("[bench] c5 failed: %r" % (e,), file=sys.stderr, flush=True)
            extra["c5"] = {"error": repr(e)}
            if world > 1:                         # the ranks are out of step: no more collectives
                if rank == 0:
                    finish()
                os._exit(0)
    fe.close()
    del dev
    torch.cuda.empty_cache()
    if rank == 0 and world == 1:
        if "c4" in args.extras:
            extra["c4"] = bench_c4(torch, L, local, log)
        if "c3" in args.extras:
            extra["c3"] = bench_c3(torch, L, local, args, log)
    if rank != 0:
        finished[0] = True
        if world > 1:
            dist.destroy_process_group()
        return
    finish()
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations, measured after the headline (CUDA events on the ctx stream)
# ------------------------------------------------------------------------------------------------------
def _event_ms(torch, fe, local, fn, reps):
    ext = torch.cuda.ExternalStream(fe.stream(), device=torch.device("cuda", local))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(ext)
    for _ in range(reps):
        out = fn()
    e1.record(ext)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps, out


def bench_c3(torch, L, local, args, log):
    """configs[2]: batch-256 1920x1080 dense lane frames, detection + descriptors (+ ground / sanity)."""
    from oracle import synth
    nb, h, w = args.c3_batch, 1080, 1920
    pool = np.stack([synth.frame(s, h, w, dense=True) for s in range(8)])
    frames = torch.from_numpy(pool[np.arange(nb) % len(pool)].copy()).cuda()
    cam, Hg = L.scaled_calibration(w, h)
    fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(h, w), top_cutoff=0, camera=cam, homography=Hg, src_size=(h, w),
                    max_batch=nb, device=local, max_segments_per_frame=4096, pinned=True)
    st = L.STAGE_DETECT | L.STAGE_GROUND | L.STAGE_DESCRIBE
    b = fe.process(frames, stages=st)
    b = fe.process(frames, stages=st)
    ms, b = _event_ms(torch, fe, local, lambda: fe.process(frames, stages=st), 2)
    fe.set_chunk_frames(-1)
    fe.process(frames, stages=st)
    kern = {k: round(v, 3) for k, v in fe.timings()}
    out = {"workload": "batch-%d 1920x1080 dense synthetic lane frames: detect + ground/sanity + LBD descriptors" % nb, "batch": nb,
           "frames_per_s": nb / (ms * 1e-3), "ms_per_batch": ms, "segments_per_frame": b.n_segments / nb,
           "distinct_frames": len(pool), "input": "resident in HBM (%.2f GB > L2)" % (nb * h * w * 3 / 1e9),
           "algorithmic_GBps_21.7N_model": 21.7 * h * w * nb / (ms * 1e-3) / 1e9, "kernels_ms_single_stream": kern}
    fe.close()
    log("c3 done: %s" % out)
    return out


def bench_c4(torch, L, local, log):
    """configs[3]: line_associator Hamming kNN, 2 000 query segments vs 100 000 accumulated map lines (256-bit codes), k = 2."""
    from oracle import synth
    q, m, src = synth.descriptor_sets(2000, 100000, seed=0)
    fe = L.FrontEnd(max_batch=1, device=local)
    dq, dm = torch.from_numpy(q).cuda(), torch.from_numpy(m).cuda()
    di = torch.empty((2000, 2), dtype=torch.int32, device="cuda"); dd = torch.empty_like(di)
    call = lambda: fe.knn_device(dq.data_ptr(), 2000, dm.data_ptr(), 100000, 2, di.data_ptr(), dd.data_ptr(), max_dist=L.MATCH_RADIUS)
    for _ in range(5):
        call()
    ms, _ = _event_ms(torch, fe, local, call, 50)
    kms = dict(fe.timings()).get("knn", None)
    ok = bool((di[:, 0].cpu().numpy() == np.where(src >= 100000 - 64, src - (100000 - 64), src)).all())
    pairs = 2000 * 100000
    out = {"workload": "Hamming kNN 2000 x 100000 x 256 bit, k=2, radius 128, reference tie order", "ms_per_call": ms, "kernel_ms": kms,
           "pairs_per_s": pairs / (ms * 1e-3), "popc32_equiv_per_s": pairs * 8 / (ms * 1e-3), "planted_neighbours_found": ok,
           "queries_per_s": 2000 / (ms * 1e-3)}
    fe.close()
    log("c4 done: %s" % out)
    return out


def bench_c5(torch, dist, L, fe, dev, n, rank, world, local, args, log):
    """configs[4]: log replay in epochs, sharded over the ranks: per epoch every rank detects + describes its shard,
    the ranks all-gather kept segments + descriptors (NCCL inside liblsf.so), every rank appends them -- moved to the map frame
    with the odometry poses -- to its replica of the map, and the next epoch is matched against that snapshot.  The log is
    args.c5_frames frames long in total: each rank cycles through its resident 1 000-frame pool (generating 100 000 distinct
    frames on the host would take minutes and measure numpy, not the path)."""
    from lane_slam_b200 import odometry
    from lane_slam_b200.replay import EpochReplay
    total, E = args.c5_frames, args.c5_epoch
    per = min(E // world, n)                            # frames per rank and epoch
    E = per * world
    n_epochs = max(1, total // E)
    tt = np.arange(n_epochs * E)
    poses = odometry.integrate((tt * 0.1e9) % 1e9, 0.30 + 0.05 * np.sin(tt / 50.0), 0.30 + 0.05 * np.cos(tt / 70.0))
    fe.map_clear()
    rp = EpochReplay(fe, rank, world, epoch_frames=E, poses=poses, k=K_NN)
    ext = torch.cuda.ExternalStream(fe.stream(), device=torch.device("cuda", local))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def run(epochs):
        nseg = 0
        for e in epochs:
            lo, hi = rp.shard(e)
            off = (e * per) % max(1, n - per)
            b, mi, md = rp.run_epoch(e, dev[off:off + per], lo)
            nseg += b.n_segments
        return nseg
    run(range(0, min(2, n_epochs)))                       # warm-up epochs (part of the log, not of the timed region)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(ext)
    nseg = run(range(min(2, n_epochs), n_epochs))
    msize = rp.finish()
    e1.record(ext)
    if world > 1:
        dist.barrier()
    e1.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
    timed_frames = (n_epochs - min(2, n_epochs)) * E
    out = {"workload": "%d-frame 640x480 log replay in epochs of %d frames over %d GPU(s): detect + describe per shard, all-gather of kept "
                       "segments + descriptors, map append with odometry poses, kNN (k=2) of epoch e against the map after epoch e-1"
                       % (n_epochs * E, E, world),
           "frames": n_epochs * E, "timed_frames": timed_frames, "epoch_frames": E, "frames_per_s": timed_frames / dt if dt > 0 else None,
           "seconds": dt, "map_lines_at_end": int(msize), "segments_matched_this_rank": int(nseg),
           "note": "every rank cycles through its resident 1000-frame pool; brute-force matching against the growing map dominates"}
    log("c5 done: %s" % out)
    return out


_OUT = None


def emit(line):
    """The one JSON line goes to the process's original stdout; everything libraries print meanwhile (NCCL's version
    banner, for instance) was diverted to stderr in main()."""
    print(json.dumps(line), file=_OUT if _OUT is not None else sys.stdout, flush=True)


def main():
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=N_FRAMES, help="frames per step per GPU (default: the 1000-frame workload)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--verbose", action="store_true", help="progress on stderr")
    ap.add_argument("--parity-frames", type=int, default=64, help="frames of the timed step re-checked by the CPU oracle (0 = skip)")
    ap.add_argument("--inflight", type=int, default=2, help="contexts (batches) in flight per GPU, one host thread each")
    ap.add_argument("--extras", default="c3,c4,c5", help="other BASELINE configs measured after the headline (c3, c4: 1 GPU only)")
    ap.add_argument("--c3-batch", type=int, default=256)
    ap.add_argument("--c5-frames", type=int, default=100000, help="length of the replayed log (total over all GPUs)")
    ap.add_argument("--c5-epoch", type=int, default=1024, help="frames per epoch (total over all GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
