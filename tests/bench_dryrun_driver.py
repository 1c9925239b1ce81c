"""TEST INFRASTRUCTURE: runs bench.py's GPU arm on a machine without a GPU, with the device pieces replaced by stand-ins, so
that the CONTROL FLOW of the benchmark (passes, threads, failure handling across ranks, the one JSON line) is exercised by
the CPU test suite.  Nothing here measures anything.

    python tests/bench_dryrun_driver.py [bench.py arguments]

environment:
    DRYRUN_FAIL = "<kind>:<ncontexts>:<rank>"   the fake FrontEnd raises in every batch of that pass on that rank
    DRYRUN_C5   = "hang"                         bench_c5 is replaced: rank 1 raises, the other ranks wait forever (DRYRUN_C5_MODE =
                                                 "raise": their collective raises instead)
with WORLD_SIZE > 1 the ranks talk over gloo (RANK / MASTER_ADDR / MASTER_PORT from the environment)."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import bench
import lane_slam_b200 as L

FAIL = os.environ.get("DRYRUN_FAIL", "")
RANK = int(os.environ.get("RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))

# ---- torch: no device --------------------------------------------------------------------------------------------------------
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.cuda.empty_cache = lambda *a, **k: None
torch.Tensor.pin_memory = lambda self, *a, **k: self
torch.Tensor.cuda = lambda self, *a, **k: self


def _no_device(fn):
    def wrapped(*a, **k):
        k.pop("device", None)
        return fn(*a, **k)
    return wrapped


torch.tensor = _no_device(torch.tensor)
torch.zeros = _no_device(torch.zeros)


class _Event(object):
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


torch.cuda.Event = _Event
torch.cuda.ExternalStream = lambda *a, **k: None

_init = dist.init_process_group
dist.init_process_group = lambda *a, **k: _init("gloo", rank=RANK, world_size=WORLD)


# ---- the front end: canned results, real collectives ------------------------------------------------------------------------
class _Batch(object):
    pass


class FakeFrontEnd(object):
    current_pass = ["warmup", 0]           # set by the patched bench.timed wrapper below

    def __init__(self, cfg=None, **kw):
        self.max_batch = kw.get("max_batch", 1)
        self.world = 1

    def reset_sequence(self):
        pass

    def _batch(self, n):
        b = _Batch()
        b.n_frames, b.n_segments = n, 3 * n
        b.keep = np.ones(3 * n, np.uint8)
        b.lines_px = np.zeros((3 * n, 4), np.float32)
        b.desc = np.zeros((3 * n, 32), np.uint8)
        return b

    def _maybe_fail(self, kind):
        if FAIL:
            k, nc, r = FAIL.split(":")
            if k == kind and int(nc) == FakeFrontEnd.current_pass[1] and int(r) == RANK and FakeFrontEnd.current_pass[0] == "timed":
                raise L.LsfError(-3, "injected failure")

    def process(self, frames, stages=0, k=0):
        self._maybe_fail("dev" if isinstance(frames, torch.Tensor) else "raw")
        return self._batch(len(frames))

    def process_jpeg(self, blob, off, stages=0, k=0):
        self._maybe_fail("jpeg")
        return self._batch(len(off) - 1)

    def prefetch(self, frames):
        pass

    def timings(self):
        return [("color_canny", 1.0), ("lsd_grow", 2.0), ("gray_sobel", 0.5), ("jpeg_h2d", 0.1), ("jpeg_decode", 0.2), ("d2h", 0.1)]

    def launch_count(self):
        return 14

    def stream(self):
        return 0

    def set_chunk_frames(self, c):
        pass

    def exchange_init(self, rank=0, world=1, unique_id=None, max_records=0):
        self.world = world

    def allgather_start(self, frame_base=0):
        if self.world > 1:                      # a real blocking collective: a rank that skips one leaves the others waiting
            t = torch.ones(1)
            dist.all_reduce(t)

    def exchange_wait(self):
        return None, 7, [7]

    def map_clear(self):
        pass

    def close(self):
        pass


L.FrontEnd = FakeFrontEnd

_timed_marker = bench.threading.Barrier      # bench.timed builds a Barrier(ncontexts + 1): use it to learn which pass is running


def _barrier(parties, *a, **k):
    FakeFrontEnd.current_pass = ["timed", parties - 1]
    return _timed_marker(parties, *a, **k)


bench.threading.Barrier = _barrier

if os.environ.get("DRYRUN_C5") == "hang":
    def _c5(torch_, dist_, L_, fe, dev, n, rank, world, local, args, log):
        if rank == 1:
            raise RuntimeError("injected c5 failure on rank 1")
        if os.environ.get("DRYRUN_C5_MODE") == "raise":
            dist_.barrier()                      # gloo notices that rank 1 is gone and raises; NCCL would wait forever:
        threading.Event().wait()                 # ... like this
        return {"never": True}
    bench.bench_c5 = _c5
    bench.EXTRAS_DEADLINE_S = 5

if __name__ == "__main__":
    bench.main()
