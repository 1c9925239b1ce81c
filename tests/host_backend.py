"""TEST INFRASTRUCTURE: a CPU stand-in with the six methods EpochReplay needs from a FrontEnd, built on the oracle
(frames -> kept records), torch.distributed/gloo (the all-gather) and numpy (the map).  It is the checker for the GPU
replay and what the world-size-2 gloo tests drive."""
import numpy as np

from lane_slam_b200 import dist as ldist
from oracle import cmodel as cm, reference_glue as rg


class _Batch(object):
    pass


class HostBackend(object):
    def __init__(self, cfg, img_size, top_cutoff, camera, homography, distributed=False):
        self.cfg, self.isz, self.cut, self.cam, self.Hg, self.distributed = cfg, img_size, top_cutoff, camera, homography, distributed
        self.map = dict(ground=np.zeros((0, 4)), color=np.zeros(0, np.uint8), frame=np.zeros(0, np.int32), desc=np.zeros((0, 32), np.uint8))
        self._gathered = []
        self._last = None

    def process(self, frames, stages=0):
        outs = [cm.front_end_frame(f, self.cfg, self.isz, self.cut, self.cam, self.Hg, descriptors=True) for f in frames]
        b = _Batch()
        b.n_frames = len(frames)
        per = [len(o["lines_px"]) for o in outs]
        b.frame_offset = np.concatenate([[0], np.cumsum(per)]).astype(np.int32)
        b.n_segments = int(b.frame_offset[-1])
        cat = lambda k, shape, dt: np.concatenate([o[k] for o in outs]) if outs else np.zeros(shape, dt)
        b.color = cat("color", (0,), np.uint8); b.ground = cat("ground", (0, 4), np.float64)
        b.keep = cat("keep", (0,), bool).astype(np.uint8); b.desc = cat("desc32", (0, 32), np.uint8)
        b.lines_px = cat("lines_px", (0, 4), np.float32)
        self._last = b
        return b

    def allgather_start(self, frame_base=0):
        rec = ldist.pack_kept(self._last, frame_base)
        if self.distributed:
            import torch.distributed as dist
            counts = [None] * dist.get_world_size()
            dist.all_gather_object(counts, len(rec))
            rec = ldist.allgather_records(rec)
        else:
            counts = [len(rec)]
        self._gathered.append((rec, counts))

    def exchange_wait(self):
        rec, counts = self._gathered.pop(0)
        return rec, len(rec), counts

    def map_append_records(self, rec, n, poses=None, pose_frame_base=0):
        u = ldist.unpack(rec[:n])
        g = rg.map_transform(u["ground"], u["frame"], poses, pose_frame_base)
        m = self.map
        m["ground"] = np.concatenate([m["ground"], g]); m["color"] = np.concatenate([m["color"], u["color"]])
        m["frame"] = np.concatenate([m["frame"], u["frame"]]); m["desc"] = np.concatenate([m["desc"], u["desc"]])

    def match_batch(self, n_segments, k=2):
        if len(self.map["desc"]) == 0 or n_segments == 0:
            return np.full((n_segments, k), -1, np.int32), np.full((n_segments, k), -1, np.int32)
        return cm.knn_mihasher(self._last.desc, self.map["desc"], k)

    def map_size(self):
        return len(self.map["desc"])

    def map_read(self):
        return self.map
