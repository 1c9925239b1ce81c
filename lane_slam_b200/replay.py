"""Log replay in epochs over several GPUs (BASELINE.json configs[4]).

Frames shard by rank inside every epoch of E frames (contiguous ranges, so that concatenating the ranks' results in rank
order gives global frame order).  The map is a snapshot replicated on every rank and updated at epoch boundaries:

    epoch e on rank r:   detect + ground projection + sanity + descriptors of the rank's frames   (no collective)
                         finish the exchange of epoch e-1, append its gathered kept lines -- moved to the map frame with
                         the odometry poses of their frames -- to the map, in global frame order     (same on every rank)
                         match the descriptors of epoch e against that snapshot                      (k-NN, no collective)
                         start the exchange of epoch e (one all-gather on a side stream; it overlaps epoch e+1's kernels)

The reference defines no association semantics (line_associator is a stub); this is the contract SURVEY.md 8e fixes:
frames of epoch e see the map after epoch e-1.  With one rank and E = 1 it degenerates to matching every frame against
everything seen before it.  `backend` is a FrontEnd (GPU, NCCL inside liblsf.so) -- or any object with the same six
methods (tests run the same loop over gloo with a host backend)."""
import numpy as np

from ._lib import STAGE_DESCRIBE, STAGE_DETECT, STAGE_GROUND
from .dist import shard_range


class EpochReplay(object):
    def __init__(self, backend, rank=0, world=1, epoch_frames=256, poses=None, k=2):
        self.be, self.rank, self.world, self.E, self.k = backend, int(rank), int(world), int(epoch_frames), int(k)
        self.poses = None if poses is None else np.ascontiguousarray(poses, np.float64).reshape(-1, 3)
        self.pending = []          # epochs whose exchange is in flight
        self.kept_total = 0

    def shard(self, epoch, n_frames_total=None):
        """Global frame range [lo, hi) of this rank in `epoch`."""
        base = epoch * self.E
        n = self.E if n_frames_total is None else max(0, min(self.E, n_frames_total - base))
        lo, hi = shard_range(n, self.rank, self.world)
        return base + lo, base + hi

    def _absorb(self):
        """Finish the oldest exchange and append its records to the map (global frame order)."""
        epoch = self.pending.pop(0)
        rec, n, counts = self.be.exchange_wait()
        base = epoch * self.E
        p = None if self.poses is None else self.poses[base:base + self.E]
        self.be.map_append_records(rec, n, poses=p, pose_frame_base=base)
        self.kept_total += n
        return n, counts

    def run_epoch(self, epoch, frames, frame_lo):
        """frames: this rank's frames of `epoch` (global ids frame_lo ..).  Returns (SegmentBatch, match_idx, match_dist):
        the matches refer to map rows (global, identical on every rank)."""
        b = self.be.process(frames, stages=STAGE_DETECT | STAGE_GROUND | STAGE_DESCRIBE)
        while self.pending:                       # the map must hold everything up to epoch - 1
            self._absorb()
        midx, mdist = self.be.match_batch(b.n_segments, self.k)
        self.be.allgather_start(frame_base=frame_lo)
        self.pending.append(epoch)
        return b, midx, mdist

    def finish(self):
        while self.pending:
            self._absorb()
        return self.be.map_size()
