"""CPU tests of the N>1 plumbing with gloo, world_size 2: frame sharding and the all-gather of kept segments."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from lane_slam_b200 import dist as ldist
from lane_slam_b200.frontend import SegmentBatch


def _fake_batch(rank, n_frames):
    rng = np.random.default_rng(100 + rank)
    counts = rng.integers(0, 5, (n_frames, 3)).astype(np.int32)
    per = counts.sum(axis=1)
    S = int(per.sum())
    fo = np.concatenate([[0], np.cumsum(per)]).astype(np.int32)
    arrays = dict(counts=counts, frame_offset=fo, color=rng.integers(0, 3, S).astype(np.uint8),
                  lines_px=np.zeros((S, 4), np.float32), normals=np.zeros((S, 2)), centers=np.zeros((S, 2), np.float32),
                  pixels_normalized=np.zeros((S, 4), np.float32), normal_f32=np.zeros((S, 2), np.float32),
                  ground=rng.normal(size=(S, 4)), keep=(rng.random(S) < 0.6).astype(np.uint8),
                  desc=rng.integers(0, 256, (S, 32), dtype=np.uint8), match_idx=None, match_dist=None)
    return SegmentBatch(n_frames, S, arrays, 0)


def _worker(rank, world, port, n_total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = ldist.shard_range(n_total, rank, world)
    b = _fake_batch(rank, hi - lo)
    rec = ldist.allgather_kept_segments(b, frame_base=lo)
    out[rank] = (lo, hi, ldist.unpack(rec))
    dist.destroy_process_group()


def test_shard_ranges_cover_and_are_contiguous():
    for n, w in [(1000, 8), (7, 2), (100000, 4), (3, 8)]:
        r = [ldist.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))


def test_allgather_kept_segments_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(2, port, 9, out), nprocs=2, join=True)
    a, b = out[0][2], out[1][2]
    for k in ("frame", "color", "ground", "desc"):
        assert np.array_equal(a[k], b[k])                 # every rank holds the same gathered map
    # content = rank 0's kept segments followed by rank 1's, with global frame ids
    exp = []
    for rank in range(2):
        lo, hi = ldist.shard_range(9, rank, 2)
        exp.append(ldist.unpack(ldist.pack_kept(_fake_batch(rank, hi - lo), frame_base=lo)))
    for k in ("frame", "color", "ground", "desc"):
        assert np.array_equal(a[k], np.concatenate([exp[0][k], exp[1][k]]))
    assert (np.diff(a["frame"]) >= 0).all() and a["frame"].max() < 9
