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


# ---- epoch replay (SURVEY 8e / BASELINE configs[4]) over gloo, world size 2, vs one process ------------------------
def _replay(rank, world, n_total, E, distributed):
    from host_backend import HostBackend
    from lane_slam_b200.replay import EpochReplay
    from oracle import reference_glue as rg, synth
    H, W = 120, 160
    cfg = rg.check_configuration(dict(rg.DEFAULT_DETECTOR_CONFIG))
    cam, Hg = rg.scaled_camera(W, H)
    rng = np.random.default_rng(3)
    poses = np.cumsum(rng.normal(0, 0.02, (n_total, 3)), axis=0)
    be = HostBackend(cfg, (H, W), 0, cam, Hg, distributed=distributed)
    rp = EpochReplay(be, rank, world, epoch_frames=E, poses=poses, k=2)
    matches = []
    for e in range((n_total + E - 1) // E):
        lo, hi = rp.shard(e, n_total)
        frames = synth.sequence(hi - lo, base_seed=0, H=H, W=W, start=lo)
        b, mi, md = rp.run_epoch(e, frames, lo)
        matches.append((lo, hi, mi, md, b.n_segments))
    rp.finish()
    return be.map_read(), matches


def _replay_worker(rank, world, port, n_total, E, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out[rank] = _replay(rank, world, n_total, E, True)
    dist.destroy_process_group()


def test_epoch_replay_gloo_world2_equals_single_process():
    """Every rank ends with the same map, in global frame order, equal to the map one process builds from the whole log;
    the matches of epoch e (against the snapshot after e-1) are the single-process ones, shard by shard."""
    n_total, E = 14, 6          # 3 epochs, the last one ragged (2 frames: rank 0 gets 1, rank 1 gets 1)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_replay_worker, args=(2, port, n_total, E, out), nprocs=2, join=True)
    ref_map, ref_matches = _replay(0, 1, n_total, E, False)
    assert len(ref_map["desc"]) > 0
    for rank in range(2):
        m = out[rank][0]
        for k in ("ground", "color", "frame", "desc"):
            assert np.array_equal(m[k], ref_map[k]), (rank, k)
    assert (np.diff(ref_map["frame"]) >= 0).all()
    for e in range(len(ref_matches)):
        parts_i = np.concatenate([out[r][1][e][2] for r in range(2)])
        parts_d = np.concatenate([out[r][1][e][3] for r in range(2)])
        assert np.array_equal(parts_i, ref_matches[e][2]) and np.array_equal(parts_d, ref_matches[e][3])
        if e > 0:
            assert (ref_matches[e][2][:, 0] >= 0).any()           # later epochs do find map lines
    assert (ref_matches[0][2] == -1).all()                        # epoch 0 sees an empty map


def test_odometry_matches_reference_node_golden():
    """lsf_odometry_step (host arithmetic inside liblsf.so) == the reference's OdometryNode run by
    tests/golden/make_golden_odometry.py: pose after every command, bit for bit, skipped updates included."""
    from lane_slam_b200 import odometry
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "odometry.npz"))
    od = odometry.Odometry(0.0)
    for i in range(len(g["nsecs"])):
        adv = od.getPose(g["nsecs"][i], g["vel_left"][i], g["vel_right"][i])
        assert adv == bool(g["advanced"][i]) and tuple(g["poses"][i]) == od.pose(), i
    assert len(od.trajectory) == int(g["advanced"].sum())
    assert np.array_equal(odometry.integrate(g["nsecs"], g["vel_left"], g["vel_right"]), g["poses"])
