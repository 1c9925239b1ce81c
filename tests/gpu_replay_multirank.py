"""Run under torchrun (one rank per GPU): the epoch replay over NCCL inside liblsf.so.  Every rank must end with the same
map, equal to what the oracle-backed host loop builds from the whole log in one process; matches likewise.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/gpu_replay_multirank.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import lane_slam_b200 as L
    from lane_slam_b200 import odometry
    from lane_slam_b200.replay import EpochReplay
    from host_backend import HostBackend
    from oracle import reference_glue as rg, synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_total, E, H, W = 40, 16, 240, 320
    cfg = rg.check_configuration(dict(rg.DEFAULT_DETECTOR_CONFIG))
    cam, Hg = rg.scaled_camera(W, H)
    poses = np.cumsum(np.random.default_rng(3).normal(0, 0.02, (n_total, 3)), axis=0)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(odometry.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=(H, W), top_cutoff=0, camera=cam, homography=Hg, src_size=(H, W),
                    max_batch=E, max_segments_per_frame=2048, device=local)
    fe.exchange_init(rank=rank, world=world, unique_id=bytes(uid.cpu().numpy()))
    rp = EpochReplay(fe, rank, world, epoch_frames=E, poses=poses, k=2)
    mine = []
    for e in range((n_total + E - 1) // E):
        lo, hi = rp.shard(e, n_total)
        b, mi, md = rp.run_epoch(e, synth.sequence(hi - lo, base_seed=0, H=H, W=W, start=lo), lo)
        mine.append((lo, hi, mi.copy(), md.copy()))
    rp.finish()
    m = fe.map_read()
    # reference: the whole log in one process on the host
    hb = HostBackend(cfg, (H, W), 0, cam, Hg)
    hr = EpochReplay(hb, 0, 1, epoch_frames=E, poses=poses, k=2)
    ref = []
    for e in range((n_total + E - 1) // E):
        lo, hi = hr.shard(e, n_total)
        b, mi, md = hr.run_epoch(e, synth.sequence(hi - lo, base_seed=0, H=H, W=W, start=lo), lo)
        ref.append((b.frame_offset, mi, md, lo))
    hr.finish()
    hm = hb.map_read()
    for k in ("color", "frame", "desc"):
        assert np.array_equal(m[k], hm[k]), (rank, k)
    assert np.abs(m["ground"] - hm["ground"]).max() <= 1e-4
    for e, (lo, hi, mi, md) in enumerate(mine):
        fo, ri, rd, base = ref[e]
        s0, s1 = fo[lo - base], fo[hi - base]
        assert np.array_equal(mi, ri[s0:s1]) and np.array_equal(md, rd[s0:s1]), (rank, e)
    fe.close()
    dist.barrier()
    if rank == 0:
        print("replay multirank ok: world %d, map %d lines" % (world, len(hm["desc"])))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
