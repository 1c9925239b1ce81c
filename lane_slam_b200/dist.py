"""Multi-GPU plumbing: frames shard by rank (no collective on the data path); the one exchange step of the
path is the all-gather of kept ground segments + descriptors ("map lines") at an epoch boundary
(SURVEY.md 8e).  torch.distributed (NCCL on GPUs, gloo on CPU for tests) is only the transport.

Record layout (72 bytes, little endian): frame_id i32 | color u8 | pad[3] | ground x1,y1,x2,y2 f64 | desc u8[32]
"""
import numpy as np

RECORD_BYTES = 72
_REC = np.dtype([("frame", "<i4"), ("color", "u1"), ("pad", "u1", (3,)), ("ground", "<f8", (4,)), ("desc", "u1", (32,))])
assert _REC.itemsize == RECORD_BYTES


def shard_range(n_frames, rank, world):
    """Contiguous frame range [lo, hi) of `rank` (keeps temporal locality for the map)."""
    lo = (n_frames * rank) // world
    hi = (n_frames * (rank + 1)) // world
    return lo, hi


def pack_kept(batch, frame_base=0):
    """Kept segments of a SegmentBatch -> structured array of 72-byte records (global frame ids)."""
    keep = batch.keep.astype(bool)
    n = int(keep.sum())
    rec = np.zeros(n, _REC)
    if n:
        frame_of = np.repeat(np.arange(batch.n_frames, dtype=np.int32), np.diff(batch.frame_offset))
        rec["frame"] = frame_of[keep] + frame_base
        rec["color"] = batch.color[keep]
        rec["ground"] = batch.ground[keep]
        rec["desc"] = batch.desc[keep]
    return rec


def unpack(records):
    r = records.view(_REC) if records.dtype != _REC else records
    return dict(frame=r["frame"].copy(), color=r["color"].copy(), ground=r["ground"].copy(), desc=r["desc"].copy())


def allgather_records(rec, device=None):
    """All-gather variable-length record arrays: counts first, then one padded payload.  Returns the
    concatenation over ranks in rank order (= global frame order for contiguous shards)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    cnt = torch.tensor([len(rec)], dtype=torch.int64, device=dev)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, cnt)
    counts = counts.cpu().numpy()
    mx = int(counts.max())
    if mx == 0:
        return np.zeros(0, _REC)
    pad = np.zeros((mx, RECORD_BYTES), np.uint8)
    pad[:len(rec)] = rec.view(np.uint8).reshape(-1, RECORD_BYTES)
    send = torch.from_numpy(pad).to(dev)
    recv = torch.empty((world, mx, RECORD_BYTES), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1))
    recv = recv.cpu().numpy()
    parts = [recv[r, :counts[r]].reshape(-1).view(_REC) for r in range(world)]
    return np.concatenate(parts)


def allgather_kept_segments(batch, frame_base=0, device=None):
    """pack + all-gather: every rank ends up with every rank's kept segments (the shared map snapshot)."""
    return allgather_records(pack_kept(batch, frame_base), device=device)


class _DeviceBytes(object):
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def allgather_kept_device(fe, frame_base=0, device=None):
    """The exchange step on the device (NCCL): lsf_pack_kept_records builds this rank's 72-byte records in HBM,
    the counts and then one padded payload are all-gathered.  Returns (records uint8 [total, 72] CUDA tensor in rank
    order, per-rank counts).  Nothing passes through host memory except the counts."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    ptr, n = fe.pack_kept_device(frame_base)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([n], dtype=torch.int64, device=dev))
    counts = counts.cpu().tolist()
    mx = max(counts)
    if mx == 0:
        return torch.zeros((0, RECORD_BYTES), dtype=torch.uint8, device=dev), counts
    send = torch.zeros((mx, RECORD_BYTES), dtype=torch.uint8, device=dev)
    if n:
        send.view(-1)[:n * RECORD_BYTES].copy_(torch.as_tensor(_DeviceBytes(ptr, n * RECORD_BYTES), device=dev))
    recv = torch.empty((world, mx, RECORD_BYTES), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv.view(-1), send.view(-1))
    return torch.cat([recv[r, :counts[r]] for r in range(world)]), counts
