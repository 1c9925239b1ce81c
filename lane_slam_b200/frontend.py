"""Host side of the B200 line front end: the node-level batch API over the C ABI (include/lsf.h).

Mirrors, for a batch of frames, what the reference's ROS nodes do per frame:
  LineDetectorNode.processImage_      src/line_detector/src/line_detector_node.py:141-213
  GroundProjectionNode.lineseglist_cb src/ground_projection/src/ground_projection_node.py:55-65
  LineSanityNode.processSegmentList   src/line_sanity/src/line_sanity_node.py:48-72
  (line_associator stub)              BinaryDescriptor::compute + BinaryDescriptorMatcher::knnMatch
All arithmetic happens in liblsf.so on the GPU; this module only owns buffers and argument checking.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (LsfConfig, LsfError, LsfSegments, MATCH_RADIUS, MEM_DEVICE, MEM_HOST, STAGE_DESCRIBE, STAGE_DETECT,
                   STAGE_GROUND, STAGE_MATCH, STAGE_MATCH_PREV, TAP, TIES_INDEX, TIES_REFERENCE)

WHITE, YELLOW, RED = 0, 1, 2          # src/duckietown_msgs/msg/Segment.msg:1-3
COLORS = ("white", "yellow", "red")

# exact key set demanded by Configurable (duckietown_utils/parameters.py:15-23, line_detector_lsd.py:20-36)
DETECTOR_PARAM_NAMES = ['hsv_white1', 'hsv_white2', 'hsv_yellow1', 'hsv_yellow2', 'hsv_red1', 'hsv_red2',
                        'hsv_red3', 'hsv_red4', 'dilation_kernel_size', 'canny_thresholds', 'hough_threshold',
                        'hough_min_line_length', 'hough_max_line_gap']

# src/duckietown/config/baseline/line_detector/line_detector_node/default.yaml:7-23
DEFAULT_DETECTOR_CONFIGURATION = dict(
    dilation_kernel_size=3, canny_thresholds=[80, 200], hough_threshold=2, hough_min_line_length=3,
    hough_max_line_gap=1, hsv_white1=[0, 0, 150], hsv_white2=[180, 60, 255], hsv_yellow1=[25, 140, 100],
    hsv_yellow2=[45, 255, 255], hsv_red1=[0, 140, 100], hsv_red2=[15, 255, 255], hsv_red3=[165, 140, 100],
    hsv_red4=[180, 255, 255])


def check_detector_configuration(configuration):
    """Same contract as Configurable.__init__ (parameters.py:2-36): dict with exactly the known keys."""
    if not isinstance(configuration, dict):
        raise ValueError('Expecting a dict, obtained %r' % (configuration,))
    extra = set(configuration) - set(DETECTOR_PARAM_NAMES)
    missing = set(DETECTOR_PARAM_NAMES) - set(configuration)
    if extra or missing:
        raise ValueError('Error while loading configuration.\nExtra parameters: %r\nMissing parameters: %r\n'
                         % (extra, missing))
    return dict(configuration)


def scaled_calibration(W, H):
    """The reference's default calibration (camera_intrinsic/default.yaml, camera_extrinsic/default.yaml,
    640x480) scaled to a W x H camera."""
    cfg = LsfConfig()
    _lib.load().lsf_default_config(C.byref(cfg))
    sx, sy = W / 640.0, H / 480.0
    K = np.array(cfg.K[:], np.float64).reshape(3, 3)
    P = np.array(cfg.P[:], np.float64).reshape(3, 4)
    K[0] *= sx; K[1] *= sy; P[0] *= sx; P[1] *= sy
    Hg = np.array(cfg.Hgnd[:], np.float64).reshape(3, 3) @ np.diag([1 / sx, 1 / sy, 1.0])
    camera = dict(width=W, height=H, K=K.ravel().tolist(), D=list(cfg.D[:]), R=list(cfg.R[:]), P=P.ravel().tolist())
    return camera, Hg.ravel().tolist()


class SegmentBatch:
    """SoA result of one batch (views into the FrontEnd's reusable host buffers; copy to keep)."""

    def __init__(self, n_frames, n_segments, arrays, k):
        self.n_frames, self.n_segments, self.k = n_frames, n_segments, k
        S = n_segments
        self.counts = arrays["counts"][:n_frames]
        self.frame_offset = arrays["frame_offset"][:n_frames + 1]
        for name in ("color", "lines_px", "normals", "centers", "pixels_normalized", "normal_f32", "ground", "keep",
                     "desc"):
            setattr(self, name, arrays[name][:S])
        self.match_idx = arrays["match_idx"][:S, :k] if k else None
        self.match_dist = arrays["match_dist"][:S, :k] if k else None

    def frame_slice(self, f):
        return slice(int(self.frame_offset[f]), int(self.frame_offset[f + 1]))

    def frame(self, f):
        s = self.frame_slice(f)
        return dict(color=self.color[s], lines_px=self.lines_px[s], normals=self.normals[s], centers=self.centers[s],
                    pixels_normalized=self.pixels_normalized[s], normal=self.normal_f32[s], ground=self.ground[s],
                    keep=self.keep[s].astype(bool), desc=self.desc[s], counts=self.counts[f].tolist())


class FrontEnd:
    """One lsf_ctx: fixed img_size / top_cutoff / detector configuration / calibration, batches of frames."""

    def __init__(self, configuration=None, img_size=(120, 160), top_cutoff=40, camera=None, homography=None,
                 src_size=(480, 640), max_batch=1, device=0, ai_scale=(1, 1, 1), ai_shift=(0, 0, 0),
                 max_segments_per_color=0, max_pixels_per_color=0, max_segments_per_frame=1024, pinned=False,
                 chunk_frames=0, tie_order=_lib.TIES_REFERENCE, grow_warps_per_sm=0):
        self._lib = _lib.load()
        conf = check_detector_configuration(configuration if configuration is not None
                                            else DEFAULT_DETECTOR_CONFIGURATION)
        cfg = LsfConfig()
        self._lib.lsf_default_config(C.byref(cfg))
        cfg.img_h, cfg.img_w, cfg.top_cutoff = int(img_size[0]), int(img_size[1]), int(top_cutoff)
        los = [conf['hsv_white1'], conf['hsv_yellow1'], conf['hsv_red1'], conf['hsv_red3']]
        his = [conf['hsv_white2'], conf['hsv_yellow2'], conf['hsv_red2'], conf['hsv_red4']]
        for i in range(4):
            for j in range(3):
                cfg.hsv_lo[i][j] = int(los[i][j])
                cfg.hsv_hi[i][j] = int(his[i][j])
        cfg.dilation_kernel_size = int(conf['dilation_kernel_size'])
        cfg.canny_lo, cfg.canny_hi = int(conf['canny_thresholds'][0]), int(conf['canny_thresholds'][1])
        for i in range(3):
            cfg.ai_scale[i] = float(ai_scale[i]); cfg.ai_shift[i] = float(ai_shift[i])
        if camera is not None:
            for name, n in (("K", 9), ("D", 5), ("R", 9), ("P", 12)):
                arr = getattr(cfg, name)
                vals = np.asarray(camera[name], np.float64).ravel()
                if len(vals) != n:
                    raise ValueError("camera[%s] must have %d entries" % (name, n))
                for i in range(n):
                    arr[i] = float(vals[i])
            cfg.cam_w, cfg.cam_h = int(camera["width"]), int(camera["height"])
        if homography is not None:
            hv = np.asarray(homography, np.float64).ravel()
            for i in range(9):
                cfg.Hgnd[i] = float(hv[i])
        cfg.max_batch = int(max_batch)
        cfg.max_src_h, cfg.max_src_w = int(src_size[0]), int(src_size[1])
        cfg.max_segments_per_color = int(max_segments_per_color)
        cfg.max_pixels_per_color = int(max_pixels_per_color)
        cfg.device = int(device)
        cfg.chunk_frames = int(chunk_frames)   # 0 auto, < 0 one stream (per-kernel timings), > 0 frames per pipeline chunk
        cfg.grow_warps_per_sm = int(grow_warps_per_sm)   # 0 = default; fewer when several contexts share the GPU
        cfg.tie_order = int(tie_order)         # order of equal-distance neighbours: the reference's (Mihasher) or ascending index
        self.cfg = cfg
        self.img_size, self.top_cutoff, self.max_batch = tuple(img_size), int(top_cutoff), int(max_batch)
        self._ctx = C.c_void_p()
        rc = self._lib.lsf_create(C.byref(cfg), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.lsf_last_error(None).decode()
            if rc == _lib.LSF_E_CONFIG:
                raise ValueError(msg)       # reference: bad configuration -> ValueError
            raise LsfError(rc, msg)
        h, w, sh, sw = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._lib.lsf_image_dims(self._ctx, C.byref(h), C.byref(w), C.byref(sh), C.byref(sw))
        self.h, self.w, self.lsd_h, self.lsd_w = h.value, w.value, sh.value, sw.value
        segcap, pixcap, outcap = C.c_int(), C.c_int(), C.c_int()
        self._lib.lsf_capacities(self._ctx, C.byref(segcap), C.byref(pixcap), C.byref(outcap))
        self.max_segments_per_color, self.max_pixels_per_color, self.max_output_rows = segcap.value, pixcap.value, outcap.value
        # host output rows: a starting size; process() grows the buffers up to the library's own row capacity and
        # retries when a batch holds more (the reference's detector returns however many lines it finds)
        self._cap = max(1, min(int(max_segments_per_frame) * self.max_batch, self.max_output_rows))
        self._pinned = pinned
        self._host = None
        self._dev = None

    # -- buffers ------------------------------------------------------------------------------------------
    def _alloc(self, shape, dtype):
        if self._pinned:
            import torch
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name), pin_memory=True)
            self._keep.append(t)
            return t.numpy()
        return np.empty(shape, dtype)

    def _host_arrays(self):
        if self._host is None:
            self._keep = []
            S, n = self._cap, self.max_batch
            a = dict(counts=self._alloc((n, 3), np.int32), frame_offset=self._alloc((n + 1,), np.int32),
                     color=self._alloc((S,), np.uint8), lines_px=self._alloc((S, 4), np.float32),
                     normals=self._alloc((S, 2), np.float64), centers=self._alloc((S, 2), np.float32),
                     pixels_normalized=self._alloc((S, 4), np.float32), normal_f32=self._alloc((S, 2), np.float32),
                     ground=self._alloc((S, 4), np.float64), keep=self._alloc((S,), np.uint8),
                     desc=self._alloc((S, 32), np.uint8), match_idx=self._alloc((S, 8), np.int32),
                     match_dist=self._alloc((S, 8), np.int32))
            self._host = a
        return self._host

    def _check(self, rc):
        if rc != 0:
            raise LsfError(rc, self._lib.lsf_last_error(self._ctx).decode())

    # -- the batch call -----------------------------------------------------------------------------------
    def process(self, frames, stages=STAGE_DETECT | STAGE_GROUND, k=0):
        """frames: uint8 [n,H,W,3] (or [H,W,3]) BGR: a C-contiguous numpy array (host) or a CUDA torch tensor.
        Returns a SegmentBatch (host numpy views)."""
        dev_ptr = None
        if isinstance(frames, np.ndarray):
            if frames.dtype != np.uint8:
                raise ValueError("frames must be uint8")
            if frames.ndim == 3:
                frames = frames[None]
            frames = np.ascontiguousarray(frames)
            n, H, W, ch = frames.shape
            ptr, kind = frames.ctypes.data, MEM_HOST
        else:  # torch tensor on the device
            if frames.dim() == 3:
                frames = frames[None]
            if not frames.is_cuda or str(frames.dtype) != "torch.uint8" or not frames.is_contiguous():
                raise ValueError("device frames must be a contiguous CUDA uint8 tensor")
            n, H, W, ch = frames.shape
            ptr, kind = frames.data_ptr(), MEM_DEVICE
            dev_ptr = frames
        if ch != 3:
            raise ValueError("frames must be BGR, 3 channels")
        a = self._host_arrays()
        seg = LsfSegments()
        seg.mem, seg.capacity = MEM_HOST, self._cap
        for name in ("counts", "frame_offset", "color", "lines_px", "normals", "centers", "pixels_normalized",
                     "normal_f32", "ground", "keep", "desc", "match_idx", "match_dist"):
            setattr(seg, name, a[name].ctypes.data)
        kk = int(k) if (stages & (STAGE_MATCH | STAGE_MATCH_PREV)) else 0
        rc = self._lib.lsf_front_end_batch(self._ctx, ptr, n, H, W, W * 3, kind, int(stages), kk, C.byref(seg))
        if rc == _lib.LSF_E_CAPACITY and seg.n_segments > self._cap and self._cap < self.max_output_rows:
            # host buffers too small for this batch: grow them (the library reports the row count it needs) and run again;
            # the carry of LSF_STAGE_MATCH_PREV is only advanced by a successful batch, so the retry matches the same frames
            self._cap = min(self.max_output_rows, int(seg.n_segments * 1.25) + 64)
            self._host = None
            return self.process(frames, stages=stages, k=k)
        self._check(rc)
        self._last_frames = dev_ptr      # device input stays alive for tap("image") until the next call
        self._last_n = n
        S = seg.n_segments
        if kk and kk != 8:
            # library wrote [S][k] densely into buffers shaped [cap][8]: re-view
            mi = a["match_idx"].reshape(-1)[:S * kk].reshape(S, kk)
            md = a["match_dist"].reshape(-1)[:S * kk].reshape(S, kk)
            arrays = dict(a, match_idx=mi, match_dist=md)
        else:
            arrays = a
        return SegmentBatch(n, S, arrays, kk)

    def process_jpeg(self, blob, offsets, stages=STAGE_DETECT | STAGE_GROUND, k=0):
        """Frames that are still JPEG files (sensor_msgs/CompressedImage.data): blob = uint8 array holding the files back to
        back (pin it for an asynchronous copy), offsets = int64 [n+1].  Decoded on the GPU exactly like cv2.imdecode (jpg.py:21-31)."""
        blob = np.ascontiguousarray(blob, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = len(offsets) - 1
        while True:
            a = self._host_arrays()
            seg = LsfSegments()
            seg.mem, seg.capacity = MEM_HOST, self._cap
            for name in ("counts", "frame_offset", "color", "lines_px", "normals", "centers", "pixels_normalized",
                         "normal_f32", "ground", "keep", "desc", "match_idx", "match_dist"):
                setattr(seg, name, a[name].ctypes.data)
            kk = int(k) if (stages & (STAGE_MATCH | STAGE_MATCH_PREV)) else 0
            rc = self._lib.lsf_front_end_batch_jpeg(self._ctx, blob.ctypes.data, offsets.ctypes.data, n, int(stages), kk, C.byref(seg))
            if rc == _lib.LSF_E_CAPACITY and seg.n_segments > self._cap and self._cap < self.max_output_rows:
                self._cap = min(self.max_output_rows, int(seg.n_segments * 1.25) + 64)
                self._host = None
                continue
            self._check(rc)
            break
        self._last_frames = None
        self._last_n = n
        S = seg.n_segments
        if kk and kk != 8:
            arrays = dict(a, match_idx=a["match_idx"].reshape(-1)[:S * kk].reshape(S, kk), match_dist=a["match_dist"].reshape(-1)[:S * kk].reshape(S, kk))
        else:
            arrays = a
        return SegmentBatch(n, S, arrays, kk)

    def hough_lines(self, hough_threshold, hough_min_line_length, hough_max_line_gap, ground=True):
        """LineDetectorHSV.detectLines (line_detector1.py:121-125) for the three colours of every frame of the LAST process() /
        process_jpeg() call, from that batch's colour masks and edges on the device: cv2.HoughLinesP, normals, endpoint order,
        wire fields and -- ground=True -- ground projection + line sanity (lsf_hough_batch).  Returns a SegmentBatch of its own
        arrays (lines_px holds the integer endpoints; no descriptors, no matches)."""
        n = self._last_n
        cap = max(1024, 512 * n)
        while True:
            a = dict(counts=np.empty((n, 3), np.int32), frame_offset=np.empty((n + 1,), np.int32), color=np.empty((cap,), np.uint8),
                     lines_px=np.empty((cap, 4), np.float32), normals=np.empty((cap, 2), np.float64), centers=np.empty((cap, 2), np.float32),
                     pixels_normalized=np.empty((cap, 4), np.float32), normal_f32=np.empty((cap, 2), np.float32),
                     ground=np.zeros((cap, 4), np.float64), keep=np.zeros((cap,), np.uint8), desc=np.zeros((0, 32), np.uint8))
            seg = LsfSegments()
            seg.mem, seg.capacity = MEM_HOST, cap
            for name in ("counts", "frame_offset", "color", "lines_px", "normals", "centers", "pixels_normalized", "normal_f32", "ground", "keep"):
                setattr(seg, name, a[name].ctypes.data)
            rc = self._lib.lsf_hough_batch(self._ctx, int(hough_threshold), int(hough_min_line_length), int(hough_max_line_gap),
                                           1 if ground else 0, C.byref(seg))
            if rc == _lib.LSF_E_CAPACITY and seg.n_segments > cap:
                cap = int(seg.n_segments * 1.25) + 64
                continue
            self._check(rc)
            break
        a["desc"] = np.zeros((seg.n_segments, 32), np.uint8)
        return SegmentBatch(n, seg.n_segments, a, 0)

    def prefetch(self, frames):
        """Streaming replay: start copying the NEXT batch of host frames (numpy uint8 [n,H,W,3], ideally pinned) to the
        device now; a later process() call with the same array consumes the staged copy (lsf_prefetch_batch)."""
        if not isinstance(frames, np.ndarray) or frames.dtype != np.uint8 or frames.ndim != 4 or frames.shape[3] != 3 \
                or not frames.flags.c_contiguous:
            raise ValueError("prefetch expects a C-contiguous uint8 [n,H,W,3] numpy array")
        n, H, W, _ = frames.shape
        self._check(self._lib.lsf_prefetch_batch(self._ctx, frames.ctypes.data, n, H, W, W * 3))

    # -- pieces of the path on their own -----------------------------------------------------------------------
    def set_color_transform(self, scale, shift):
        """AntiInstagramTransform update (line_detector_node.py:112-114); takes effect at the next batch."""
        sc = (C.c_float * 3)(*[float(x) for x in scale]); sf = (C.c_float * 3)(*[float(x) for x in shift])
        self._check(self._lib.lsf_set_color_transform(self._ctx, sc, sf))

    def set_tie_order(self, tie_order):
        """Order of equal-distance neighbours: TIES_REFERENCE (the reference's Mihasher order) or TIES_INDEX."""
        self._check(self._lib.lsf_set_tie_order(self._ctx, int(tie_order)))

    def cancel_prefetch(self):
        """Drop batches staged by prefetch() that will not be consumed."""
        self._check(self._lib.lsf_cancel_prefetch(self._ctx))

    def set_chunk_frames(self, chunk_frames):
        """Pipeline chunking of the next batches: 0 automatic, < 0 one stream (timings() then lists every kernel)."""
        self._check(self._lib.lsf_set_chunk_frames(self._ctx, int(chunk_frames)))

    def project_filter(self, pixels_normalized, color):
        """ground_projection + line_sanity over S segments -> (ground f64 [S,4], keep bool [S])."""
        p = np.ascontiguousarray(pixels_normalized, np.float32).reshape(-1, 4)
        c = np.ascontiguousarray(color, np.uint8).reshape(-1)
        g = np.empty((len(p), 4), np.float64); kp = np.empty(len(p), np.uint8)
        self._check(self._lib.lsf_project_filter_batch(self._ctx, p.ctypes.data, c.ctypes.data, len(p), MEM_HOST,
                                                       g.ctypes.data, kp.ctypes.data))
        return g, kp.astype(bool)

    def describe(self, lines_px, frame_offset):
        """LBD descriptors of caller-supplied segments on the frames of the last process() call."""
        lp = np.ascontiguousarray(lines_px, np.float32).reshape(-1, 4)
        fo = np.ascontiguousarray(frame_offset, np.int32)
        desc = np.empty((len(lp), 32), np.uint8)
        seg = LsfSegments()
        seg.mem, seg.capacity, seg.n_frames, seg.n_segments = MEM_HOST, len(lp), len(fo) - 1, len(lp)
        seg.lines_px, seg.frame_offset, seg.desc = lp.ctypes.data, fo.ctypes.data, desc.ctypes.data
        self._check(self._lib.lsf_describe_batch(self._ctx, C.byref(seg)))
        return desc

    def knn(self, query, train, k=1, max_dist=256):
        """Exact Hamming kNN of 32-byte codes -> (idx i32 [Q,k], dist i32 [Q,k]); ties -> smallest train index."""
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        m = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        idx = np.empty((len(q), k), np.int32); dist = np.empty((len(q), k), np.int32)
        self._check(self._lib.lsf_knn_hamming(self._ctx, q.ctypes.data, len(q), m.ctypes.data, len(m), int(k),
                                              int(max_dist), MEM_HOST, idx.ctypes.data, dist.ctypes.data))
        return idx, dist

    def knn_device(self, q_ptr, nq, m_ptr, nm, k, idx_ptr, dist_ptr, max_dist=256):
        """Device-resident variant (raw device pointers)."""
        self._check(self._lib.lsf_knn_hamming(self._ctx, q_ptr, nq, m_ptr, nm, int(k), int(max_dist), MEM_DEVICE,
                                              idx_ptr, dist_ptr))

    def pack_kept_device(self, frame_base=0):
        """Kept segments of the last batch as 72-byte exchange records in a ctx-owned DEVICE buffer -> (ptr, count)."""
        ptr, cnt = C.c_void_p(), C.c_int()
        self._check(self._lib.lsf_pack_kept_records(self._ctx, int(frame_base), C.byref(ptr), C.byref(cnt)))
        return ptr.value, cnt.value

    def map_add_device(self, ptr, n):
        """Append n 32-byte descriptors that already live on the device."""
        self._check(self._lib.lsf_map_add(self._ctx, ptr, int(n), MEM_DEVICE))

    # -- map of accumulated lines / epoch exchange ------------------------------------------------------------
    def map_append(self, poses=None, frame_base=0):
        """Append the kept ground segments of the last batch to the device map, moved to the map frame with the per-frame
        poses [n_frames, 3] = (x, y, theta) (show_map.py:28-43 + the map->duck TF of odometry.py:110-120)."""
        p = None if poses is None else np.ascontiguousarray(poses, np.float64).reshape(self._last_n, 3)
        self._check(self._lib.lsf_map_append(self._ctx, None if p is None else p.ctypes.data, int(frame_base)))

    def map_append_records(self, records, n=None, poses=None, pose_frame_base=0):
        """Append 72-byte exchange records: a numpy structured / uint8 array (host) or a raw device pointer with n."""
        p = None if poses is None else np.ascontiguousarray(poses, np.float64).reshape(-1, 3)
        if isinstance(records, np.ndarray):
            r = np.ascontiguousarray(records).view(np.uint8).reshape(-1, 72)
            ptr, n, kind = r.ctypes.data, len(r), MEM_HOST
        else:
            ptr, kind = int(records), MEM_DEVICE
        self._check(self._lib.lsf_map_append_records(self._ctx, ptr, int(n), kind, None if p is None else p.ctypes.data,
                                                     int(pose_frame_base), 0 if p is None else len(p)))

    def map_read(self, first=0, count=None):
        """-> dict(ground f64 [M,4] in the map frame, color u8 [M], frame i32 [M], desc u8 [M,32])."""
        count = self.map_size() - first if count is None else count
        g = np.empty((count, 4), np.float64); c = np.empty(count, np.uint8); f = np.empty(count, np.int32)
        d = np.empty((count, 32), np.uint8)
        self._check(self._lib.lsf_map_read(self._ctx, int(first), int(count), g.ctypes.data, c.ctypes.data, f.ctypes.data, d.ctypes.data))
        return dict(ground=g, color=c, frame=f, desc=d)

    def match_batch(self, n_segments, k=2):
        """kNN of the last batch's descriptors against the map as it is now -> (idx, dist) i32 [S, k]."""
        idx = np.empty((n_segments, k), np.int32); dist = np.empty((n_segments, k), np.int32)
        self._check(self._lib.lsf_match_batch(self._ctx, int(k), MEM_HOST, idx.ctypes.data, dist.ctypes.data))
        return idx, dist

    def exchange_init(self, rank=0, world=1, unique_id=None, max_records=0):
        """One NCCL communicator per ctx for the epoch exchange (world = 1 needs none).  unique_id: the 128 bytes rank 0 got
        from nccl_unique_id(), handed to every rank by any transport (torch.distributed broadcast in bench.py)."""
        uid = None if unique_id is None else (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        self._check(self._lib.lsf_exchange_init(self._ctx, uid, int(rank), int(world), int(max_records)))
        self._world = int(world)

    def allgather_start(self, frame_base=0):
        """Start the exchange of the last batch's kept segments (returns at once; overlaps the next batch)."""
        self._check(self._lib.lsf_allgather_segments(self._ctx, None, int(frame_base)))

    def exchange_wait(self):
        """Finish the oldest exchange -> (device pointer of the gathered 72-byte records, total, per-rank counts)."""
        ptr, n = C.c_void_p(), C.c_int()
        counts = (C.c_int * self._world)()
        self._check(self._lib.lsf_exchange_wait(self._ctx, C.byref(ptr), C.byref(n), counts, self._world))
        return ptr.value, n.value, list(counts)

    def read_device_records(self, ptr, n):
        """Copy n 72-byte records from a device pointer to a host structured array (dist._REC layout)."""
        from . import dist as _dist
        out = np.zeros(n, _dist._REC)
        if n:
            import ctypes
            cudart = ctypes.CDLL("libcudart.so.12") if not hasattr(self, "_cudart") else self._cudart
            self._cudart = cudart
            cudart.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
            rc = cudart.cudaMemcpy(out.ctypes.data, ptr, n * 72, 2)
            if rc != 0:
                raise LsfError(_lib.LSF_E_CUDA, "cudaMemcpy of exchange records failed: %d" % rc)
        return out

    def lane_votes(self, delta_d=0.02, delta_phi=0.1):
        """Vote histograms of the last batch, int32 [n_frames, nd, nphi]: the counts that
        LaneFilterHistogram.generate_measurement_likelihood (lane_filter.py:82-102) accumulates before normalising.
        The grid is np.mgrid[d_min:d_max:delta_d, phi_min:phi_max:delta_phi] like the reference's (23 x 30 by default)."""
        cfg = self.cfg
        nd = len(np.arange(cfg.d_min, cfg.d_max, delta_d))
        nphi = len(np.arange(cfg.phi_min, cfg.phi_max, delta_phi))
        n = self._last_n
        hist = np.zeros((n, nd, nphi), np.int32)
        self._check(self._lib.lsf_lane_votes(self._ctx, float(delta_d), float(delta_phi), nd, nphi, MEM_HOST, hist.ctypes.data))
        return hist

    def reset_sequence(self):
        """Start a new sequence for STAGE_MATCH_PREV (forget the previous batch's last frame)."""
        self._check(self._lib.lsf_reset_sequence(self._ctx))

    def map_clear(self):
        self._check(self._lib.lsf_map_clear(self._ctx))

    def map_add(self, desc):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self._check(self._lib.lsf_map_add(self._ctx, d.ctypes.data, len(d), MEM_HOST))

    def map_size(self):
        return self._lib.lsf_map_size(self._ctx)

    def tap(self, name, frame=0):
        """Dense stage map of one frame of the last batch (parity checks)."""
        t = TAP[name]
        if name == "image":
            out = np.empty((self.h, self.w, 3), np.uint8)
        elif name in ("dx", "dy"):
            out = np.empty((self.h, self.w), np.int16)
        else:
            out = np.empty((self.h, self.w), np.uint8)
        self._check(self._lib.lsf_get_tap(self._ctx, t, int(frame), out.ctypes.data, out.nbytes))
        return out

    def timings(self):
        names = (C.c_char_p * 32)(); ms = (C.c_float * 32)()
        n = self._lib.lsf_last_timings(self._ctx, names, ms, 32)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    def launch_count(self):
        return int(self._lib.lsf_launch_count(self._ctx))

    def stream(self):
        return self._lib.lsf_stream(self._ctx)

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.lsf_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
