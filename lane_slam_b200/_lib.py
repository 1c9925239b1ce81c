"""ctypes binding of liblsf.so (include/lsf.h).  No CPU fallback: a missing library or GPU raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblsf.so")

LSF_OK, LSF_E_CONFIG, LSF_E_ARG, LSF_E_CAPACITY, LSF_E_CUDA, LSF_E_NCCL, LSF_E_INTERNAL = 0, -1, -2, -3, -4, -5, -6
MEM_HOST, MEM_PINNED, MEM_DEVICE = 0, 1, 2
TIES_REFERENCE, TIES_INDEX = 0, 1
MATCH_RADIUS = 128
STAGE_DETECT, STAGE_GROUND, STAGE_DESCRIBE, STAGE_MATCH, STAGE_MATCH_PREV = 1, 2, 4, 8, 16
TAP = dict(image=0, labels=1, edges=2, bw_white=3, bw_yellow=4, bw_red=5, ec_white=6, ec_yellow=7, ec_red=8,
           gray=9, dx=10, dy=11)


class LsfConfig(C.Structure):
    _fields_ = [
        ("img_h", C.c_int32), ("img_w", C.c_int32), ("top_cutoff", C.c_int32),
        ("hsv_lo", (C.c_int32 * 3) * 4), ("hsv_hi", (C.c_int32 * 3) * 4),
        ("dilation_kernel_size", C.c_int32), ("canny_lo", C.c_int32), ("canny_hi", C.c_int32),
        ("ai_scale", C.c_float * 3), ("ai_shift", C.c_float * 3),
        ("K", C.c_double * 9), ("D", C.c_double * 5), ("R", C.c_double * 9), ("P", C.c_double * 12),
        ("cam_w", C.c_int32), ("cam_h", C.c_int32), ("Hgnd", C.c_double * 9),
        ("lanewidth", C.c_double), ("linewidth_white", C.c_double), ("linewidth_yellow", C.c_double),
        ("d_min", C.c_double), ("d_max", C.c_double), ("phi_min", C.c_double), ("phi_max", C.c_double),
        ("max_batch", C.c_int32), ("max_src_h", C.c_int32), ("max_src_w", C.c_int32),
        ("max_segments_per_color", C.c_int32), ("max_pixels_per_color", C.c_int32), ("device", C.c_int32),
        ("chunk_frames", C.c_int32), ("tie_order", C.c_int32), ("grow_warps_per_sm", C.c_int32), ("reserved", C.c_int32 * 4),
    ]


class LsfSegments(C.Structure):
    _fields_ = [
        ("mem", C.c_int32), ("capacity", C.c_int32), ("n_frames", C.c_int32), ("n_segments", C.c_int32),
        ("counts", C.c_void_p), ("frame_offset", C.c_void_p), ("color", C.c_void_p), ("lines_px", C.c_void_p),
        ("normals", C.c_void_p), ("centers", C.c_void_p), ("pixels_normalized", C.c_void_p),
        ("normal_f32", C.c_void_p), ("ground", C.c_void_p), ("keep", C.c_void_p), ("desc", C.c_void_p),
        ("match_idx", C.c_void_p), ("match_dist", C.c_void_p),
    ]


class LsfOdometry(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("theta", C.c_double), ("last_t", C.c_double), ("dt", C.c_double)]


_EXPORTS = [
    "lsf_map_append", "lsf_map_append_records", "lsf_map_read", "lsf_match_batch", "lsf_odometry_init", "lsf_odometry_step",
    "lsf_nccl_unique_id", "lsf_exchange_init", "lsf_allgather_segments", "lsf_exchange_wait",
    "lsf_front_end_batch_jpeg", "lsf_lane_filter_init", "lsf_lane_filter_reset", "lsf_lane_filter_batch", "lsf_lane_filter_belief",
    "lsf_default_config", "lsf_create", "lsf_destroy", "lsf_last_error", "lsf_set_color_transform", "lsf_set_chunk_frames",
    "lsf_set_tie_order", "lsf_capacities", "lsf_cancel_prefetch",
    "lsf_front_end_batch", "lsf_prefetch_batch", "lsf_detect_batch", "lsf_describe_batch", "lsf_project_filter_batch",
    "lsf_knn_hamming", "lsf_pack_kept_records", "lsf_lane_votes", "lsf_map_clear", "lsf_map_add", "lsf_map_size", "lsf_reset_sequence", "lsf_get_tap", "lsf_image_dims",
    "lsf_last_timings", "lsf_launch_count", "lsf_stream", "lsf_version", "lsf_hough_batch",
]

_lib = None


class LsfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("lsf error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load liblsf.so.  Raises (loudly) when the CUDA extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("lane_slam_b200: %s is missing -- build it with `python -m lane_slam_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.lsf_default_config.argtypes = [C.POINTER(LsfConfig)]
    lib.lsf_create.argtypes = [C.POINTER(LsfConfig), C.POINTER(vp)]
    lib.lsf_destroy.argtypes = [vp]
    lib.lsf_destroy.restype = None
    lib.lsf_last_error.argtypes = [vp]
    lib.lsf_last_error.restype = C.c_char_p
    lib.lsf_set_color_transform.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.lsf_set_chunk_frames.argtypes = [vp, i32]
    lib.lsf_set_tie_order.argtypes = [vp, i32]
    lib.lsf_capacities.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.lsf_cancel_prefetch.argtypes = [vp]
    lib.lsf_front_end_batch.argtypes = [vp, vp, i32, i32, i32, sz, i32, i32, i32, C.POINTER(LsfSegments)]
    lib.lsf_prefetch_batch.argtypes = [vp, vp, i32, i32, i32, sz]
    lib.lsf_front_end_batch_jpeg.argtypes = [vp, vp, vp, i32, i32, i32, C.POINTER(LsfSegments)]
    lib.lsf_detect_batch.argtypes = [vp, vp, i32, i32, i32, sz, i32, C.POINTER(LsfSegments)]
    lib.lsf_describe_batch.argtypes = [vp, C.POINTER(LsfSegments)]
    lib.lsf_project_filter_batch.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    lib.lsf_knn_hamming.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, vp, vp]
    lib.lsf_pack_kept_records.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i32)]
    lib.lsf_lane_votes.argtypes = [vp, C.c_double, C.c_double, i32, i32, i32, vp]
    lib.lsf_map_clear.argtypes = [vp]
    lib.lsf_map_add.argtypes = [vp, vp, i32, i32]
    lib.lsf_map_size.argtypes = [vp]
    lib.lsf_map_append.argtypes = [vp, vp, i32]
    lib.lsf_map_append_records.argtypes = [vp, vp, i32, i32, vp, i32, i32]
    lib.lsf_map_read.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    lib.lsf_match_batch.argtypes = [vp, i32, i32, vp, vp]
    lib.lsf_hough_batch.argtypes = [vp, i32, i32, i32, i32, C.POINTER(LsfSegments)]
    lib.lsf_odometry_init.argtypes = [C.POINTER(LsfOdometry), C.c_double]
    lib.lsf_odometry_step.argtypes = [C.POINTER(LsfOdometry), C.c_double, C.c_double, C.c_double]
    lib.lsf_lane_filter_init.argtypes = [vp, vp]
    lib.lsf_lane_filter_reset.argtypes = [vp]
    lib.lsf_lane_filter_batch.argtypes = [vp, vp, i32, vp]
    lib.lsf_lane_filter_belief.argtypes = [vp, vp]
    lib.lsf_nccl_unique_id.argtypes = [vp]
    lib.lsf_exchange_init.argtypes = [vp, vp, i32, i32, i32]
    lib.lsf_allgather_segments.argtypes = [vp, vp, i32]
    lib.lsf_exchange_wait.argtypes = [vp, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), i32]
    lib.lsf_reset_sequence.argtypes = [vp]
    lib.lsf_get_tap.argtypes = [vp, i32, i32, vp, sz]
    lib.lsf_image_dims.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.lsf_last_timings.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), i32]
    lib.lsf_launch_count.argtypes = [vp]
    lib.lsf_launch_count.restype = C.c_longlong
    lib.lsf_stream.argtypes = [vp]
    lib.lsf_stream.restype = vp
    lib.lsf_version.restype = C.c_char_p
    _lib = lib
    return lib


def exported_symbols():
    return list(_EXPORTS)
