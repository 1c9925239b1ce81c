/*
 * lsf.h -- C ABI of the B200 line front end ("lane-slam front", lsf).
 *
 * Drop-in boundary for the per-frame line path of mandanasmi/lane-slam.  Every entry point
 * names the reference interface it replaces (file:line relative to the reference repo).
 * Plain pointers and sizes only; no C++/torch types.  All functions return 0 (LSF_OK) or a
 * negative lsf_status; no exception crosses the boundary; lsf_last_error() gives the text.
 *
 * Ownership: the caller allocates every output with an explicit capacity; the library owns
 * only device scratch inside lsf_ctx.  A ctx is single-owner / not re-entrant (the reference
 * admits one processImage_ at a time: src/line_detector/src/line_detector_node.py:129-139);
 * lsf_set_color_transform is the only call allowed concurrently with a batch (from another thread; it swaps the six
 * values under a lock and the batch snapshots them under the same lock when it starts).  Several contexts -- on the
 * same device or on different devices -- may live in one process and be driven from different threads.
 */
#ifndef LSF_H
#define LSF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LSF_API __attribute__((visibility("default")))
#else
#define LSF_API
#endif

typedef enum lsf_status {
    LSF_OK = 0,
    LSF_E_CONFIG = -1,   /* bad configuration      (reference: ValueError, duckietown_utils/parameters.py:5-23) */
    LSF_E_ARG = -2,      /* bad argument           (reference: Exception, line_detector_lsd.py:49) */
    LSF_E_CAPACITY = -3, /* caller/scratch capacity too small; lsf_last_error names the required size */
    LSF_E_CUDA = -4,     /* CUDA runtime / driver failure, or no sm_100 device */
    LSF_E_NCCL = -5,
    LSF_E_INTERNAL = -6
} lsf_status;

typedef enum lsf_mem_kind { LSF_MEM_HOST = 0, LSF_MEM_PINNED = 1, LSF_MEM_DEVICE = 2 } lsf_mem_kind;

/* Order of neighbours of EQUAL Hamming distance in every k-NN result.
 * LSF_TIES_REFERENCE (default): the order BinaryDescriptorMatcher::knnMatch itself returns, i.e. the discovery order of
 *   Mihasher(256, 32) (binary_descriptor_matcher.cpp:276, :634-753) -- by (smallest per-byte XOR popcount, first byte
 *   reaching it, that XOR byte, train index); verified against the reference's compiled code (tests/golden/lbd_reference.npz).
 * LSF_TIES_INDEX: ascending train index (what cv2.BFMatcher returns). */
enum { LSF_TIES_REFERENCE = 0, LSF_TIES_INDEX = 1 };

/* Search radius of the matching STAGES (lsf_front_end_batch): Mihasher's D = ceil(256 / 2) (binary_descriptor_matcher.cpp:761);
 * farther neighbours are reported as -1.  lsf_knn_hamming takes the radius as an argument. */
#define LSF_MATCH_RADIUS 128

/* Segment.msg:1-3 */
enum { LSF_WHITE = 0, LSF_YELLOW = 1, LSF_RED = 2 };

/*
 * POD mirror of the YAML / rosparam configuration of the path:
 *   line_detector_node/default.yaml:1-23 (img_size, top_cutoff, detector configuration),
 *   AntiInstagramTransform (line_detector_node.py:112-114), camera_intrinsic/default.yaml:1-20,
 *   camera_extrinsic/default.yaml:1, line_sanity_node.py:17-23.
 * LSD itself runs with OpenCV's defaults + LSD_REFINE_ADV (line_detector_lsd.py:65); those are
 * fixed in the kernels (scale 0.8, sigma_scale 0.6, quant 2, ang_th 22.5, log_eps 0, density 0.7, 1024 bins).
 */
typedef struct lsf_config {
    int32_t img_h, img_w;        /* img_size = [h, w]; frames of another size are nearest-resized to it */
    int32_t top_cutoff;          /* rows [0, top_cutoff) of the resized frame are dropped */
    int32_t hsv_lo[4][3];        /* white1, yellow1, red1, red3 */
    int32_t hsv_hi[4][3];        /* white2, yellow2, red2, red4 */
    int32_t dilation_kernel_size;/* 3 (cross) or 1 (none); other sizes -> LSF_E_CONFIG */
    int32_t canny_lo, canny_hi;  /* canny_thresholds */
    float ai_scale[3], ai_shift[3]; /* AntiInstagram scale / shift per B,G,R channel; identity = 1 / 0 */
    double K[9], D[5], R[9], P[12]; /* CameraInfo */
    int32_t cam_w, cam_h;
    double Hgnd[9];              /* ground homography */
    double lanewidth, linewidth_white, linewidth_yellow, d_min, d_max, phi_min, phi_max;
    /* capacities (0 = library default) */
    int32_t max_batch;           /* frames per call the scratch is sized for */
    int32_t max_src_h, max_src_w;/* largest input frame */
    int32_t max_segments_per_color; /* per frame and colour */
    int32_t max_pixels_per_color;   /* LSD support pixels per frame and colour; 0 = sized by the library, which then grows it and reruns
                                       the batch when a frame needs more (a value set here is a hard limit: LSF_E_CAPACITY) */
    int32_t device;              /* CUDA device ordinal */
    int32_t chunk_frames;        /* frames per pipeline chunk: a batch is cut into chunks whose host->device copy and
                                    kernels overlap on several streams.  0 = automatic (n/8 for host frames, n/2 for
                                    frames already on the device, one chunk below 64 frames), < 0 = never chunk (one
                                    stream; lsf_last_timings then lists every kernel) */
    int32_t tie_order;           /* LSF_TIES_REFERENCE (0, default) or LSF_TIES_INDEX */
    int32_t grow_warps_per_sm;   /* persistent region-growing warps per SM (0 = default 16, the fastest for ONE batch in flight).
                                    Each holds ~122 registers x 32 lanes; with several contexts in flight on one GPU (one
                                    host thread each) 6-10 leave room for the dense kernels of the other batches, so that
                                    the latency-bound search of one batch runs under the issue-bound kernels of another */
    int32_t reserved[4];
} lsf_config;

/*
 * SoA segment batch.  Order = frame, then white / yellow / red, then LSD acceptance order
 * (line_detector_node.py:197-205).  Arrays the caller leaves NULL are skipped.
 * All arrays live in `mem` (host/pinned or device).
 */
typedef struct lsf_segments {
    int32_t mem;               /* lsf_mem_kind of every pointer below */
    int32_t capacity;          /* rows available in the per-segment arrays */
    int32_t n_frames;          /* out */
    int32_t n_segments;        /* out: total rows written */
    int32_t *counts;           /* [n_frames][3] segments per frame and colour (Detections.lines lengths) */
    int32_t *frame_offset;     /* [n_frames+1] first row of every frame */
    uint8_t *color;            /* [S]      Segment.color */
    float *lines_px;           /* [S][4]   Detections.lines: x1,y1,x2,y2 in cropped-image pixels, after endpoint ordering */
    double *normals;           /* [S][2]   Detections.normals (float64 in the reference) */
    float *centers;            /* [S][2]   Detections.centers */
    float *pixels_normalized;  /* [S][4]   Segment.pixels_normalized (Vector2D float32) */
    float *normal_f32;         /* [S][2]   Segment.normal (Vector2D float32) */
    double *ground;            /* [S][4]   Segment.points x1,y1,x2,y2 (geometry_msgs/Point float64, z = 0) */
    uint8_t *keep;             /* [S]      1 iff line_sanity keeps the segment */
    uint8_t *desc;             /* [S][32]  LBD binary descriptor */
    int32_t *match_idx;        /* [S][k]   kNN train index in the map (or -1) */
    int32_t *match_dist;       /* [S][k]   Hamming distance (or -1) */
} lsf_segments;

typedef struct lsf_ctx lsf_ctx;

/* stage selectors for lsf_front_end_batch */
enum {
    LSF_STAGE_DETECT = 1,   /* line_detector_node: resize/crop/colour-correct, HSV masks, Canny, 3x LSD, normals */
    LSF_STAGE_GROUND = 2,   /* ground_projection_node + line_sanity_node */
    LSF_STAGE_DESCRIBE = 4, /* LSDDetectorC KeyLine fill + BinaryDescriptor::compute */
    LSF_STAGE_MATCH = 8,    /* BinaryDescriptorMatcher::knnMatch against the ctx map */
    LSF_STAGE_MATCH_PREV = 16 /* knnMatch(frame t, frame t-1): frame-to-frame association; trainIdx is local to frame t-1;
                                 frame 0 of a batch matches the last frame of the previous batch (lsf_reset_sequence clears it) */
};

/* debug / parity taps (dense per-frame maps, one byte per pixel, 0/255 or label values) */
enum {
    LSF_TAP_IMAGE = 0,     /* processed BGR image [h][w][3] (after resize/crop/colour correction) */
    LSF_TAP_LABELS = 1,    /* bit0 white, bit1 yellow, bit2 red raw inRange masks; bits 4-5 Canny NMS class (0,1,2) */
    LSF_TAP_EDGES = 2,     /* cv2.Canny output 0/255 */
    LSF_TAP_BW_WHITE = 3, LSF_TAP_BW_YELLOW = 4, LSF_TAP_BW_RED = 5,      /* Detections.area */
    LSF_TAP_EC_WHITE = 6, LSF_TAP_EC_YELLOW = 7, LSF_TAP_EC_RED = 8,      /* edge_color fed to LSD */
    LSF_TAP_GRAY = 9,      /* BGR2GRAY [h][w] */
    LSF_TAP_DX = 10, LSF_TAP_DY = 11 /* int16 [h][w] Sobel of the 5x5-blurred gray (descriptor path) */
};

/* Fill *cfg with the reference defaults (default.yaml files cited above) for a src_h x src_w camera. */
LSF_API int lsf_default_config(lsf_config *cfg);

/* Replaces: instantiate(c[0], c[1]) / LineDetectorLSD.__init__ (line_detector_node.py:83-90,
 * line_detector_lsd.py:14-36) + GroundProjection.__init__ (GroundProjection.py:17-36) +
 * LineSanityNode.__init__ (line_sanity_node.py:15-23). */
LSF_API int lsf_create(const lsf_config *cfg, lsf_ctx **out);
LSF_API void lsf_destroy(lsf_ctx *ctx);
LSF_API const char *lsf_last_error(const lsf_ctx *ctx); /* ctx may be NULL: last create() error */

/* Replaces: LineDetectorNode.cbTransform (line_detector_node.py:112-114).  Takes effect next batch. */
LSF_API int lsf_set_color_transform(lsf_ctx *ctx, const float scale[3], const float shift[3]);

/* Change lsf_config.chunk_frames of a live ctx (same meaning); takes effect at the next batch. */
LSF_API int lsf_set_chunk_frames(lsf_ctx *ctx, int chunk_frames);

/* Change lsf_config.tie_order of a live ctx. */
LSF_API int lsf_set_tie_order(lsf_ctx *ctx, int tie_order);

/* Capacities the ctx was created with: segments / LSD support pixels per frame and colour, and output rows per batch
 * (= what lsf_segments.capacity never needs to exceed). */
LSF_API int lsf_capacities(const lsf_ctx *ctx, int *max_segments_per_color, int *max_pixels_per_color, int *max_output_rows);

/* Replaces: LineDetectorNode.processImage_ (line_detector_node.py:141-213) for n frames at once,
 * i.e. cv2.resize/crop, AntiInstagram.applyTransform + convertScaleAbs, LineDetectorLSD.setImage and
 * detectLines x3 (line_detector_lsd.py:38-139), normalisation and toSegmentMsg (:195-205, :251-265);
 * with stages |= GROUND also GroundProjectionNode.lineseglist_cb (ground_projection_node.py:55-65) and
 * LineSanityNode.processSegmentList (line_sanity_node.py:48-72); with DESCRIBE also
 * LSDDetectorC::detectImpl KeyLine fill (LSDDetector_custom.cpp:176-197) and BinaryDescriptor::compute
 * (binary_descriptor_custom.cpp:539-687); with MATCH also BinaryDescriptorMatcher::knnMatch
 * (binary_descriptor_matcher.cpp:258-335) of every segment against the ctx map (k neighbours).
 * bgr: n frames of src_h x src_w x 3 bytes, row pitch `pitch` bytes, frame stride src_h*pitch. */
LSF_API int lsf_front_end_batch(lsf_ctx *ctx, const uint8_t *bgr, int n, int src_h, int src_w, size_t pitch,
                                int mem_kind, int stages, int k, lsf_segments *out);

/* Same as lsf_front_end_batch for frames that are still JPEG files -- what the node receives: sensor_msgs/CompressedImage,
 * decoded by image_cv_from_jpg = cv2.imdecode(..., IMREAD_COLOR) (src/duckietown/include/duckietown_utils/jpg.py:21-31, called at
 * line_detector_node.py:155).  blob holds the n files back to back (host memory, pinned for an asynchronous copy), file i =
 * bytes [offsets[i], offsets[i+1]).  The files cross PCIe compressed and are decoded on the GPU, bit-identical to cv2.imdecode
 * (libjpeg-turbo defaults: islow IDCT, fancy chroma upsampling, fixed-point YCbCr->BGR).  Scope: baseline Huffman JPEG, 8 bit,
 * 1 or 3 components, sampling factors <= 2 (4:4:4, 4:2:2, 4:4:0, 4:2:0, gray), no restart intervals, all frames of a batch
 * the same size and sampling; anything else -> LSF_E_ARG (the reference logs and skips undecodable frames,
 * line_detector_node.py:154-158: decode such a frame on the host and use lsf_front_end_batch). */
LSF_API int lsf_front_end_batch_jpeg(lsf_ctx *ctx, const uint8_t *blob, const int64_t *offsets, int n, int stages, int k, lsf_segments *out);

/* Streaming replay: start the host->device copy of the NEXT batch of host frames now (returns immediately; the copy
 * runs on the ctx's copy stream into the spare staging buffer) so that it overlaps the kernels of the batch being
 * processed.  A following lsf_front_end_batch call with the same `bgr` pointer and geometry consumes the staged
 * frames instead of copying again; up to two batches may be staged ahead.  The caller must keep `bgr` (pinned
 * memory for a truly asynchronous copy) unchanged until that call returns.  A staged batch is matched by pointer and
 * geometry only: any lsf_front_end_batch call on host frames that does not match drops everything staged, and
 * lsf_cancel_prefetch drops it explicitly (abandoned replay, buffer about to be reused with other contents). */
LSF_API int lsf_prefetch_batch(lsf_ctx *ctx, const uint8_t *bgr, int n, int src_h, int src_w, size_t pitch);
LSF_API int lsf_cancel_prefetch(lsf_ctx *ctx);

/* = lsf_front_end_batch(..., LSF_STAGE_DETECT, 0, out) */
LSF_API int lsf_detect_batch(lsf_ctx *ctx, const uint8_t *bgr, int n, int src_h, int src_w, size_t pitch,
                             int mem_kind, lsf_segments *out);

/* Replaces: BinaryDescriptor::compute (binary_descriptor_custom.cpp:524-687) on the frames of the last
 * lsf_detect_batch / lsf_front_end_batch call (still resident on the device): segs->lines_px,
 * segs->frame_offset (n_frames+1) in -> segs->desc out. */
LSF_API int lsf_describe_batch(lsf_ctx *ctx, lsf_segments *segs);

/* Replaces: GroundProjectionNode.lineseglist_cb + LineSanityNode.processSegmentList for S segments:
 * pixels_normalized f32 [S][4], color u8 [S]  ->  ground f64 [S][4], keep u8 [S].  Host or device per mem_kind. */
LSF_API int lsf_project_filter_batch(lsf_ctx *ctx, const float *pixels_normalized, const uint8_t *color, int n_seg,
                                     int mem_kind, double *ground, uint8_t *keep);

/* Replaces: BinaryDescriptorMatcher::knnMatch(query, train, k) (binary_descriptor_matcher.cpp:258-335).
 * Exact brute force over 32-byte codes; rows ascending by distance, ties per lsf_config.tie_order (default: the
 * reference's own order); neighbours farther than max_dist (Mihasher D = 128 = LSF_MATCH_RADIUS; pass 256 for
 * unbounded) are reported as -1 (the reference leaves those entries uninitialised). */
LSF_API int lsf_knn_hamming(lsf_ctx *ctx, const uint8_t *query, int nq, const uint8_t *train, int nm, int k,
                            int max_dist, int mem_kind, int32_t *idx, int32_t *dist);

/* Map of accumulated lines, device resident: per line the segment in the MAP frame, colour, the global id of the frame it
 * was seen in, and its descriptor.  It is what LSF_STAGE_MATCH / lsf_match_batch search (BinaryDescriptorMatcher::add /
 * train / clear, binary_descriptor_matcher.cpp:55-105) and what show_map accumulates (src/show_map/src/show_map.py:28-43:
 * an append-only store of every received segment; RViz places them with the map->duck transform the odometry node
 * broadcasts, src/odometry/src/odometry.py:110-120). */
LSF_API int lsf_map_clear(lsf_ctx *ctx);
/* descriptors only (segment = 0, colour = 255, frame = -1) */
LSF_API int lsf_map_add(lsf_ctx *ctx, const uint8_t *desc, int n, int mem_kind);
LSF_API int lsf_map_size(lsf_ctx *ctx);
/* Append the segments line_sanity kept in the last batch: the ground segment of frame f (robot frame "duck") is moved to the
 * map frame with that frame's pose, p_map = R(theta_f) p + (x_f, y_f); poses = [n_frames][3] = x, y, theta per frame of the
 * batch (what lsf_odometry_step integrates), NULL = identity.  frame_base = global id of the batch's first frame. */
LSF_API int lsf_map_append(lsf_ctx *ctx, const double *poses, int frame_base);
/* Same for n 72-byte exchange records (layout below; host or device) -- e.g. the gathered records of all ranks.  The pose
 * of record r is poses[frame_id(r) - pose_frame_base]; records of frames outside [pose_frame_base, +n_poses) are appended
 * untransformed. */
LSF_API int lsf_map_append_records(lsf_ctx *ctx, const void *records, int n, int mem_kind, const double *poses, int pose_frame_base,
                                   int n_poses);
/* Read lines [first, first + count) back to host arrays (NULL = skip): ground f64 [count][4], colour u8, frame i32, desc u8 [count][32]. */
LSF_API int lsf_map_read(lsf_ctx *ctx, int first, int count, double *ground, uint8_t *color, int32_t *frame, uint8_t *desc);

/* BinaryDescriptorMatcher::knnMatch (dataset form, binary_descriptor_matcher.cpp:339-424) of every descriptor of the LAST batch
 * against the map as it is now: match_idx / match_dist [n_segments][k], host or device per mem_kind.  The epoch replay
 *  detects epoch e, appends the gathered lines of epoch e-1, then matches. */
LSF_API int lsf_match_batch(lsf_ctx *ctx, int k, int mem_kind, int32_t *match_idx, int32_t *match_dist);

/* Diff-drive odometry of the reference (OdometryNode.getPose / drive, src/odometry/src/odometry.py:66-120): host arithmetic.
 * stamp_nsecs is the `nsecs` field of the wheels-command stamp (the reference uses only that field).  lsf_odometry_step
 * returns 1 when the pose was advanced (0 < dt < 0.3), 0 when the reference would skip the update, < 0 on error. */
typedef struct lsf_odometry { double x, y, theta, last_t, dt; } lsf_odometry;
LSF_API int lsf_odometry_init(lsf_odometry *st, double stamp_nsecs);
LSF_API int lsf_odometry_step(lsf_odometry *st, double stamp_nsecs, double vel_left, double vel_right);

/* Multi-GPU exchange step: pack the segments of the last batch that line_sanity kept into 72-byte
 * records  { int32 frame_id (+ frame_base), uint8 color, pad[3], double ground[4], uint8 desc[32] }  in a
 * ctx-owned DEVICE buffer, ready to be all-gathered (NCCL) and appended to every rank's map.  *records stays
 * valid until the next call on the ctx.  Needs LSF_STAGE_GROUND (and LSF_STAGE_DESCRIBE for the descriptors). */
LSF_API int lsf_pack_kept_records(lsf_ctx *ctx, int frame_base, void **records, int *n_records);

/* The exchange step itself, in the library (no host round trip between packing and the collective):
 *   lsf_nccl_unique_id   rank 0 creates the 128-byte NCCL id and hands it to the other ranks (any transport);
 *   lsf_exchange_init    one communicator + double-buffered fixed-capacity slots per ctx; max_records = kept segments one rank
 *                        may contribute per exchange (0 = 64 per frame of max_batch); world = 1 needs no NCCL;
 *   lsf_allgather_segments  START the exchange of the last batch's kept segments (needs LSF_STAGE_GROUND): the packing kernel
 *                        runs on the ctx stream, then ONE ncclAllGather of { count, records } slots and a compaction kernel run on
 *                        the ctx's exchange stream behind an event, overlapping whatever is launched next (up to two in flight).
 *                        comm: an ncclComm_t of the caller, or NULL for the ctx's own;
 *   lsf_exchange_wait    finish the oldest exchange in flight: *records = device array of all ranks' 72-byte records in rank
 *                        order (= global frame order for contiguous shards), *n_total, counts[rank].
 * NCCL (libnccl.so.2, e.g. the one PyTorch ships; LSF_NCCL_LIB overrides the name) is loaded at run time; failures -> LSF_E_NCCL. */
LSF_API int lsf_nccl_unique_id(void *id128);
LSF_API int lsf_exchange_init(lsf_ctx *ctx, const void *unique_id128, int rank, int world, int max_records);
LSF_API int lsf_allgather_segments(lsf_ctx *ctx, void *comm, int frame_base);
LSF_API int lsf_exchange_wait(lsf_ctx *ctx, void **records, int *n_total, int *counts, int counts_cap);

/* First consumer of the path.  Replaces the vote loop of LaneFilterHistogram.generate_measurement_likelihood
 * (src/lane_filter/include/lane_filter/lane_filter.py:82-102, generateVote :123-155) for every frame of the last batch:
 * hist[f][i][j] = number of kept ground segments of frame f whose vote (d_i, phi_i) falls into cell
 * i = floor((d_i - d_min) / delta_d), j = floor((phi_i - phi_min) / delta_phi)   (int32 [n_frames][nd][nphi], host or
 * device per mem_kind; the measurement likelihood is hist[f] / sum(hist[f])).  d_min .. phi_max, lanewidth and the line
 * widths are those of lsf_config (identical in the reference's line_sanity and lane_filter defaults).
 * Needs LSF_STAGE_GROUND in the last batch. */
LSF_API int lsf_lane_votes(lsf_ctx *ctx, double delta_d, double delta_phi, int nd, int nphi, int mem_kind, int32_t *hist);

/* The whole histogram lane filter for the frames of the last batch, in order: per frame predict(dt, v, w) (optional), update with
 * the frame's votes, getEstimate / getMax -- LaneFilterHistogram.predict / update / getEstimate / getMax
 * (src/lane_filter/include/lane_filter/lane_filter.py:47-112) as LaneFilterNode.processSegments drives them
 * (src/lane_filter/src/lane_filter_node.py:53-65).  The belief lives on the device between batches.  Bit-identical to the
 * reference class: every float64 operation in numpy's / scipy.ndimage's order; the tables that involve transcendental functions
 * come from the host (numpy / scipy there are what the reference calls):
 *   d_grid, phi_grid [nd*nphi]  np.mgrid[d_min:d_max:delta_d, phi_min:phi_max:delta_phi]          (lane_filter.py:38)
 *   sin_phi [nd*nphi]           np.sin(phi_grid)                                                    (:49)
 *   w_d [r_d + 1], w_phi [r_phi + 1]   scipy.ndimage Gaussian mask weights at distance 0 .. r for sigma_d_mask / sigma_phi_mask,
 *                               r = int(4 sigma + 0.5)                                              (:66, gaussian_filter)
 *   belief0 [nd*nphi]           multivariate_normal(mean_0, cov_0).pdf(grid)                        (:114-120)
 * lsf_lane_filter_batch: dt_v_w [n_frames][3] host (NULL when use_propagation = 0); estimates [n_frames][3] host = d, phi of the
 * belief's arg-max cell centre and the belief maximum after each frame. */
typedef struct lsf_lane_filter_config {
    int32_t nd, nphi, r_d, r_phi;
    double d_min, d_max, phi_min, phi_max, delta_d, delta_phi;
    const double *d_grid, *phi_grid, *sin_phi, *w_d, *w_phi, *belief0;
} lsf_lane_filter_config;
LSF_API int lsf_lane_filter_init(lsf_ctx *ctx, const lsf_lane_filter_config *cfg);
LSF_API int lsf_lane_filter_reset(lsf_ctx *ctx);                 /* belief <- belief0 (LaneFilterHistogram.initialize) */
LSF_API int lsf_lane_filter_batch(lsf_ctx *ctx, const double *dt_v_w, int use_propagation, double *estimates);
LSF_API int lsf_lane_filter_belief(lsf_ctx *ctx, double *belief); /* [nd*nphi] host */

/* Forget the previous batch's last frame (start of a new sequence for LSF_STAGE_MATCH_PREV). */
LSF_API int lsf_reset_sequence(lsf_ctx *ctx);

/* Parity taps: copy a dense stage map of frame `frame` of the last batch to host memory `dst`.  LSF_TAP_IMAGE re-reads
 * the input frames: for LSF_MEM_DEVICE input the caller must keep them alive until then; it fails with LSF_E_ARG once
 * a later lsf_prefetch_batch has overwritten the staging buffer that held them. */
LSF_API int lsf_get_tap(lsf_ctx *ctx, int tap, int frame, void *dst, size_t dst_bytes);

/* Processed-image geometry for an input of src_h x src_w: h = img_h - top_cutoff, w = img_w. */
LSF_API int lsf_image_dims(const lsf_ctx *ctx, int *h, int *w, int *lsd_h, int *lsd_w);

/* Device-time of the kernels of the last batch (ms, CUDA events on the ctx stream), by stage name.
 * names/ms arrays of capacity cap; returns the number of entries. */
LSF_API int lsf_last_timings(lsf_ctx *ctx, const char **names, float *ms, int cap);

/* Number of kernels this library launched since ctx creation. */
LSF_API long long lsf_launch_count(const lsf_ctx *ctx);

/* The CUDA stream (cudaStream_t) the ctx launches on, for callers that time with their own events. */
LSF_API void *lsf_stream(lsf_ctx *ctx);

LSF_API const char *lsf_version(void);

/* ---- alternative detector: LineDetectorHSV (src/line_detector/include/line_detector/line_detector1.py:11-137; the class eight of
 * the ten shipped line_detector_node YAML files select).  setImage and the colour filter are the ones of LineDetectorLSD and are
 * on the device after lsf_front_end_batch / lsf_front_end_batch_jpeg (LSF_STAGE_DETECT); this call replaces detectLines for the
 * three colours of every frame of that LAST batch: cv2.HoughLinesP(edge_color, 1, pi/180, hough_threshold, minLineLength =
 * hough_min_line_length, maxLineGap = hough_max_line_gap) (:63-69, bit-identical line lists, same order), _findNormal and
 * _correctPixelOrdering (:71-119, float64), toSegmentMsg's normalisation (line_detector_node.py:251-265) and -- project_ground != 0
 * -- ground projection + line sanity as in LSF_STAGE_GROUND.  out: color, lines_px (the int32 endpoints, exactly representable),
 * normals, centers, pixels_normalized, normal_f32, [ground, keep], counts [n][3], frame_offset [n + 1]; desc / match_* are not
 * written (the reference's LineDetectorHSV has no descriptors).  Errors: LSF_E_ARG without a previous batch, LSF_E_CONFIG for
 * parameters HoughLinesP would turn into degenerate lines (min length < 1), LSF_E_CAPACITY. */
LSF_API int lsf_hough_batch(lsf_ctx *ctx, int hough_threshold, int hough_min_line_length, int hough_max_line_gap, int project_ground,
                            lsf_segments *out);

#ifdef __cplusplus
}
#endif
#endif /* LSF_H */
