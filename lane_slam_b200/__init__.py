"""lane_slam_b200 -- B200-native (sm_100a) line front end of mandanasmi/lane-slam.

Drop-in for the per-frame line path only (line_detector -> ground_projection -> line_sanity, plus LBD
descriptors and Hamming association): the plugin class ``LineDetectorB200`` mirrors
``line_detector.LineDetectorLSD`` and ``FrontEnd`` is the node-level batch API.  All arithmetic runs in
``liblsf.so`` (hand-written CUDA behind the C ABI of include/lsf.h); importing this package never
touches ``oracle/`` and there is no CPU fallback.
"""
from ._lib import (LsfError, STAGE_DESCRIBE, STAGE_DETECT, STAGE_GROUND, STAGE_MATCH, STAGE_MATCH_PREV, MEM_DEVICE, MEM_HOST,
                   TIES_REFERENCE, TIES_INDEX, MATCH_RADIUS, LIB_PATH, exported_symbols)
from .frontend import (FrontEnd, SegmentBatch, DEFAULT_DETECTOR_CONFIGURATION, DETECTOR_PARAM_NAMES, COLORS,
                       WHITE, YELLOW, RED, scaled_calibration, check_detector_configuration)
from .line_detector import LineDetectorB200, LineDetectorHSVB200, Detections, LineDetectorInterface
from .lane_filter import LaneFilterB200
from . import wire, odometry, replay
from .messages import Segment, SegmentList, Vector2D, Point, segment_lists_from_batch

STAGE_ALL = STAGE_DETECT | STAGE_GROUND | STAGE_DESCRIBE
