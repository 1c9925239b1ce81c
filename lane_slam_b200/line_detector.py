"""Plugin point 1: the detector class selectable from the YAML ``detector:`` entry, exactly like
``line_detector.LineDetectorLSD`` (src/line_detector/include/line_detector/line_detector_lsd.py:11-142,
interface src/line_detector/include/line_detector/line_detector_interface.py:6-19):

    detector:
      - lane_slam_b200.LineDetectorB200
      - configuration: {dilation_kernel_size: 3, canny_thresholds: [80,200], hsv_white1: ..., ...}

``instantiate(c[0], c[1])`` (line_detector_node.py:83-90) calls ``LineDetectorB200(configuration=dict)``.
``setImage`` runs the whole three-colour detection for the frame in one GPU batch call; ``detectLines``
slices the cached result.
"""
from collections import namedtuple

import numpy as np

from .frontend import COLORS, FrontEnd, check_detector_configuration
from ._lib import STAGE_DETECT

Detections = namedtuple('Detections', ['lines', 'normals', 'area', 'centers'])


class LineDetectorInterface(object):
    def setImage(self, bgr):
        raise NotImplementedError

    def detectLines(self, color):
        raise NotImplementedError


class LineDetectorB200(LineDetectorInterface):
    def __init__(self, configuration, device=0):
        # same errors as Configurable: not a dict / extra / missing keys -> ValueError
        self._configuration = check_detector_configuration(configuration)
        for name, value in self._configuration.items():
            if isinstance(value, list) and len(value) == 3:
                value = np.array(value)       # parameters.py:29-32
            setattr(self, name, value)
        self._device = device
        self._fe = None
        self._shape = None
        self._batch = None
        self.bgr = np.empty(0)

    def _front_end(self, shape):
        if self._fe is None or self._shape != shape:
            if self._fe is not None:
                self._fe.close()
            h, w = shape
            # the node has already resized and cropped the frame (line_detector_node.py:163-169)
            self._fe = FrontEnd(self._configuration, img_size=(h, w), top_cutoff=0, src_size=(h, w), max_batch=1,
                                device=self._device)
            self._shape = shape
        return self._fe

    def setImage(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        if bgr.ndim != 3 or bgr.shape[2] != 3:
            raise ValueError("setImage expects an HxWx3 uint8 BGR image")
        self.bgr = np.copy(bgr)
        fe = self._front_end(bgr.shape[:2])
        b = fe.process(self.bgr, stages=STAGE_DETECT)
        # copy out of the reusable buffers
        self._batch = dict(counts=b.counts[0].copy(), lines=b.lines_px.copy(), normals=b.normals.copy(),
                           centers=b.centers.copy())
        self._area = {}

    def detectLines(self, color):
        if color not in COLORS:
            raise Exception('Error: Undefined color strings...')
        if self._batch is None:
            raise Exception('setImage must be called before detectLines')
        ci = COLORS.index(color)
        if ci not in self._area:
            self._area[ci] = self._fe.tap("bw_" + color, 0)
        cnt = self._batch["counts"]
        lo = int(cnt[:ci].sum()); hi = lo + int(cnt[ci])
        if hi == lo:
            # the reference returns plain [] for lines / normals / centers (line_detector_lsd.py:68-71, 87-88)
            return Detections(lines=[], normals=[], area=self._area[ci], centers=[])
        return Detections(lines=self._batch["lines"][lo:hi], normals=self._batch["normals"][lo:hi],
                          area=self._area[ci], centers=self._batch["centers"][lo:hi])

    def getImage(self):
        return self.bgr


class LineDetectorHSVB200(LineDetectorB200):
    """``line_detector.LineDetectorHSV`` (src/line_detector/include/line_detector/line_detector1.py:11-137), the detector eight of the
    ten shipped YAML files select: same configuration keys, same setImage / colour filter, lines from cv2.HoughLinesP with the
    YAML's hough_threshold / hough_min_line_length / hough_max_line_gap.  The Hough transform, the normals and the endpoint
    ordering run on the GPU (lsf_hough_batch) from the maps setImage left on the device.  lines are int32 like the reference's."""

    def setImage(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8)
        if bgr.ndim != 3 or bgr.shape[2] != 3:
            raise ValueError("setImage expects an HxWx3 uint8 BGR image")
        self.bgr = np.copy(bgr)
        fe = self._front_end(bgr.shape[:2])
        fe.process(self.bgr, stages=STAGE_DETECT)          # colour masks + edges (the LSD stage of the batch is not used here)
        b = fe.hough_lines(self.hough_threshold, self.hough_min_line_length, self.hough_max_line_gap, ground=False)
        self._batch = dict(counts=b.counts[0].copy(), lines=b.lines_px.astype(np.int32), normals=b.normals.copy(),
                           centers=b.centers.astype(np.float64))
        self._area = {}
