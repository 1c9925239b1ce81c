"""CPU ORACLE (test infrastructure): the reference's Python glue restated for OpenCV 4.

Every function cites the reference file:line (relative to /root/reference) it follows.  The
per-pixel arithmetic is done by the SAME third-party library the reference calls (cv2; the
contract version is opencv-python-headless 4.13.0), so this module is
the ground truth the CUDA path is compared with.  ROS message plumbing is replaced by plain
numpy arrays with the message field dtypes (Vector2D float32, Point float64).
"""
from collections import namedtuple

import cv2
import numpy as np

# src/line_detector/include/line_detector/line_detector_interface.py:6-7
Detections = namedtuple("Detections", ["lines", "normals", "area", "centers"])

# src/duckietown/config/baseline/line_detector/line_detector_node/default.yaml:1-23
DEFAULT_DETECTOR_CONFIG = dict(
    dilation_kernel_size=3, canny_thresholds=[80, 200],
    hough_threshold=2, hough_min_line_length=3, hough_max_line_gap=1,
    hsv_white1=[0, 0, 150], hsv_white2=[180, 60, 255],
    hsv_yellow1=[25, 140, 100], hsv_yellow2=[45, 255, 255],
    hsv_red1=[0, 140, 100], hsv_red2=[15, 255, 255],
    hsv_red3=[165, 140, 100], hsv_red4=[180, 255, 255])
DEFAULT_IMG_SIZE = (120, 160)
DEFAULT_TOP_CUTOFF = 40

# src/duckietown/config/baseline/calibration/camera_intrinsic/default.yaml:1-20
DEFAULT_CAMERA = dict(
    width=640, height=480,
    K=[307.7379294605756, 0, 329.692367951685, 0, 314.9827773443905, 244.4605588877848, 0, 0, 1],
    D=[-0.2565888993516047, 0.04481160508242147, -0.00505275149956019, 0.001308569367976665, 0],
    R=[1, 0, 0, 0, 1, 0, 0, 0, 1],
    P=[210.1107940673828, 0, 327.2577820024981, 0, 0, 253.8408660888672, 239.9969353923052, 0, 0, 0, 1, 0])
# src/duckietown/config/baseline/calibration/camera_extrinsic/default.yaml:1
DEFAULT_HOMOGRAPHY = [-4.89775e-05, -0.0002150858, -0.1818273, 0.00099274, 1.202336e-06, -0.3280241,
                      -0.0004281805, -0.007185673, 1]

WHITE, YELLOW, RED = 0, 1, 2  # src/duckietown_msgs/msg/Segment.msg:1-3
COLORS = ("white", "yellow", "red")

PARAM_NAMES = ['hsv_white1', 'hsv_white2', 'hsv_yellow1', 'hsv_yellow2', 'hsv_red1', 'hsv_red2', 'hsv_red3',
               'hsv_red4', 'dilation_kernel_size', 'canny_thresholds', 'hough_threshold',
               'hough_min_line_length', 'hough_max_line_gap']


def scaled_camera(W, H):
    """default.yaml calibration scaled from 640x480 to (W,H)."""
    sx, sy = W / 640.0, H / 480.0
    K = np.array(DEFAULT_CAMERA["K"], np.float64).reshape(3, 3).copy()
    P = np.array(DEFAULT_CAMERA["P"], np.float64).reshape(3, 4).copy()
    K[0] *= sx; K[1] *= sy
    P[0] *= sx; P[1] *= sy
    Hg = np.array(DEFAULT_HOMOGRAPHY, np.float64).reshape(3, 3) @ np.diag([1 / sx, 1 / sy, 1.0])
    return dict(width=W, height=H, K=K.ravel().tolist(), D=list(DEFAULT_CAMERA["D"]),
                R=list(DEFAULT_CAMERA["R"]), P=P.ravel().tolist()), Hg.ravel().tolist()


def check_configuration(configuration):
    """duckietown_utils/parameters.py:2-36: exact key set or ValueError; 3-element lists become np.array."""
    if not isinstance(configuration, dict):
        raise ValueError('Expecting a dict, obtained %r' % (configuration,))
    extra = set(configuration) - set(PARAM_NAMES)
    missing = set(PARAM_NAMES) - set(configuration)
    if extra or missing:
        raise ValueError('Extra parameters: %r\nMissing parameters: %r\n' % (extra, missing))
    out = {}
    for name in PARAM_NAMES:
        v = configuration[name]
        out[name] = np.array(v) if (isinstance(v, list) and len(v) == 3) else v
    return out


def color_mask_cv(hsv, cfg, color):
    """line_detector_lsd.py:38-54: inRange (red = OR of two ranges) then dilate with the k x k ellipse."""
    if color == 'white':
        raw = cv2.inRange(hsv, cfg['hsv_white1'], cfg['hsv_white2'])
    elif color == 'yellow':
        raw = cv2.inRange(hsv, cfg['hsv_yellow1'], cfg['hsv_yellow2'])
    elif color == 'red':
        raw = cv2.bitwise_or(cv2.inRange(hsv, cfg['hsv_red1'], cfg['hsv_red2']),
                             cv2.inRange(hsv, cfg['hsv_red3'], cfg['hsv_red4']))
    else:
        raise Exception('Error: Undefined color strings...')
    k = cfg['dilation_kernel_size']
    return cv2.dilate(raw, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)))


def lsd_lines_cv(edge_img):
    """line_detector_lsd.py:64-72.  OpenCV 4 renamed the `_refine` kwarg, so it is passed positionally.
    Returns float32 [S,4] or the plain list [] when LSD finds nothing."""
    found = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV).detect(edge_img)[0]
    return [] if found is None else np.array(found[:, 0])


def normals_and_ordering(bw, lines):
    """line_detector_lsd.py:74-125 (_checkBounds, _correctPixelOrdering, _findNormal), keeping numpy's
    dtype promotions: lengths / unit vectors / centres in float32, probe pixels truncated toward zero,
    normals promoted to float64 by the int64 sign, endpoints swapped in place."""
    if len(lines) == 0:
        return [], []
    p1, p2 = lines[:, 0:2], lines[:, 2:4]
    length = np.sum((p1 - p2) ** 2, axis=1, keepdims=True) ** 0.5
    dx = 1. * (lines[:, 3:4] - lines[:, 1:2]) / length
    dy = 1. * (lines[:, 0:1] - lines[:, 2:3]) / length
    centers = np.hstack([(lines[:, 0:1] + lines[:, 2:3]) / 2, (lines[:, 1:2] + lines[:, 3:4]) / 2])
    h, w = bw.shape[:2]

    def probe(sign):
        px = np.clip((centers[:, 0:1] + sign * 3. * dx).astype('int'), 0, w - 1)
        py = np.clip((centers[:, 1:2] + sign * 3. * dy).astype('int'), 0, h - 1)
        return bw[py, px]

    # "x3 = c - 3.*dx": written as c + (-1)*3.*dx above; (-3.)*dx == -(3.*dx) exactly in IEEE
    flag_signs = np.logical_and(probe(-1.) > 0, probe(+1.) == 0).astype('int') * 2 - 1
    normals = np.hstack([dx, dy]) * flag_signs
    swap = ((lines[:, 2] - lines[:, 0]) * normals[:, 1] - (lines[:, 3] - lines[:, 1]) * normals[:, 0]) > 0
    lines[swap] = lines[swap][:, [2, 3, 0, 1]]
    return centers, normals


def hough_lines_cv(edge_img, threshold, min_line_length, max_line_gap):
    """line_detector1.py:63-69 (_HoughLine): int32 [S,4], or the plain list [] when nothing is found."""
    lines = cv2.HoughLinesP(edge_img, 1, np.pi / 180, threshold, np.empty(1), min_line_length, max_line_gap)
    return [] if lines is None else np.array(lines[:, 0])


class LineDetectorHSV:
    """Same plugin contract as src/line_detector/include/line_detector/line_detector1.py:11-137 (the detector eight of the ten
    shipped YAML files select): setImage and the colour filter are the LSD detector's, the lines come from cv2.HoughLinesP and are
    int32, so _findNormal's arithmetic (the same source lines as in line_detector_lsd.py) runs in float64 here."""

    def __init__(self, configuration):
        self.cfg = check_configuration(configuration)
        self.hough = tuple(int(configuration[k]) for k in ('hough_threshold', 'hough_min_line_length', 'hough_max_line_gap'))
        self.bgr = self.hsv = self.edges = np.empty(0)

    def setImage(self, bgr):  # :127-130
        self.bgr = np.copy(bgr)
        self.hsv = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
        lo, hi = self.cfg['canny_thresholds']
        self.edges = cv2.Canny(self.bgr, lo, hi, apertureSize=3)

    def detectLines(self, color):  # :121-125
        bw = color_mask_cv(self.hsv, self.cfg, color)
        edge_color = cv2.bitwise_and(bw, self.edges)
        lines = hough_lines_cv(edge_color, *self.hough)
        centers, normals = normals_and_ordering(bw, lines)
        return Detections(lines=lines, normals=normals, area=bw, centers=centers)

    def getImage(self):
        return self.bgr


class LineDetectorLSD:
    """Same plugin contract as src/line_detector/include/line_detector/line_detector_lsd.py:11-142
    (setImage / detectLines / getImage), built from the cv2 calls above."""

    def __init__(self, configuration):
        self.cfg = check_configuration(configuration)
        self.bgr = self.hsv = self.edges = np.empty(0)

    def setImage(self, bgr):  # :135-139 (the BGR2GRAY result is never used by the reference)
        self.bgr = np.copy(bgr)
        self.hsv = cv2.cvtColor(bgr, cv2.COLOR_BGR2HSV)
        lo, hi = self.cfg['canny_thresholds']
        self.edges = cv2.Canny(self.bgr, lo, hi, apertureSize=3)  # :60-62, 3-channel input

    def detectLines(self, color):  # :127-133
        bw = color_mask_cv(self.hsv, self.cfg, color)
        edge_color = cv2.bitwise_and(bw, self.edges)  # :56
        lines = lsd_lines_cv(edge_color)
        centers, normals = normals_and_ordering(bw, lines)
        return Detections(lines=lines, normals=normals, area=bw, centers=centers)

    def getImage(self):
        return self.bgr


def scaleandshift2(img, scale, shift):
    """src/anti_instagram/include/anti_instagram/scale_and_shift.py:25-33."""
    img_shift = np.zeros(img.shape, dtype='float32')
    for i in range(3):
        s = np.array(scale[i]).astype('float32')
        p = np.array(shift[i]).astype('float32')
        np.multiply(img[:, :, i], s, out=img_shift[:, :, i])
        img_shift[:, :, i] += p
    return img_shift


def preprocess(image_cv, img_size, top_cutoff, scale=(1, 1, 1), shift=(0, 0, 0)):
    """src/line_detector/src/line_detector_node.py:163-175 (resize, crop, AntiInstagram, convertScaleAbs)."""
    hei_original, wid_original = image_cv.shape[0:2]
    if img_size[0] != hei_original or img_size[1] != wid_original:
        image_cv = cv2.resize(image_cv, (img_size[1], img_size[0]), interpolation=cv2.INTER_NEAREST)
    image_cv = image_cv[top_cutoff:, :, :]
    image_cv_corr = scaleandshift2(image_cv, scale, shift)
    return cv2.convertScaleAbs(image_cv_corr)


def detect_frame(image_cv, detector, img_size, top_cutoff, scale=(1, 1, 1), shift=(0, 0, 0)):
    """line_detector_node.py:141-213 minus ROS: returns dict with per-colour Detections and the
    flattened SegmentList arrays (white, yellow, red order; LSD order within a colour):
    color u8[S], pixels_normalized f32[S,4], normal f32[S,2]; also lines_px f32[S,4] (cropped-image px)."""
    img = preprocess(image_cv, img_size, top_cutoff, scale, shift)
    detector.setImage(img)
    dets = [detector.detectLines(c) for c in COLORS]
    arr_cutoff = np.array((0, top_cutoff, 0, top_cutoff))
    arr_ratio = np.array((1. / img_size[1], 1. / img_size[0], 1. / img_size[1], 1. / img_size[0]))
    col, pixn, nrm, lpx = [], [], [], []
    for ci, d in enumerate(dets):
        if len(d.lines) > 0:
            ln = (d.lines + arr_cutoff) * arr_ratio  # float64
            # toSegmentMsg (:251-265): Vector2D fields are float32 on the wire
            pixn.append(ln.astype(np.float32))
            nrm.append(np.asarray(d.normals).astype(np.float32))
            col.append(np.full(len(d.lines), ci, np.uint8))
            lpx.append(np.asarray(d.lines, np.float32))
    if col:
        out = dict(color=np.concatenate(col), pixels_normalized=np.concatenate(pixn), normal=np.concatenate(nrm),
                   lines_px=np.concatenate(lpx))
    else:
        out = dict(color=np.zeros(0, np.uint8), pixels_normalized=np.zeros((0, 4), np.float32),
                   normal=np.zeros((0, 2), np.float32), lines_px=np.zeros((0, 4), np.float32))
    out["detections"] = dets
    out["image"] = img
    out["counts"] = [len(d.lines) for d in dets]
    return out


class GroundProjection:
    """src/ground_projection/include/ground_projection/GroundProjection.py:38-78 with
    image_geometry.PinholeCameraModel.rectifyPoint == cv2.undistortPoints(uv, K, D, R=R, P=P)."""

    def __init__(self, camera=None, homography=None):
        camera = camera or DEFAULT_CAMERA
        self.cw, self.ch = camera["width"], camera["height"]
        self.K = np.array(camera["K"], np.float64).reshape(3, 3)
        self.D = np.array(camera["D"], np.float64)
        self.R = np.array(camera["R"], np.float64).reshape(3, 3)
        self.P = np.array(camera["P"], np.float64).reshape(3, 4)
        self.H = np.array(homography or DEFAULT_HOMOGRAPHY, np.float64).reshape(3, 3)

    def vector2pixel(self, x, y):  # :38-48
        u = self.cw * x
        v = self.ch * y
        if u < 0: u = 0
        if u > self.cw - 1: u = self.cw - 1
        if v < 0: v = 0
        if v > self.ch - 1: v = 0
        return u, v

    def pixel2ground(self, u, v):  # :64-78
        uv = cv2.undistortPoints(np.array([[[u, v]]], np.float64), self.K, self.D, R=self.R, P=self.P)[0, 0]
        g = np.dot(self.H, np.array([uv[0], uv[1], 1.0]))
        return g[0] / g[2], g[1] / g[2]

    def vector2ground(self, x, y):  # :56-58
        return self.pixel2ground(*self.vector2pixel(float(x), float(y)))

    def project_segments(self, pixels_normalized):
        """ground_projection_node.py:55-65 -> points f64 [S,4] (x1,y1,x2,y2), z = 0."""
        out = np.zeros((len(pixels_normalized), 4), np.float64)
        for i, p in enumerate(pixels_normalized):
            out[i, 0:2] = self.vector2ground(p[0], p[1])
            out[i, 2:4] = self.vector2ground(p[2], p[3])
        return out


# src/line_sanity/src/line_sanity_node.py:17-23
LANEWIDTH, LINEWIDTH_WHITE, LINEWIDTH_YELLOW = 0.23, 0.05, 0.025
D_MIN, D_MAX, PHI_MIN, PHI_MAX = -0.15, 0.3, -1.5, 1.5


def fancy_filters(p1, p2, color):
    """line_sanity_node.py:75-117."""
    state = 0
    p1 = np.array(p1); p2 = np.array(p2)
    with np.errstate(all="ignore"):
        t_hat = (p2 - p1) / np.linalg.norm(p2 - p1)
        n_hat = np.array([-t_hat[1], t_hat[0]])
        d1 = np.inner(n_hat, p1); d2 = np.inner(n_hat, p2)
        l1 = np.inner(t_hat, p1); l2 = np.inner(t_hat, p2)
        if l1 < 0: l1 = -l1
        if l2 < 0: l2 = -l2
        l_i = (l1 + l2) / 2
        d_i = (d1 + d2) / 2
        phi_i = np.arcsin(t_hat[1])
    if color == WHITE:
        if p1[0] > p2[0]:
            d_i = d_i - LINEWIDTH_WHITE; state = 1
        else:
            d_i = -d_i; phi_i = -phi_i; state = 2
        d_i = d_i - LANEWIDTH / 2
    elif color == YELLOW:
        if p2[0] > p1[0]:
            d_i = d_i - LINEWIDTH_YELLOW; phi_i = -phi_i; state = 3
        else:
            d_i = -d_i; state = 4
        d_i = LANEWIDTH / 2 - d_i
    return d_i, phi_i, l_i, state


def sanity_keep(points, colors):
    """line_sanity_node.py:48-72 -> bool keep mask (order preserved)."""
    keep = np.zeros(len(points), bool)
    for i, (g, c) in enumerate(zip(points, colors)):
        if g[0] < 0 or g[2] < 0:
            continue
        if c != WHITE and c != YELLOW:
            continue
        d_i, phi_i, l_i, state = fancy_filters(g[0:2], g[2:4], c)
        if state == 0:
            continue
        if d_i > D_MAX or d_i < D_MIN or phi_i < PHI_MIN or phi_i > PHI_MAX:
            continue
        keep[i] = True
    return keep


def lane_filter_votes(points, colors, delta_d=0.02, delta_phi=0.1):
    """Vote counts of LaneFilterHistogram.generate_measurement_likelihood (src/lane_filter/include/lane_filter/
    lane_filter.py:82-102; generateVote :123-155 has the arithmetic of fancyFilters) for the ground segments of ONE
    frame -> int array [23, 30] for the default grid np.mgrid[d_min:d_max:delta_d, phi_min:phi_max:delta_phi]
    (the measurement likelihood is counts / counts.sum(), or None when there is no vote)."""
    from math import floor
    d, _ = np.mgrid[D_MIN:D_MAX:delta_d, PHI_MIN:PHI_MAX:delta_phi]
    counts = np.zeros(d.shape, np.int64)
    for g, c in zip(points, colors):
        if c != WHITE and c != YELLOW:
            continue
        if g[0] < 0 or g[2] < 0:
            continue
        d_i, phi_i, l_i, state = fancy_filters(g[0:2], g[2:4], c)
        if d_i > D_MAX or d_i < D_MIN or phi_i < PHI_MIN or phi_i > PHI_MAX:
            continue
        i = int(floor((d_i - D_MIN) / delta_d))
        j = int(floor((phi_i - PHI_MIN) / delta_phi))
        counts[i, j] += 1
    return counts


def front_end_frame(image_cv, detector, gp, img_size, top_cutoff, scale=(1, 1, 1), shift=(0, 0, 0)):
    """detector -> ground_projection -> line_sanity for one frame (show_map_complete.launch:10-43 chain)."""
    r = detect_frame(image_cv, detector, img_size, top_cutoff, scale, shift)
    r["ground"] = gp.project_segments(r["pixels_normalized"])
    r["keep"] = sanity_keep(r["ground"], r["color"])
    return r


def knn_hamming_bf(q, m, k):
    """Stand-in for the un-compilable Mihasher: cv2.BFMatcher(NORM_HAMMING).knnMatch (exact, ties ->
    smallest train index)."""
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    res = bf.knnMatch(np.ascontiguousarray(q), np.ascontiguousarray(m), k=k)
    idx = np.full((len(q), k), -1, np.int32)
    dist = np.full((len(q), k), -1, np.int32)
    for i, row in enumerate(res):
        for j, dm in enumerate(row):
            idx[i, j] = dm.trainIdx
            dist[i, j] = int(dm.distance)
    return idx, dist


def map_transform(ground, frame_ids, poses, pose_frame_base=0):
    """What RViz does with show_map's markers (frame '/duck', src/show_map/src/show_map.py:45-75) and the map->duck transform
    the odometry node broadcasts (src/odometry/src/odometry.py:110-120: translation (x, y, 0), rotation theta about z):
    p_map = R(theta) p_duck + (x, y), per segment with the pose of ITS frame.  ground f64 [S,4] -> f64 [S,4]."""
    ground = np.asarray(ground, np.float64).reshape(-1, 4)
    out = ground.copy()
    if poses is None:
        return out
    poses = np.asarray(poses, np.float64).reshape(-1, 3)
    for i, f in enumerate(np.asarray(frame_ids)):
        pf = int(f) - pose_frame_base
        if pf < 0 or pf >= len(poses):
            continue
        x, y, th = poses[pf]
        c, s = np.cos(th), np.sin(th)
        for o in (0, 2):
            px, py = ground[i, o], ground[i, o + 1]
            out[i, o] = (c * px - s * py) + x
            out[i, o + 1] = (s * px + c * py) + y
    return out


class LaneFilterHistogram(object):
    """src/lane_filter/include/lane_filter/lane_filter.py:12-120 restated (same numpy / scipy calls, no ROS): initialize,
    predict, update (votes via lane_filter_votes above), getEstimate, getMax."""
    DEFAULTS = dict(mean_d_0=0, mean_phi_0=0, sigma_d_0=0.1, sigma_phi_0=0.1, delta_d=0.02, delta_phi=0.1, d_max=0.3, d_min=-0.15,
                    phi_min=-1.5, phi_max=1.5, cov_v=0.5, linewidth_white=0.05, linewidth_yellow=0.025, lanewidth=0.23, min_max=0.1,
                    sigma_d_mask=1.0, sigma_phi_mask=2.0)

    def __init__(self, configuration=None):
        for k, v in dict(self.DEFAULTS, **(configuration or {})).items():
            setattr(self, k, v)
        self.d, self.phi = np.mgrid[self.d_min:self.d_max:self.delta_d, self.phi_min:self.phi_max:self.delta_phi]   # :38
        self.mean_0 = [self.mean_d_0, self.mean_phi_0]
        self.cov_0 = [[self.sigma_d_0, 0], [0, self.sigma_phi_0]]
        self.cov_mask = [self.sigma_d_mask, self.sigma_phi_mask]
        self.initialize()

    def initialize(self):  # :114-120
        from scipy.stats import multivariate_normal
        pos = np.empty(self.d.shape + (2,))
        pos[:, :, 0] = self.d
        pos[:, :, 1] = self.phi
        self.belief = multivariate_normal(self.mean_0, self.cov_0).pdf(pos)

    def predict(self, dt, v, w):  # :47-72
        from math import floor
        from scipy.ndimage import gaussian_filter
        d_t = self.d + v * dt * np.sin(self.phi)
        phi_t = self.phi + w * dt
        p_belief = np.zeros(self.belief.shape)
        for i in range(self.belief.shape[0]):
            for j in range(self.belief.shape[1]):
                if self.belief[i, j] > 0:
                    if d_t[i, j] > self.d_max or d_t[i, j] < self.d_min or phi_t[i, j] < self.phi_min or phi_t[i, j] > self.phi_max:
                        continue
                    i_new = int(floor((d_t[i, j] - self.d_min) / self.delta_d))
                    j_new = int(floor((phi_t[i, j] - self.phi_min) / self.delta_phi))
                    p_belief[i_new, j_new] += self.belief[i, j]
        s_belief = np.zeros(self.belief.shape)
        gaussian_filter(p_belief, self.cov_mask, output=s_belief, mode='constant')
        if np.sum(s_belief) == 0:
            return
        self.belief = s_belief / np.sum(s_belief)

    def update(self, ground, colors):  # :75-102
        counts = lane_filter_votes(ground, colors, self.delta_d, self.delta_phi).astype(np.float64)
        if np.linalg.norm(counts) == 0:
            return None
        ml = counts / np.sum(counts)
        self.belief = np.multiply(self.belief, ml)
        if np.sum(self.belief) == 0:
            self.belief = ml
        else:
            self.belief = self.belief / np.sum(self.belief)
        return ml

    def getEstimate(self):  # :104-109
        maxids = np.unravel_index(self.belief.argmax(), self.belief.shape)
        return [self.d_min + (maxids[0] + 0.5) * self.delta_d, self.phi_min + (maxids[1] + 0.5) * self.delta_phi]

    def getMax(self):
        return self.belief.max()
