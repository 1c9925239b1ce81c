"""CPU ORACLE (test infrastructure): ctypes bindings of csrc/lane_oracle.c.

cv2-free, plain-C models of every primitive on the path; used for stage-level goldens of each CUDA
kernel and as the scalar "port" CPU baseline for LBD (which has no runnable reference).
"""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.orc_nfa.restype = C.c_double
        _lib.orc_nfa.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        _lib.orc_fast_atan2.restype = C.c_float
        _lib.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        _lib.orc_lsd_detect.restype = C.c_int
        _lib.orc_sanity_keep.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


def preprocess(src, img_size, top_cutoff, scale=(1, 1, 1), shift=(0, 0, 0)):
    src = _u8(src)
    sh, sw = src.shape[:2]
    dh, dw = img_size
    out = np.empty((dh - top_cutoff, dw, 3), np.uint8)
    sc = np.asarray(scale, np.float32); sf = np.asarray(shift, np.float32)
    lib().orc_preprocess(_p(src), sh, sw, dh, dw, top_cutoff, _p(sc), _p(sf), _p(out))
    return out


def bgr2hsv(bgr):
    bgr = _u8(bgr)
    out = np.empty_like(bgr)
    lib().orc_bgr2hsv(_p(bgr), C.c_size_t(bgr.size // 3), _p(out))
    return out


def bgr2gray(bgr):
    bgr = _u8(bgr)
    out = np.empty(bgr.shape[:-1], np.uint8)
    lib().orc_bgr2gray(_p(bgr), C.c_size_t(bgr.size // 3), _p(out))
    return out


def hsv_bounds(cfg):
    lo = np.array([cfg['hsv_white1'], cfg['hsv_yellow1'], cfg['hsv_red1'], cfg['hsv_red3']], np.int32)
    hi = np.array([cfg['hsv_white2'], cfg['hsv_yellow2'], cfg['hsv_red2'], cfg['hsv_red4']], np.int32)
    return np.ascontiguousarray(lo), np.ascontiguousarray(hi)


def color_mask(hsv, cfg, colour_idx):
    hsv = _u8(hsv)
    H, W = hsv.shape[:2]
    lo, hi = hsv_bounds(cfg)
    out = np.empty((H, W), np.uint8)
    lib().orc_color_mask(_p(hsv), H, W, _p(lo), _p(hi), colour_idx, _p(out))
    return out


def dilate(img, ksize=3):
    img = _u8(img)
    H, W = img.shape
    out = np.empty_like(img)
    lib().orc_dilate(_p(img), H, W, int(ksize), _p(out))
    return out


def canny_bgr(bgr, lo, hi):
    bgr = _u8(bgr)
    H, W = bgr.shape[:2]
    edges = np.empty((H, W), np.uint8)
    nms = np.empty((H, W), np.uint8)
    lib().orc_canny_bgr(_p(bgr), H, W, int(lo), int(hi), _p(edges), _p(nms))
    return edges, nms


def lsd_scaled_size(H, W):
    sh, sw = C.c_int(), C.c_int()
    lib().orc_lsd_scaled_size(H, W, C.byref(sh), C.byref(sw))
    return sh.value, sw.value


def lsd_blur_resize(img):
    img = _u8(img)
    H, W = img.shape
    sh, sw = lsd_scaled_size(H, W)
    out = np.empty((sh, sw), np.uint8)
    lib().orc_lsd_blur_resize(_p(img), H, W, _p(out))
    return out


def lsd_detect(img, refine=2, cap=8192, stages=False):
    """-> lines f32 [S,4], extra f64 [S,3] (width, p, log_nfa) [, scaled u8, ang_deg f32]."""
    img = _u8(img)
    H, W = img.shape
    lines = np.empty((cap, 4), np.float32)
    extra = np.empty((cap, 3), np.float64)
    sh, sw = lsd_scaled_size(H, W)
    scaled = np.empty((sh, sw), np.uint8) if stages else None
    ang = np.empty((sh, sw), np.float32) if stages else None
    n = lib().orc_lsd_detect(_p(img), H, W, int(refine), _p(lines), _p(extra), cap, _p(scaled), _p(ang))
    if n > cap:
        return lsd_detect(img, refine, cap=n, stages=stages)
    if stages:
        return lines[:n].copy(), extra[:n].copy(), scaled, ang
    return lines[:n].copy(), extra[:n].copy()


def find_normals(bw, lines):
    """in-place endpoint swap on a copy; -> lines f32, normals f64, centers f32."""
    bw = _u8(bw)
    H, W = bw.shape
    lines = np.array(lines, np.float32).reshape(-1, 4).copy()
    n = len(lines)
    normals = np.empty((n, 2), np.float64)
    centers = np.empty((n, 2), np.float32)
    lib().orc_find_normals(_p(bw), H, W, _p(lines), n, _p(normals), _p(centers))
    return lines, normals, centers


def _cam_arrays(camera, homography):
    K = np.ascontiguousarray(camera["K"], np.float64); D = np.ascontiguousarray(camera["D"], np.float64)
    R = np.ascontiguousarray(camera["R"], np.float64); P = np.ascontiguousarray(camera["P"], np.float64)
    Hg = np.ascontiguousarray(homography, np.float64)
    return K, D, R, P, Hg


def undistort_points(uv, camera):
    K, D, R, P, _ = _cam_arrays(camera, np.zeros(9))
    uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
    out = np.empty_like(uv)
    lib().orc_undistort_points(_p(uv), len(uv), _p(K), _p(D), _p(R), _p(P), _p(out))
    return out


def project_filter(lines_px, colour, img_size, top_cutoff, camera, homography):
    """-> pixels_normalized f32 [S,4], ground f64 [S,4], keep u8 [S]."""
    K, D, R, P, Hg = _cam_arrays(camera, homography)
    lines_px = np.ascontiguousarray(lines_px, np.float32).reshape(-1, 4)
    colour = _u8(colour)
    n = len(lines_px)
    pixn = np.empty((n, 4), np.float32); ground = np.empty((n, 4), np.float64); keep = np.empty(n, np.uint8)
    lib().orc_project_filter(_p(lines_px), _p(colour), n, int(img_size[0]), int(img_size[1]), int(top_cutoff),
                             _p(K), _p(D), _p(R), _p(P), int(camera["width"]), int(camera["height"]), _p(Hg),
                             _p(pixn), _p(ground), _p(keep))
    return pixn, ground, keep


def gauss5_sobel(gray):
    gray = _u8(gray)
    H, W = gray.shape
    blur = np.empty((H, W), np.uint8); dx = np.empty((H, W), np.int16); dy = np.empty((H, W), np.int16)
    lib().orc_gauss5_sobel(_p(gray), H, W, _p(blur), _p(dx), _p(dy))
    return blur, dx, dy


def lbd(lines_px, dx, dy):
    """-> keylines f32 [S,8], desc72 f32 [S,72], desc32 u8 [S,32]."""
    lines_px = np.ascontiguousarray(lines_px, np.float32).reshape(-1, 4)
    dx = np.ascontiguousarray(dx, np.int16); dy = np.ascontiguousarray(dy, np.int16)
    H, W = dx.shape
    n = len(lines_px)
    kls = np.empty((n, 8), np.float32); d72 = np.empty((n, 72), np.float32); d32 = np.empty((n, 32), np.uint8)
    lib().orc_lbd(_p(lines_px), n, _p(dx), _p(dy), H, W, _p(kls), _p(d72), _p(d32))
    return kls, d72, d32


def knn_hamming(q, m, k, max_dist=256):
    q = _u8(q).reshape(-1, 32); m = _u8(m).reshape(-1, 32)
    idx = np.empty((len(q), k), np.int32); dist = np.empty((len(q), k), np.int32)
    lib().orc_knn_hamming(_p(q), len(q), _p(m), len(m), int(k), int(max_dist), _p(idx), _p(dist))
    return idx, dist


COLORS = ("white", "yellow", "red")


def front_end_frame(image, cfg, img_size, top_cutoff, camera, homography, scale=(1, 1, 1), shift=(0, 0, 0),
                    descriptors=False):
    """The whole per-frame path in plain C (no cv2): the stage-level twin of reference_glue.front_end_frame."""
    img = preprocess(image, img_size, top_cutoff, scale, shift)
    hsv = bgr2hsv(img)
    edges, nms = canny_bgr(img, cfg['canny_thresholds'][0], cfg['canny_thresholds'][1])
    out = dict(image=img, hsv=hsv, edges=edges, nms=nms, bw=[], edge_color=[], counts=[])
    L, Nn, Cc, col = [], [], [], []
    for ci in range(3):
        bw = dilate(color_mask(hsv, cfg, ci), cfg['dilation_kernel_size'])
        ec = bw & edges
        lines, _ = lsd_detect(ec)
        lines, normals, centers = find_normals(bw, lines)
        out["bw"].append(bw); out["edge_color"].append(ec); out["counts"].append(len(lines))
        L.append(lines); Nn.append(normals); Cc.append(centers); col.append(np.full(len(lines), ci, np.uint8))
    out["lines_px"] = np.concatenate(L); out["normal64"] = np.concatenate(Nn)
    out["normal"] = out["normal64"].astype(np.float32)
    out["centers"] = np.concatenate(Cc); out["color"] = np.concatenate(col)
    pixn, ground, keep = project_filter(out["lines_px"], out["color"], img_size, top_cutoff, camera, homography)
    out["pixels_normalized"] = pixn; out["ground"] = ground; out["keep"] = keep.astype(bool)
    if descriptors:
        gray = bgr2gray(img)
        blur, dx, dy = gauss5_sobel(gray)
        kls, d72, d32 = lbd(out["lines_px"], dx, dy)
        out.update(gray=gray, dx=dx, dy=dy, keylines=kls, desc72=d72, desc32=d32)
    return out


def knn_mihasher(q, m, k):
    """knnMatch in the reference's own (Mihasher) result order: distance, then hash-discovery order; D = 128."""
    q = _u8(q).reshape(-1, 32); m = _u8(m).reshape(-1, 32)
    idx = np.empty((len(q), k), np.int32); dist = np.empty((len(q), k), np.int32)
    lib().orc_knn_mihasher(_p(q), len(q), _p(m), len(m), int(k), _p(idx), _p(dist))
    return idx, dist


def jpeg_decode(data):
    """Baseline JPEG -> BGR uint8 [H,W,3] exactly like cv2.imdecode(data, cv2.IMREAD_COLOR) (libjpeg-turbo defaults: islow
    IDCT, fancy upsampling); duckietown_utils/jpg.py:21-31.  Raises ValueError for streams outside the restated scope."""
    buf = np.ascontiguousarray(np.frombuffer(bytes(data), np.uint8) if not isinstance(data, np.ndarray) else data, np.uint8)
    W, H, nc = C.c_int(), C.c_int(), C.c_int()
    rc = lib().orc_jpeg_info(_p(buf), C.c_size_t(len(buf)), C.byref(W), C.byref(H), C.byref(nc))
    if rc != 0:
        raise ValueError("not a baseline JPEG (%d)" % rc)
    out = np.empty((H.value, W.value, 3), np.uint8)
    rc = lib().orc_jpeg_decode_bgr(_p(buf), C.c_size_t(len(buf)), _p(out))
    if rc != 0:
        raise ValueError("JPEG decode failed (%d)" % rc)
    return out
