"""Generate tests/golden/real_images.npz: real Duckietown camera frames through the REFERENCE's own detector class.

Run in the authoring container only (needs /root/reference).  The reference ships 173 real 640x480 JPEGs
(src/anti_instagram/annotation-tool/images/, five recording sessions); this script takes a fixed spread of them,
decodes them exactly like the reference does (duckietown_utils/jpg.py:21-31 -> cv2.imdecode(..., cv2.IMREAD_COLOR))
and runs the UNMODIFIED LineDetectorLSD (imported from /root/reference by make_golden.reference_detector) on

  * the native 640x480 frame, top_cutoff 0               (the benchmark geometry),
  * the reference default 160x120 nearest resize, cut 40  (line_detector_node/default.yaml:1-2),
  * every distinct threshold set of the shipped YAML files (line_detector_node/*.yaml) at the default geometry.

Stored: the JPEG byte streams themselves (small; decoding on the GPU box uses the same cv2 wheel -- a CRC of the
decoded pixels is stored and checked), the Detections (lines f32, normals f64, centers f32) per colour, CRCs of the
`area` masks and the Canny map at native size and the packed masks at the small size.
"""
import glob
import os
import sys
import zlib

import cv2
import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (reference_detector(): loads the reference's classes)

IMG_DIR = "/root/reference/src/anti_instagram/annotation-tool/images"
YAML_DIR = "/root/reference/src/duckietown/config/baseline/line_detector/line_detector_node"
PER_FAMILY = 6
LSD_KEYS = ['hsv_white1', 'hsv_white2', 'hsv_yellow1', 'hsv_yellow2', 'hsv_red1', 'hsv_red2', 'hsv_red3', 'hsv_red4',
            'dilation_kernel_size', 'canny_thresholds', 'hough_threshold', 'hough_min_line_length', 'hough_max_line_gap']


def pick_images():
    fams = {}
    for p in sorted(glob.glob(IMG_DIR + "/*.jpg")):
        fams.setdefault(os.path.basename(p).rsplit("_", 1)[0], []).append(p)
    out = []
    for name in sorted(fams):
        fs = fams[name]
        idx = sorted(set(np.linspace(0, len(fs) - 1, min(PER_FAMILY, len(fs))).round().astype(int).tolist()))
        out += [fs[i] for i in idx]
    return out


def yaml_threshold_sets():
    """name -> configuration dict for LineDetectorLSD (exact key set), one per DISTINCT threshold set."""
    seen, out = {}, {}
    for p in sorted(glob.glob(YAML_DIR + "/*.yaml")):
        y = yaml.safe_load(open(p))
        conf = y["detector"][1]["configuration"]
        c = {k: conf.get(k, {"hough_threshold": 2, "hough_min_line_length": 3, "hough_max_line_gap": 1}.get(k)) for k in LSD_KEYS}
        key = repr([c[k] for k in LSD_KEYS if k.startswith("hsv") or k in ("canny_thresholds", "dilation_kernel_size")])
        if key in seen:
            continue
        seen[key] = 1
        out[os.path.basename(p)[:-5]] = c
    return out


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def store(out, prefix, det, full_maps):
    for c in ("white", "yellow", "red"):
        d = det.detectLines(c)
        n = len(d.lines)
        out["%s_%s_lines" % (prefix, c)] = np.asarray(d.lines, np.float32).reshape(n, 4)
        out["%s_%s_normals" % (prefix, c)] = np.asarray(d.normals, np.float64).reshape(n, 2)
        out["%s_%s_centers" % (prefix, c)] = np.asarray(d.centers, np.float32).reshape(n, 2)
        if full_maps:
            out["%s_%s_area" % (prefix, c)] = np.packbits(d.area > 0)
        else:
            out["%s_%s_area_crc" % (prefix, c)] = crc(np.packbits(d.area > 0))
    if full_maps:
        out["%s_edges" % prefix] = np.packbits(det.edges > 0)
    else:
        out["%s_edges_crc" % prefix] = crc(np.packbits(det.edges > 0))


def main():
    det, _ = mg.reference_detector()
    lsd_mod = sys.modules["line_detector.line_detector_lsd"]
    files = pick_images()
    blobs = [open(p, "rb").read() for p in files]
    out = {"names": np.array([os.path.basename(p) for p in files]),
           "jpeg_bytes": np.frombuffer(b"".join(blobs), np.uint8),
           "jpeg_offsets": np.cumsum([0] + [len(b) for b in blobs]).astype(np.int64)}
    crcs = []
    nseg = [0, 0]
    for i, b in enumerate(blobs):
        img = cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR)     # jpg.py:21-31
        assert img.shape == (480, 640, 3)
        crcs.append(crc(img))
        det.setImage(img)                                                    # native geometry
        store(out, "n%d" % i, det, full_maps=False)
        nseg[0] += sum(len(out["n%d_%s_lines" % (i, c)]) for c in ("white", "yellow", "red"))
        small = cv2.resize(img, (160, 120), interpolation=cv2.INTER_NEAREST)[40:, :, :]   # line_detector_node.py:163-169
        det.setImage(small)
        store(out, "d%d" % i, det, full_maps=True)
        nseg[1] += sum(len(out["d%d_%s_lines" % (i, c)]) for c in ("white", "yellow", "red"))
    out["decoded_crc"] = np.array(crcs, np.uint32)
    # shipped YAML threshold sets, default geometry, a few frames each
    sets = yaml_threshold_sets()
    out["yaml_names"] = np.array(sorted(sets))
    yaml_frames = list(range(0, len(blobs), 5))
    out["yaml_frames"] = np.array(yaml_frames, np.int32)
    for name in sorted(sets):
        conf = sets[name]
        out["yaml_%s_conf" % name] = np.array(yaml.safe_dump(conf))
        dy = lsd_mod.LineDetectorLSD(configuration=dict(conf))
        for i in yaml_frames:
            img = cv2.imdecode(np.frombuffer(blobs[i], np.uint8), cv2.IMREAD_COLOR)
            dy.setImage(cv2.resize(img, (160, 120), interpolation=cv2.INTER_NEAREST)[40:, :, :])
            store(out, "y_%s_%d" % (name, i), dy, full_maps=True)
    path = os.path.join(HERE, "real_images.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d images, %d native / %d default-geometry segments, %d yaml sets (%s), %.1f MB, cv2 %s" % (
        path, len(blobs), nseg[0], nseg[1], len(sets), ", ".join(sorted(sets)), os.path.getsize(path) / 1e6, cv2.__version__))


if __name__ == "__main__":
    main()
