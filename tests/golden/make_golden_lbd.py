"""Generate tests/golden/lbd_reference.npz with the reference's OWN compiled line_descriptor code
(oracle/_ref/libref_line_descriptor.so, built by oracle/build_ref.py from /root/reference/src/line_descriptor/src/*.cpp,
unmodified).  Run in the authoring container only.

For synthetic frames (oracle/synth.py seeds) and real frames (tests/golden/real_images.npz) the segment lists found by
the detector are handed to LSDDetectorC (KeyLine fill, LSDDetector_custom.cpp:130-215) and BinaryDescriptor::compute
(binary_descriptor_custom.cpp:524-687, computeLBD :1026-1372); stored: the input lines, the KeyLine fields, the 72-float
descriptor and the 32-byte binary descriptor.  Also BinaryDescriptorMatcher::knnMatch (binary_descriptor_matcher.cpp:258-335,
Mihasher) on a tie-heavy descriptor set."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import realset
    from oracle import cmodel as cm, reference_glue as rg, refnative as rn, synth
    cfg = rg.check_configuration(dict(rg.DEFAULT_DETECTOR_CONFIG))
    out = {}
    cases = [("synth", s, 480, 640, False) for s in range(4)] + [("synth", 5, 480, 640, True), ("synth", 2, 240, 320, False)] + \
            [("real", i, 480, 640, False) for i in range(0, realset.count(), 4)]
    out["cases"] = np.array([[0 if c[0] == "synth" else 1, c[1], c[2], c[3], int(c[4])] for c in cases], np.int32)
    nl = 0
    for k, (kind, idx, H, W, dense) in enumerate(cases):
        img = synth.frame(idx, H, W, dense=dense) if kind == "synth" else realset.image(idx)
        o = cm.front_end_frame(img, cfg, (H, W), 0, *rg.scaled_camera(W, H))
        gray = cv2.cvtColor(o["image"], cv2.COLOR_BGR2GRAY)
        kl, d72, d32 = rn.keylines_lbd(gray, o["lines_px"])
        out["%d_lines" % k] = o["lines_px"]; out["%d_keylines" % k] = kl; out["%d_desc72" % k] = d72; out["%d_desc32" % k] = d32
        nl += len(d32)
    # matcher: tie-heavy set (few distinct rows, small perturbations) + the standard planted set
    rng = np.random.default_rng(5)
    base = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    m = base[rng.integers(0, 40, 3000)].copy()
    flips = rng.random((3000, 256)) < 0.01
    m ^= np.packbits(flips, axis=1, bitorder="little")
    q = base[rng.integers(0, 40, 200)].copy()
    q ^= np.packbits(rng.random((200, 256)) < 0.02, axis=1, bitorder="little")
    out["knn_q"] = q; out["knn_m"] = m
    for k in (1, 2, 4, 8):
        i, d = rn.knn_match(q, m, k)
        out["knn_idx_%d" % k] = i; out["knn_dist_%d" % k] = d
    path = os.path.join(HERE, "lbd_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote %s: %d cases, %d lines, %.2f MB" % (path, len(cases), nl, os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    main()
