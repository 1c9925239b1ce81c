"""Developer tool (not a test): run the CUDA path on a few frames and print stage-by-stage parity
against the C oracle, plus per-kernel timings.  Usage on the GPU box:  python tests/gpu_debug.py [n] [H W]"""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import cmodel as cm, reference_glue as rg, synth  # noqa: E402
import lane_slam_b200 as L  # noqa: E402


def section(name):
    print("\n=== %s ===" % name, flush=True)


def compare_frames(frames, isz, cut, dense=False, scale=(1, 1, 1), shift=(0, 0, 0), describe=True):
    n, H, W = frames.shape[:3]
    cam, Hg = rg.scaled_camera(W, H) if (W, H) != (640, 480) else (rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY)
    cfg = rg.check_configuration(dict(rg.DEFAULT_DETECTOR_CONFIG))
    fe = L.FrontEnd(dict(L.DEFAULT_DETECTOR_CONFIGURATION), img_size=isz, top_cutoff=cut, camera=cam, homography=Hg,
                    src_size=(H, W), max_batch=n, ai_scale=scale, ai_shift=shift, max_segments_per_frame=4096)
    stages = L.STAGE_DETECT | L.STAGE_GROUND | (L.STAGE_DESCRIBE if describe else 0)
    t = time.time()
    b = fe.process(frames, stages=stages)
    print("process: %.1f ms, S=%d, launches=%d" % ((time.time() - t) * 1e3, b.n_segments, fe.launch_count()))
    print("timings:", ["%s=%.3f" % x for x in fe.timings()])
    nbad = 0
    for f in range(n):
        o = cm.front_end_frame(frames[f], cfg, isz, cut, cam, Hg, scale, shift, descriptors=describe)
        res = {}
        res["image"] = np.array_equal(fe.tap("image", f), o["image"])
        lab = fe.tap("labels", f)
        raw = [(cm.color_mask(o["hsv"], cfg, c) > 0) for c in range(3)]
        res["masks"] = all(np.array_equal(((lab >> c) & 1).astype(bool), raw[c]) for c in range(3))
        res["nms"] = np.array_equal(lab >> 4, o["nms"])
        res["edges"] = np.array_equal(fe.tap("edges", f), o["edges"])
        res["bw"] = all(np.array_equal(fe.tap("bw_" + c, f), o["bw"][i]) for i, c in enumerate(L.COLORS))
        res["ec"] = all(np.array_equal(fe.tap("ec_" + c, f), o["edge_color"][i]) for i, c in enumerate(L.COLORS))
        g = b.frame(f)
        res["counts"] = g["counts"] == o["counts"]
        if res["counts"]:
            res["lines_exact"] = np.array_equal(g["lines_px"], o["lines_px"])
            res["lines_err"] = float(np.abs(g["lines_px"] - o["lines_px"]).max()) if len(o["lines_px"]) else 0.0
            res["normals"] = np.array_equal(g["normals"], o["normal64"])
            res["pixn"] = np.array_equal(g["pixels_normalized"], o["pixels_normalized"])
            res["ground_err"] = float(np.abs(g["ground"] - o["ground"]).max()) if len(o["ground"]) else 0.0
            res["keep"] = np.array_equal(g["keep"], o["keep"])
            if describe:
                res["gray"] = np.array_equal(fe.tap("gray", f), o["gray"])
                res["dx"] = np.array_equal(fe.tap("dx", f), o["dx"]) and np.array_equal(fe.tap("dy", f), o["dy"])
                dm = (g["desc"] != o["desc32"])
                res["desc_rows_bad"] = int(dm.any(axis=1).sum())
                res["desc_bits_bad"] = int(np.unpackbits(g["desc"] ^ o["desc32"]).sum())
        ok = all(v is True or (isinstance(v, float) and v < 1e-4) or (isinstance(v, int) and not isinstance(v, bool) and v == 0)
                 for v in res.values())
        nbad += not ok
        print("frame %d: %s %s gpu=%s oracle=%s" % (f, "OK " if ok else "BAD", res, g["counts"], o["counts"]), flush=True)
        if not res["counts"] or not res.get("lines_exact", True):
            # dump the first differing segment per colour
            off_g = np.cumsum([0] + g["counts"]); off_o = np.cumsum([0] + o["counts"])
            for c in range(3):
                lg = g["lines_px"][off_g[c]:off_g[c + 1]]; lo = o["lines_px"][off_o[c]:off_o[c + 1]]
                m = min(len(lg), len(lo))
                # oracle lines are post-swap; compare unordered endpoints
                for i in range(m):
                    if not (np.allclose(lg[i], lo[i], atol=1e-3) or np.allclose(lg[i], lo[i][[2, 3, 0, 1]], atol=1e-3)):
                        print("   colour %d first diff at %d: gpu %s oracle %s" % (c, i, lg[i], lo[i]))
                        break
    fe.close()
    return nbad


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    H, W = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (480, 640)
    bad = 0
    for name, fn in [
        ("native cut0", lambda: compare_frames(np.stack([synth.frame(s, H, W) for s in range(n)]), (H, W), 0)),
        ("resize 120x160 cut40", lambda: compare_frames(np.stack([synth.frame(s, H, W) for s in range(n)]), (120, 160), 40)),
        ("native cut H/3 + colour transform", lambda: compare_frames(
            np.stack([synth.frame(s + 10, H, W) for s in range(n)]), (H, W), H // 3, scale=(1.1, 0.93, 1.27), shift=(3.5, -7.25, 12.0))),
        ("dense", lambda: compare_frames(np.stack([synth.frame(s, H, W, dense=True) for s in range(2)]), (H, W), 0)),
        ("odd size 123x161", lambda: compare_frames(np.stack([synth.frame(s, 123, 161) for s in range(2)]), (123, 161), 0)),
    ]:
        section(name)
        try:
            bad += fn()
        except Exception:
            traceback.print_exc()
            bad += 1
    section("knn")
    try:
        q, m, src = synth.descriptor_sets(500, 20000)
        fe = L.FrontEnd(max_batch=1)
        t = time.time()
        idx, dist = fe.knn(q, m, k=2)
        print("knn %.1f ms" % ((time.time() - t) * 1e3), fe.timings())
        oi, od = cm.knn_hamming(q, m, 2)
        print("knn idx equal", np.array_equal(idx, oi), "dist equal", np.array_equal(dist, od))
        bad += not (np.array_equal(idx, oi) and np.array_equal(dist, od))
        fe.close()
    except Exception:
        traceback.print_exc()
        bad += 1
    print("\nTOTAL BAD:", bad)


if __name__ == "__main__":
    main()
