"""CPU tests: pin the oracle.  (1) restated glue + C model vs the golden vectors produced by the REFERENCE's own
Python files (tests/golden/make_golden.py); (2) C model vs cv2 4.13 primitives, stage by stage; (3) known
answers for LBD / kNN restatements (no runnable reference: "parity unpinned" beyond self-consistency)."""
import os

import cv2
import numpy as np
import pytest

from oracle import cmodel as cm, reference_glue as rg, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_detections.npz")
CFG = rg.check_configuration(dict(rg.DEFAULT_DETECTOR_CONFIG))


def _prep(case):
    seed, H, W, ih, iw, cut = [int(v) for v in case]
    return synth.frame(seed, H, W), (ih, iw), cut


def test_golden_reference_detections_glue_and_cmodel():
    g = np.load(GOLD)
    det = rg.LineDetectorLSD(dict(rg.DEFAULT_DETECTOR_CONFIG))
    for i, case in enumerate(g["cases"]):
        img, isz, cut = _prep(case)
        proc = rg.preprocess(img, isz, cut)
        det.setImage(proc)
        c_out = cm.front_end_frame(img, CFG, isz, cut, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY)
        assert np.array_equal(np.packbits(det.edges > 0), g["%d_edges" % i])
        assert np.array_equal(np.packbits(c_out["edges"] > 0), g["%d_edges" % i])
        assert np.array_equal(c_out["hsv"].astype(np.int64).sum(axis=(0, 1)), g["%d_hsv_sum" % i])
        off = 0
        for ci, c in enumerate(rg.COLORS):
            d = det.detectLines(c)
            gl, gn, gc = g["%d_%s_lines" % (i, c)], g["%d_%s_normals" % (i, c)], g["%d_%s_centers" % (i, c)]
            n = len(gl)
            assert len(d.lines) == n
            assert np.array_equal(np.packbits(d.area > 0), g["%d_%s_area" % (i, c)])
            assert np.array_equal(np.packbits(c_out["bw"][ci] > 0), g["%d_%s_area" % (i, c)])
            if n:
                assert np.array_equal(np.asarray(d.lines), gl) and np.array_equal(np.asarray(d.normals), gn)
                assert np.array_equal(np.asarray(d.centers), gc)
            assert c_out["counts"][ci] == n
            assert np.array_equal(c_out["lines_px"][off:off + n], gl)
            assert np.array_equal(c_out["normal64"][off:off + n], gn)
            assert np.array_equal(c_out["centers"][off:off + n], gc)
            off += n


def test_golden_undistort():
    g = np.load(GOLD)
    out = cm.undistort_points(g["undist_uv"], rg.DEFAULT_CAMERA)
    assert np.abs(out - g["undist_out"]).max() < 1e-9


def test_hsv_all_colors_vs_cv2():
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([(v & 255), (v >> 8) & 255, (v >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(cm.bgr2hsv(img), cv2.cvtColor(img, cv2.COLOR_BGR2HSV))
    assert np.array_equal(cm.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("seed,shape", [(0, (480, 640)), (1, (123, 161)), (2, (240, 320))])
def test_canny_masks_vs_cv2(seed, shape):
    img = synth.frame(seed, *shape)
    edges, _ = cm.canny_bgr(img, 80, 200)
    assert np.array_equal(edges, cv2.Canny(img, 80, 200, apertureSize=3))
    hsv = cv2.cvtColor(img, cv2.COLOR_BGR2HSV)
    for ci, c in enumerate(rg.COLORS):
        assert np.array_equal(cm.dilate(cm.color_mask(hsv, CFG, ci), 3), rg.color_mask_cv(hsv, CFG, c))
    rng = np.random.default_rng(seed)
    noise = cv2.GaussianBlur(rng.integers(0, 256, shape + (3,), dtype=np.uint8), (5, 5), 1.5)
    e2, _ = cm.canny_bgr(noise, 50, 150)
    assert np.array_equal(e2, cv2.Canny(noise, 50, 150, apertureSize=3))


@pytest.mark.parametrize("seed,shape,dense", [(3, (480, 640), False), (4, (97, 203), False), (5, (240, 320), True)])
def test_lsd_vs_cv2(seed, shape, dense):
    img = synth.frame(seed, *shape, dense=dense)
    det = rg.LineDetectorLSD(dict(rg.DEFAULT_DETECTOR_CONFIG))
    det.setImage(img)
    lsd = cv2.createLineSegmentDetector(cv2.LSD_REFINE_ADV)
    for c in rg.COLORS:
        bw = rg.color_mask_cv(det.hsv, CFG, c)
        ec = cv2.bitwise_and(bw, det.edges)
        lines, width, prec, nfa = lsd.detect(ec)
        mine, extra = cm.lsd_detect(ec)
        if lines is None:
            assert len(mine) == 0
            continue
        assert np.array_equal(mine, lines[:, 0])
        assert np.allclose(extra[:, 0], width[:, 0]) and np.allclose(extra[:, 2], nfa[:, 0], rtol=1e-9, atol=1e-9)


def test_preprocess_resize_crop_transform_vs_cv2():
    img = synth.frame(9)
    for isz, cut in [((120, 160), 40), ((100, 130), 7), ((480, 640), 160), ((77, 201), 0)]:
        for sc, sf in [((1, 1, 1), (0, 0, 0)), ((1.1, 0.93, 1.27), (3.5, -7.25, 12.0)), ((0.5, 2.5, 1), (-40, 0.5, 100))]:
            assert np.array_equal(cm.preprocess(img, isz, cut, sc, sf), rg.preprocess(img, isz, cut, sc, sf)), (isz, cut, sc)


def test_ground_projection_and_sanity_vs_glue():
    rng = np.random.default_rng(1)
    lines = rng.uniform(-20, 200, (400, 4)).astype(np.float32)
    col = rng.integers(0, 3, 400).astype(np.uint8)
    pixn, ground, keep = cm.project_filter(lines, col, (120, 160), 40, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY)
    gp = rg.GroundProjection()
    ref_pix = ((lines + np.array((0, 40, 0, 40))) * np.array((1. / 160, 1. / 120, 1. / 160, 1. / 120))).astype(np.float32)
    assert np.array_equal(pixn, ref_pix)
    ref_g = gp.project_segments(ref_pix)
    ok = np.isfinite(ref_g).all(axis=1)
    assert np.abs(ground[ok] - ref_g[ok]).max() / max(1.0, np.abs(ref_g[ok]).max()) < 1e-9
    assert np.array_equal(keep.astype(bool), rg.sanity_keep(ground, col))


def test_lbd_image_prep_vs_cv2_and_descriptor_properties():
    img = synth.frame(11)
    gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    blur, dx, dy = cm.gauss5_sobel(gray)
    ref_blur = cv2.GaussianBlur(gray, (5, 5), 1)
    assert np.array_equal(blur, ref_blur)
    assert np.array_equal(dx, cv2.Sobel(ref_blur, cv2.CV_16S, 1, 0, ksize=3))
    assert np.array_equal(dy, cv2.Sobel(ref_blur, cv2.CV_16S, 0, 1, ksize=3))
    o = cm.front_end_frame(img, CFG, (480, 640), 0, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY, descriptors=True)
    d72, d32 = o["desc72"], o["desc32"]
    assert np.allclose(np.linalg.norm(d72, axis=1), 1.0, atol=1e-5)     # final L2 renormalisation (B.2)
    assert (d72 <= 0.4 / np.linalg.norm(np.minimum(d72, 0.4), axis=1, keepdims=True).min() + 1e-3).all()
    # the binary code is exactly the 32 band-pair comparisons of the float descriptor
    comb = [(a, b) for a in range(9) for b in range(a + 1, 9)]
    comb = [c for c in comb if c not in [(0, 7), (0, 8), (1, 7), (1, 8)]]
    assert len(comb) == 32
    for i in range(min(10, len(d72))):
        for k, (a, b) in enumerate(comb):
            bits = sum((1 << j) for j in range(8) if d72[i, 8 * a + j] > d72[i, 8 * b + j])
            assert d32[i, k] == bits
    # keyline: numOfPixels = max(|dx|,|dy|)+1 on rounded endpoints
    kl = o["keylines"]
    e = np.rint(kl[:, :4]).astype(int)
    assert np.array_equal(kl[:, 6].astype(int), np.maximum(abs(e[:, 2] - e[:, 0]), abs(e[:, 3] - e[:, 1])) + 1)


def test_knn_oracle_vs_bfmatcher_and_ties():
    q, m, src = synth.descriptor_sets(300, 5000, seed=2)
    for k in (1, 2, 4):
        oi, od = cm.knn_hamming(q, m, k)
        bi, bd = rg.knn_hamming_bf(q, m, k)
        assert np.array_equal(oi, bi) and np.array_equal(od, bd)
    oi, od = cm.knn_hamming(q, m, 1)
    assert (oi[:64, 0] == np.arange(64)).all()
    oi, od = cm.knn_hamming(q[:4], m[:2], 3)
    assert (oi[:, 2] == -1).all() and (od[:, 2] == -1).all()


def test_nfa_known_answers():
    # survey known-answer: NFA values of the 160x80 white image of synthetic seed 0 are reproduced by cv2 itself;
    # here: closed-form corner cases of nfa()
    import math
    lib = cm.lib()
    logNT = 5 * (math.log10(128) + math.log10(64)) / 2 + math.log10(11.0)
    assert lib.orc_nfa(0, 0, 0.125, logNT) == -logNT
    assert abs(lib.orc_nfa(10, 10, 0.125, logNT) - (-logNT - 10 * math.log10(0.125))) < 1e-12
    assert lib.orc_nfa(17, 15, 0.125, logNT) > 0 > lib.orc_nfa(16, 14, 0.125, logNT)


def test_golden_lane_filter_votes():
    """the restated vote loop equals the reference's own LaneFilterHistogram.generate_measurement_
    likelihood (golden generated by tests/golden/make_golden_lane_filter.py from /root/reference)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lane_filter_votes.npz"))
    for k in range(len(g["seeds"])):
        counts = rg.lane_filter_votes(g["ground_%d" % k], g["color_%d" % k])
        assert counts.shape == (23, 30)
        ml = g["likelihood_%d" % k]
        assert counts.sum() > 0
        assert np.array_equal(counts / counts.sum(), ml)
        # the votes are exactly the segments line_sanity keeps
        assert counts.sum() == rg.sanity_keep(g["ground_%d" % k], g["color_%d" % k]).sum()


# ---- real camera frames (tests/golden/real_images.npz: the reference's own JPEGs through the reference's own class) ----
import realset  # noqa: E402


def _check_against_golden(o, gd, det=None):
    assert o["counts"] == gd["counts"]
    assert np.array_equal(o["lines_px"], gd["lines"]) and np.array_equal(o["normal64"], gd["normals"])
    assert np.array_equal(o["centers"], gd["centers"])
    for ci, c in enumerate(realset.COLORS):
        packed = np.packbits(o["bw"][ci] > 0)
        if c + "_area" in gd:
            assert np.array_equal(packed, gd[c + "_area"])
        else:
            assert realset.crc(packed) == gd[c + "_area_crc"]
    packed = np.packbits(o["edges"] > 0)
    if "edges" in gd:
        assert np.array_equal(packed, gd["edges"])
    else:
        assert realset.crc(packed) == gd["edges_crc"]


def test_golden_real_images_cmodel_native_and_default_geometry():
    """28 real 640x480 frames of the reference repo: the C model reproduces the reference class's Detections exactly
    (segment order, endpoints, normals, centres, area masks, Canny map) at 640x480 and at 160x120 cut 40."""
    nseg = 0
    for i in range(realset.count()):
        img = realset.image(i)
        for pre, isz, cut in (("n%d" % i, (480, 640), 0), ("d%d" % i, (120, 160), 40)):
            o = cm.front_end_frame(img, CFG, isz, cut, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY)
            _check_against_golden(o, realset.golden(pre))
            nseg += sum(o["counts"])
    assert nseg > 4000


def test_golden_real_images_glue():
    det = rg.LineDetectorLSD(dict(rg.DEFAULT_DETECTOR_CONFIG))
    gp = rg.GroundProjection()
    for i in range(0, realset.count(), 3):
        r = rg.front_end_frame(realset.image(i), det, gp, (120, 160), 40)
        gd = realset.golden("d%d" % i)
        assert r["counts"] == gd["counts"] and np.array_equal(r["lines_px"], gd["lines"])


def test_golden_yaml_threshold_sets():
    """Every distinct threshold set of the shipped line_detector_node/*.yaml files (canny 50/150, 60/150, the other HSV
    ranges): C model == the reference class run with that configuration."""
    sets, frames = realset.yaml_sets()
    assert {"bad_lighting", "universal", "226-night"} <= set(sets)
    for name, conf in sets.items():
        cfg = rg.check_configuration(dict(conf))
        for i in frames:
            o = cm.front_end_frame(realset.image(i), cfg, (120, 160), 40, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY)
            _check_against_golden(o, realset.golden("y_%s_%d" % (name, i)))


# ---- LBD / KeyLine / matcher: pinned to the reference's own compiled code (oracle/_ref via tests/golden/make_golden_lbd.py) ----
LBD_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lbd_reference.npz")


def _lbd_case_image(case):
    kind, idx, H, W, dense = [int(v) for v in case]
    return synth.frame(idx, H, W, dense=bool(dense)) if kind == 0 else realset.image(idx)


def test_golden_lbd_keylines_and_descriptors_vs_compiled_reference():
    """KeyLine fill (LSDDetector_custom.cpp:130-215) and computeLBD (binary_descriptor_custom.cpp:1026-1372) of the C model
    equal the outputs of the reference's compiled C++: every KeyLine field, all 72 floats and all 256 bits, exactly."""
    g = np.load(LBD_GOLD)
    nl = 0
    for k, case in enumerate(g["cases"]):
        img = _lbd_case_image(case)
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        blur, dx, dy = cm.gauss5_sobel(gray)
        kl, d72, d32 = cm.lbd(g["%d_lines" % k], dx, dy)
        ref = g["%d_keylines" % k]     # startX, startY, endX, endY, lineLength, numOfPixels, angle, response, size, class_id
        assert np.array_equal(kl[:, :5], ref[:, :5]) and np.array_equal(kl[:, 6], ref[:, 5])
        assert np.array_equal(kl[:, 5], ref[:, 6]) and np.array_equal(kl[:, 7], ref[:, 7])
        assert np.array_equal(d32, g["%d_desc32" % k])
        assert np.array_equal(d72, g["%d_desc72" % k])
        nl += len(d32)
    assert nl >= 2000


def test_golden_knn_reference_order():
    """orc_knn_mihasher == BinaryDescriptorMatcher::knnMatch of the compiled reference (indices AND order inside equal
    distances) on a tie-heavy set; the smallest-index rule (cv2.BFMatcher) differs there, distances never do."""
    g = np.load(LBD_GOLD)
    q, m = g["knn_q"], g["knn_m"]
    differs = 0
    for k in (1, 2, 4, 8):
        i, d = cm.knn_mihasher(q, m, k)
        assert np.array_equal(i, g["knn_idx_%d" % k]) and np.array_equal(d, g["knn_dist_%d" % k])
        bi, bd = cm.knn_hamming(q, m, k, 128)
        assert np.array_equal(bd, d)
        differs += int((bi != i).sum())
    assert differs > 0      # the set really exercises the tie rule


def test_live_compiled_reference_when_present():
    """Where oracle/_ref exists (authoring container, or the prebuilt library that travels to the GPU box): fresh frames,
    C model vs the reference's compiled code."""
    from oracle import refnative as rn
    if not rn.available():
        pytest.skip("oracle/_ref not built here")
    for seed in (31, 32):
        o = cm.front_end_frame(synth.frame(seed), CFG, (480, 640), 0, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY, descriptors=True)
        kl, d72, d32 = rn.keylines_lbd(o["gray"], o["lines_px"])
        assert np.array_equal(d32, o["desc32"]) and np.array_equal(d72, o["desc72"])
    q, m, _ = synth.descriptor_sets(100, 3000, seed=9)
    for k in (1, 3):
        ri, rd = rn.knn_match(q, m, k)
        oi, od = cm.knn_mihasher(q, m, k)
        assert np.array_equal(ri, oi) and np.array_equal(rd, od)


def test_golden_lane_filter_belief_chain():
    """predict / update / getEstimate / getMax of the restated filter == the reference's LaneFilterHistogram driven over a
    24-frame sequence (golden from tests/golden/make_golden_lane_filter.py), bit for bit."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lane_filter_votes.npz"))
    f = rg.LaneFilterHistogram()
    assert np.array_equal(f.belief, g["filter_belief0"])
    frames = synth.sequence(24, base_seed=60)
    for t in range(24):
        o = cm.front_end_frame(frames[t], CFG, (480, 640), 0, rg.DEFAULT_CAMERA, rg.DEFAULT_HOMOGRAPHY)
        f.predict(*g["filter_dvw"][t])
        f.update(o["ground"], o["color"])
        assert f.getEstimate() + [f.getMax()] == g["filter_estimates"][t].tolist()
        if t in (0, 11, 23):
            assert np.array_equal(f.belief, g["filter_belief_%d" % t])


def test_jpeg_oracle_equals_cv2_imdecode():
    """The restated baseline JPEG decode (islow IDCT, fancy upsampling, fixed-point YCbCr) == cv2.imdecode (libjpeg-turbo) bit
    for bit: the reference's real frames and re-encoded synthetic frames (subsamplings, qualities, restart intervals, odd sizes)."""
    for i in range(realset.count()):
        assert np.array_equal(cm.jpeg_decode(realset.jpeg(i)), realset.image(i)), i
    n = 0
    for (H, W) in [(480, 640), (123, 161), (97, 203)]:
        im = synth.frame(5, H, W)
        for q in (40, 90):
            for ss in (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
                       cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440):
                for rst in (0, 5):
                    ok, enc = cv2.imencode('.jpg', im, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss,
                                                        cv2.IMWRITE_JPEG_RST_INTERVAL, rst])
                    assert np.array_equal(cm.jpeg_decode(enc), cv2.imdecode(enc, cv2.IMREAD_COLOR)), (H, W, q, ss, rst)
                    n += 1
    ok, enc = cv2.imencode('.jpg', cv2.cvtColor(synth.frame(1), cv2.COLOR_BGR2GRAY))
    assert np.array_equal(cm.jpeg_decode(enc), cv2.imdecode(enc, cv2.IMREAD_COLOR))
    ok, enc = cv2.imencode('.jpg', synth.frame(1), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    with pytest.raises(ValueError):
        cm.jpeg_decode(enc)
    assert n == 48
