"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/lsf.h declares; host-side
argument checking mirrors the reference's errors; nothing computes without a GPU (loud failure, no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import lane_slam_b200 as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lsf.h")).read()
    declared = set(re.findall(r"LSF_API\s+[\w\s\*]+?\b(lsf_\w+)\s*\(", hdr))
    assert declared, "no LSF_API declarations parsed"
    assert declared == set(L.exported_symbols())
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name


def test_config_struct_matches_header_layout():
    lib = ctypes.CDLL(L.LIB_PATH)
    from lane_slam_b200._lib import LsfConfig, LsfSegments
    cfg = LsfConfig()
    lib.lsf_default_config.argtypes = [ctypes.POINTER(LsfConfig)]
    assert lib.lsf_default_config(ctypes.byref(cfg)) == 0
    # defaults = the reference's YAML files
    assert (cfg.img_h, cfg.img_w, cfg.top_cutoff) == (120, 160, 40)
    assert list(cfg.hsv_lo[1]) == [25, 140, 100] and list(cfg.hsv_hi[3]) == [180, 255, 255]
    assert (cfg.canny_lo, cfg.canny_hi, cfg.dilation_kernel_size) == (80, 200, 3)
    assert abs(cfg.K[0] - 307.7379294605756) < 1e-12 and abs(cfg.Hgnd[7] + 0.007185673) < 1e-12
    assert (cfg.cam_w, cfg.cam_h) == (640, 480) and abs(cfg.lanewidth - 0.23) < 1e-15 and cfg.phi_max == 1.5
    assert ctypes.sizeof(LsfSegments) == 16 + 13 * 8


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(L.LsfError) as e:
        L.FrontEnd()
    assert e.value.code == -4 and "no CPU fallback" in str(e.value)
    det = L.LineDetectorB200(dict(L.DEFAULT_DETECTOR_CONFIGURATION))     # construction only checks the config
    with pytest.raises(L.LsfError):
        det.setImage(np.zeros((80, 160, 3), np.uint8))


def test_configuration_errors_match_reference():
    with pytest.raises(ValueError):
        L.LineDetectorB200("not a dict")
    with pytest.raises(ValueError):
        L.LineDetectorB200(dict(L.DEFAULT_DETECTOR_CONFIGURATION, extra_key=1))
    cfg = dict(L.DEFAULT_DETECTOR_CONFIGURATION); cfg.pop("canny_thresholds")
    with pytest.raises(ValueError):
        L.LineDetectorB200(cfg)
    det = L.LineDetectorB200(dict(L.DEFAULT_DETECTOR_CONFIGURATION))
    assert isinstance(det.hsv_white1, np.ndarray) and det.dilation_kernel_size == 3   # parameters.py:29-32
    with pytest.raises(Exception):
        det.detectLines("white")        # before setImage


def test_product_package_does_not_import_oracle():
    import subprocess, sys
    code = "import sys; import lane_slam_b200; assert not any(m.startswith('oracle') for m in sys.modules), 'oracle imported'"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for fn in os.listdir(os.path.join(ROOT, "lane_slam_b200")):
        if fn.endswith(".py"):
            src = open(os.path.join(ROOT, "lane_slam_b200", fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|from\s+\.\.?\s*import\s+oracle|import_module\(.oracle", src, re.M), fn


def test_messages_from_batch():
    from lane_slam_b200.frontend import SegmentBatch
    from lane_slam_b200 import segment_lists_from_batch
    S = 3
    arrays = dict(counts=np.array([[1, 1, 0], [1, 0, 0]], np.int32), frame_offset=np.array([0, 2, 3], np.int32),
                  color=np.array([0, 1, 0], np.uint8), lines_px=np.zeros((S, 4), np.float32), normals=np.zeros((S, 2)),
                  centers=np.zeros((S, 2), np.float32), pixels_normalized=np.arange(12, dtype=np.float32).reshape(S, 4),
                  normal_f32=np.ones((S, 2), np.float32), ground=np.arange(12, dtype=np.float64).reshape(S, 4),
                  keep=np.array([1, 0, 1], np.uint8), desc=np.zeros((S, 32), np.uint8), match_idx=None, match_dist=None)
    b = SegmentBatch(2, S, arrays, 0)
    det = segment_lists_from_batch(b, "detector")
    assert [len(x.segments) for x in det] == [2, 1] and det[0].segments[1].color == 1
    assert det[0].segments[0].pixels_normalized[1].x == np.float32(2.0)
    san = segment_lists_from_batch(b, "sanity")
    assert [len(x.segments) for x in san] == [1, 1] and san[1].segments[0].points[0].x == 8.0
