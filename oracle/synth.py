"""Synthetic Duckietown-style lane frames.  Test/benchmark input generator.

The reference ships no generator; this file defines the inputs used by the parity tests and
by bench.py.  Colours are chosen to sit inside the default HSV ranges of
src/duckietown/config/baseline/line_detector/line_detector_node/default.yaml:16-23.
Deterministic for a given (seed, H, W, dense) with the cv2 / numpy versions of this image.
"""
import math

import cv2
import numpy as np

WHITE = (255, 255, 255)
YELLOW = (0, 230, 240)   # BGR -> H_cv=29, S=255
RED = (30, 30, 220)      # BGR -> H_cv=0
ROAD = (60, 60, 60)
SKY = (200, 160, 120)


def base_scene(H, W, dense=False, seed=0):
    """Un-warped scene: road, background, two white lines, dashed yellow centre, red stop bar."""
    img = np.empty((H, W, 3), np.uint8)
    img[:] = ROAD
    img[: H // 3] = SKY
    tw = max(2, W // 40)
    ty = max(2, W // 64)
    cv2.line(img, (int(W * 0.15), H - 1), (int(W * 0.42), H // 3), WHITE, tw)
    cv2.line(img, (int(W * 0.95), H - 1), (int(W * 0.60), H // 3), WHITE, tw)
    for t in np.linspace(0, 1, 6)[:-1]:
        p0 = (int(W * (0.55 - 0.05 * t)), int(H - 1 - (H * 2 / 3) * t))
        p1 = (int(W * (0.55 - 0.05 * (t + 0.1))), int(H - 1 - (H * 2 / 3) * (t + 0.1)))
        cv2.line(img, p0, p1, YELLOW, ty)
    cv2.line(img, (int(W * 0.3), int(H * 0.45)), (int(W * 0.7), int(H * 0.45)), RED, ty)
    if dense:
        rng = np.random.default_rng(1000003 + seed)
        n = 700
        xs = rng.integers(0, W, n)
        ys = rng.integers(H // 3, H, n)
        ang = rng.uniform(0, math.pi, n)
        ln = rng.integers(max(12, W // 50), max(24, W // 20), n)
        for i in range(n):
            q = (int(xs[i] + ln[i] * math.cos(ang[i])), int(ys[i] + ln[i] * math.sin(ang[i])))
            cv2.line(img, (int(xs[i]), int(ys[i])), q, WHITE if i % 2 == 0 else YELLOW, 2)
    return img


def _warp(img, rot_deg, scale, tx=0.0, ty=0.0):
    H, W = img.shape[:2]
    M = cv2.getRotationMatrix2D((W / 2.0, H / 2.0), rot_deg, scale)
    M[0, 2] += tx
    M[1, 2] += ty
    return cv2.warpAffine(img, M, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)


def frame(seed, H=480, W=640, dense=False):
    """One uint8 BGR frame [H,W,3]: base scene, per-seed similarity warp, additive noise 0..5."""
    rng = np.random.default_rng(seed)
    img = base_scene(H, W, dense=dense, seed=seed)
    rot = float(rng.uniform(-15, 15))
    sc = float(rng.uniform(0.9, 1.2))
    img = _warp(img, rot, sc)
    noise = rng.integers(0, 6, img.shape, dtype=np.uint8)
    return cv2.add(img, noise)


def sequence(n, base_seed=0, H=480, W=640, dense=False, start=0):
    """Frames start .. start+n-1 ([n,H,W,3]) of the sequence `base_seed`: one scene following a smooth pose trajectory
    (frame t: noise seed = base_seed + t)."""
    scene = base_scene(H, W, dense=dense, seed=base_seed)
    out = np.empty((n, H, W, 3), np.uint8)
    for i in range(n):
        t = start + i
        rot = 12.0 * math.sin(2 * math.pi * t / 240.0)
        sc = 1.05 + 0.12 * math.sin(2 * math.pi * t / 173.0 + 0.7)
        tx = 0.04 * W * math.sin(2 * math.pi * t / 97.0)
        ty = 0.03 * H * math.cos(2 * math.pi * t / 131.0)
        img = _warp(scene, rot, sc, tx, ty)
        rng = np.random.default_rng(base_seed + t)
        out[i] = cv2.add(img, rng.integers(0, 6, img.shape, dtype=np.uint8))
    return out


def descriptor_sets(nq=2000, nm=100000, seed=0, flip_frac=0.12, n_dup=64):
    """C4 workload: map = nm random 256-bit codes; queries = map rows with ~12 % bit flips, plus a
    block of duplicated map rows (exercises the smallest-index tie rule).  Returns (q, m, src_rows)."""
    rng = np.random.default_rng(seed)
    m = rng.integers(0, 256, (nm, 32), dtype=np.uint8)
    if n_dup and nm >= 4 * n_dup:
        m[nm - n_dup:] = m[:n_dup]          # rows [0,n_dup) appear twice: lower index must win
    src = rng.integers(0, nm, nq)
    src[: min(n_dup, nq)] = np.arange(min(n_dup, nq))
    flips = rng.random((nq, 256)) < flip_frac
    q = m[src] ^ np.packbits(flips, axis=1, bitorder="little")
    return q, m, src
