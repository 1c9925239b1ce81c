"""CPU test of the GPU JPEG decoder's algorithm: lane_slam_b200/csrc/jpeg_core.cuh (the arithmetic the CUDA kernels run) is
compiled for the host and driven by oracle/csrc/jpeg_parallel_check.cpp exactly the way k_jpeg.cu orchestrates it --
stuffing removal, self-synchronising parallel Huffman decode simulated thread by thread, prefix sums, DC prediction, islow IDCT,
fancy upsampling, colour conversion -- and must reproduce cv2.imdecode bit for bit."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest

import realset
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("jpc") / "libjpc.so")
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "oracle", "csrc", "jpeg_parallel_check.cpp")])
    lib = C.CDLL(so)

    def decode(data, shape, sub_bytes, gpu_loop=False):
        b = np.ascontiguousarray(data, np.uint8)
        out = np.zeros(shape, np.uint8)
        r, n = C.c_int(), C.c_int()
        fn = lib.jpc_decode_gpu_loop if gpu_loop else lib.jpc_decode
        rc = fn(b.ctypes.data_as(C.c_void_p), C.c_size_t(len(b)), out.ctypes.data_as(C.c_void_p), int(sub_bytes), C.byref(r), C.byref(n))
        return rc, out, r.value, n.value
    return decode


def test_parallel_decode_equals_cv2_on_real_frames(sim):
    for i in range(0, realset.count(), 2):
        img = realset.image(i)
        for sub in (48, 152):
            rc, out, rounds, nsub = sim(realset.jpeg(i), img.shape, sub)
            assert rc == 0 and np.array_equal(out, img), (i, sub)
            assert rounds < nsub          # the self-synchronisation converges long before the sequential worst case


def test_parallel_decode_samplings_sizes_qualities(sim):
    for (H, W) in [(480, 640), (123, 161), (97, 203)]:
        im = synth.frame(5, H, W)
        for q in (40, 95):
            for ss in (cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
                       cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440):
                ok, enc = cv2.imencode('.jpg', im, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss])
                ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
                rc, out, rounds, nsub = sim(enc, ref.shape, 64)
                assert rc == 0 and np.array_equal(out, ref), (H, W, q, ss)
    ok, enc = cv2.imencode('.jpg', cv2.cvtColor(synth.frame(1), cv2.COLOR_BGR2GRAY))
    ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    rc, out, _, _ = sim(enc, ref.shape, 64)
    assert rc == 0 and np.array_equal(out, ref)
    ok, enc = cv2.imencode('.jpg', synth.frame(1), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert sim(enc, (480, 640, 3), 64)[0] == -2                 # outside the scope: reported, not mis-decoded
    ok, enc = cv2.imencode('.jpg', synth.frame(1), [cv2.IMWRITE_JPEG_RST_INTERVAL, 4])
    assert sim(enc, (480, 640, 3), 64)[0] == -2


def test_the_kernels_own_symbol_loop_equals_cv2(sim):
    """k_jpeg.cu decodes with a loop of its own (64-bit bit buffer, two-level table word per symbol, packed table selector); its
    host restatement (gpu_scan_span in jpeg_parallel_check.cpp), run with the kernel's geometry -- 512 subsequences per image,
    20 zero bytes and then garbage behind the stream -- must converge and reproduce cv2.imdecode: real frames, every sampling,
    optimised Huffman tables (long codes, second-level tables), very low and very high quality, noise, grayscale."""
    for i in range(realset.count()):
        img = realset.image(i)
        rc, out, rounds, nsub = sim(realset.jpeg(i), img.shape, 0, gpu_loop=True)
        assert rc == 0 and np.array_equal(out, img), i
        assert nsub <= 512 and rounds < 64
    rng = np.random.default_rng(0)
    images = [synth.frame(2), synth.frame(7, 123, 161), rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)]
    P = cv2
    for q in (5, 60, 100):
        for extra in ([], [P.IMWRITE_JPEG_OPTIMIZE, 1], [P.IMWRITE_JPEG_SAMPLING_FACTOR, P.IMWRITE_JPEG_SAMPLING_FACTOR_444],
                      [P.IMWRITE_JPEG_SAMPLING_FACTOR, P.IMWRITE_JPEG_SAMPLING_FACTOR_422, P.IMWRITE_JPEG_OPTIMIZE, 1],
                      [P.IMWRITE_JPEG_SAMPLING_FACTOR, P.IMWRITE_JPEG_SAMPLING_FACTOR_440]):
            for im in images:
                ok, enc = cv2.imencode('.jpg', im, [P.IMWRITE_JPEG_QUALITY, q] + extra)
                ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
                for sub in (0, 16):
                    rc, out, _, _ = sim(enc, ref.shape, sub, gpu_loop=True)
                    assert rc == 0 and np.array_equal(out, ref), (q, extra, im.shape, sub)
    ok, enc = cv2.imencode('.jpg', cv2.cvtColor(synth.frame(1), cv2.COLOR_BGR2GRAY), [P.IMWRITE_JPEG_QUALITY, 100, P.IMWRITE_JPEG_OPTIMIZE, 1])
    ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)
    rc, out, _, _ = sim(enc, ref.shape, 0, gpu_loop=True)
    assert rc == 0 and np.array_equal(out, ref)


def _tsan_runtime():
    p = subprocess.run(["g++", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    return p if os.path.isabs(p) and os.path.exists(p) else None


def test_huffman_kernel_source_runs_on_host_threads(tmp_path):
    """The REAL source text of the four kernels of k_jpeg.cu, cut out of the file and run as thread blocks of host threads
    (oracle/jpeg_huff_emu.py, oracle/csrc/cuda_threads_emu.h) with the grids the library launches: un-stuffing, Huffman rounds with
    compaction, scans, coefficient writes, DC prediction, the IDCT that reads flagged rows only and must leave the coefficient
    buffer zero, both colour kernels -- the frames equal cv2.imdecode, with 64 and with 512 Huffman threads per image."""
    import sys
    from oracle import jpeg_huff_emu as J
    for jt, big in ((64, False), (512, False)):
        d = tmp_path / ("jt%d" % jt)
        d.mkdir()
        so = J.build(str(d), jt=jt)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "jpeg_huff_emu_run.py"), so] + (["big"] if big else ["one"] if jt == 512 else []),
                             capture_output=True, text=True, timeout=900)
        assert out.returncode == 0 and "emulated k_jpeg kernels ok" in out.stdout, out.stderr[-2000:]


def test_huffman_kernel_has_no_shared_memory_race(tmp_path):
    """The same build under ThreadSanitizer: the barriers are the only synchronisation between the host threads, so every
    shared-memory access pair a missing __syncthreads leaves unordered is reported (the CPU stand-in for compute-sanitizer
    racecheck).  Control: with the ordering this kernel shipped with for one day -- the 'changed' flag cleared BEFORE the barrier at
    the top of a re-decode round, while the next warp may still be reading it -- the detector must fire."""
    import sys
    from oracle import jpeg_huff_emu as J
    rt = _tsan_runtime()
    if rt is None:
        pytest.skip("no ThreadSanitizer runtime")
    env = dict(os.environ, LD_PRELOAD=rt, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=0 history_size=7")

    def run(so, *more):
        return subprocess.run([sys.executable, os.path.join(ROOT, "tests", "jpeg_huff_emu_run.py"), so] + list(more), capture_output=True, text=True,
                              timeout=1200, env=env)
    good = tmp_path / "good"; good.mkdir()
    out = run(J.build(str(good), jt=512, sanitize=True), "huff")         # the shipped block size; a real 640x480 frame through the Huffman pass
    assert out.returncode == 0 and "emulated k_jpeg kernels ok" in out.stdout, out.stderr[-3000:]
    assert "ThreadSanitizer" not in out.stderr, out.stderr[-3000:]
    # control
    fixed = "        __syncthreads();\n        s_chg[t] = 0;"
    text = J.kernel_text()
    assert fixed in text
    orig = J.kernel_text
    J.kernel_text = lambda: text.replace(fixed, "        s_chg[t] = 0;\n        __syncthreads();", 1)
    try:
        bad = tmp_path / "bad"; bad.mkdir()
        so = J.build(str(bad), jt=64, sanitize=True)
    finally:
        J.kernel_text = orig
    seen = False
    for _ in range(3):          # the detector keeps a bounded history per memory cell: give it three runs to catch the pair
        out = run(so)
        if "WARNING: ThreadSanitizer: data race" in out.stderr and "k_jpeg_huff" in out.stderr:
            seen = True
            break
    assert seen, out.stderr[-2000:]


def test_malformed_files_never_reach_out_of_bounds(tmp_path):
    """600 mutated files (flipped header bytes, truncations, flipped entropy bytes, garbage) through jd::parse -- the host-side step
    of lsf_front_end_batch_jpeg -- and the decoder's logic under AddressSanitizer: an error code or some image, no stray access."""
    import sys
    rt = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not (os.path.isabs(rt) and os.path.exists(rt)):
        pytest.skip("no AddressSanitizer runtime")
    so = str(tmp_path / "libjpc_asan.so")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-g", "-fsanitize=address,bounds,null,alignment", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "oracle", "csrc", "jpeg_parallel_check.cpp")])
    env = dict(os.environ, LD_PRELOAD=rt, ASAN_OPTIONS="detect_leaks=0")
    for seed in (11, 12):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "jpeg_fuzz_run.py"), so, str(seed), "200"], capture_output=True, text=True,
                             timeout=1200, env=env)
        assert out.returncode == 0 and "fuzz ok" in out.stdout and "ERROR: AddressSanitizer" not in out.stderr, out.stderr[-3000:]
    # the kernels themselves (their real source on host threads) on corrupted entropy data
    from oracle import jpeg_huff_emu as J
    d = tmp_path / "emu"; d.mkdir()
    so = J.build(str(d), jt=64, sanitize="address")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "jpeg_kernel_fuzz_run.py"), so, "5", "40"], capture_output=True, text=True,
                         timeout=1200, env=env)
    assert out.returncode == 0 and "kernel fuzz ok" in out.stdout and "ERROR: AddressSanitizer" not in out.stderr, out.stderr[-3000:]
