"""CPU test: the float sinf / cosf / atan2f restatements the LBD kernel uses (lane_slam_b200/csrc/libm_f32.cuh)
are bit-identical to the C library the reference's C++ would call.  The header is compiled for the host; the default
run samples every 97th float (a few seconds); `LMF_EXHAUSTIVE=1` checks every float (minutes; 0 mismatches recorded
in DESIGN.md)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_libm_f32_restatement_matches_glibc(tmp_path):
    exe = str(tmp_path / "libm_f32_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-o", exe,
                           os.path.join(ROOT, "oracle", "csrc", "libm_f32_check.c"), "-lm"])
    args = ["1", "1500000000"] if os.environ.get("LMF_EXHAUSTIVE") else ["97", "20000000"]
    out = subprocess.run([exe] + args, capture_output=True, text=True)
    assert out.returncode == 0 and "sincos_bad 0 atan_bad 0" in out.stdout, out.stdout + out.stderr
