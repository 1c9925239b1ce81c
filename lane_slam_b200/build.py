"""Build liblsf.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m lane_slam_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblsf.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # no FMA contraction: the fixed-point / float models are op-for-op
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--threads", "0",
    "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "lsf.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = ["-DLSF_GROW_PROF"] if os.environ.get("LSF_GROW_PROF") else []   # developer aid: cycle counters in k_lsd_grow
    if os.environ.get("LSF_JPEG_PROF"):                                      # developer aid: phase cycle counts in k_jpeg_huff
        extra.append("-DLSF_JPEG_PROF")
    if os.environ.get("LSF_JPEG_LUT_BITS"):
        extra.append("-DLSF_JPEG_LUT_BITS=" + os.environ["LSF_JPEG_LUT_BITS"])
    if os.environ.get("LSF_JT"):
        extra.append("-DLSF_JT=" + os.environ["LSF_JT"])
    if os.environ.get("LSF_GROW_PER_SM_BUILD"):                              # developer aid: occupancy experiments
        extra.append("-DLSF_GROW_PER_SM_BUILD=" + os.environ["LSF_GROW_PER_SM_BUILD"])
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + sources() + ["-o", OUT]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
