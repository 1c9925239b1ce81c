"""Build recipe for oracle/_ref/libref_line_descriptor.so (TEST INFRASTRUCTURE, git-ignored output).

The reference's only native code on the path -- src/line_descriptor (vendored opencv_contrib line_descriptor: KeyLine
fill, computeLBD, BinaryDescriptorMatcher/Mihasher) -- is compiled here UNMODIFIED, from the files where they lie under
/root/reference, with plain g++ (no cmake / catkin, which the reference's own build would need, and no OpenCV C++
development files, which this image lacks): the OpenCV headers those sources include are satisfied by the small
self-written stand-in oracle/ref_shim/cvshim.hpp (see its header for what is functional), and the three image
primitives on the descriptor path (BGR2GRAY, GaussianBlur 5x5, Sobel 3x3) come from the C oracle, which is pinned
bit-exactly to cv2 4.13.  The library only exists where /root/reference exists (this container); the GPU box uses
the prebuilt file that travels with the snapshot, and the golden vectors generated from it (tests/golden/lbd_reference.npz).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/line_descriptor"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libref_line_descriptor.so")
SRCS = ["LSDDetector_custom.cpp", "binary_descriptor_custom.cpp", "binary_descriptor_matcher.cpp"]


def available():
    return all(os.path.exists(os.path.join(REF, "src", s)) for s in SRCS)


def build(force=False):
    """-> path of the library, or None when neither the reference sources nor a prebuilt library are present."""
    if not available():
        return OUT if os.path.exists(OUT) else None
    deps = [os.path.join(REF, "src", s) for s in SRCS] + [os.path.join(HERE, "ref_shim", f) for f in ("cvshim.hpp", "ref_harness.cpp")] + \
        [os.path.join(HERE, "csrc", "lane_oracle.c")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    obj = os.path.join(OUT_DIR, "lane_oracle.o")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-fvisibility=hidden", "-c",
                           os.path.join(HERE, "csrc", "lane_oracle.c"), "-o", obj])
    cmd = ["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fno-fast-math", "-w", "-shared", "-fPIC",
           "-I", os.path.join(HERE, "ref_shim"), "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "src"),
           os.path.join(HERE, "ref_shim", "ref_harness.cpp")] + [os.path.join(REF, "src", s) for s in SRCS] + [obj, "-o", OUT, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
