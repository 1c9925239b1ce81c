"""Build recipe for the C oracle (gcc only; outputs into oracle/_build/, git-ignored)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "lane_oracle.c")
SRCS = [SRC, os.path.join(HERE, "csrc", "jpeg_oracle.c")]
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liblane_oracle.so")


def build(force=False):
    """Compile csrc/lane_oracle.c -> _build/liblane_oracle.so (no FMA contraction, strict IEEE)."""
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in SRCS):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-fvisibility=hidden",
           ] + SRCS + ["-o", OUT, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
