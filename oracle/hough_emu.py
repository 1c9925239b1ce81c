"""TEST INFRASTRUCTURE: the REAL source text of the kernels of lane_slam_b200/csrc/k_hough.cu (k_hough_p, k_hough_segments), cut out of
the file at build time and run as thread blocks made of host threads (oracle/csrc/cuda_threads_emu.h), with hough_core.cuh in its
device form (32 lanes, shuffles).  Checks what the host builds of the core cannot: the kernels' own indexing of the bit-planes, the
task loop, counts / offsets and the output rows."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lane_slam_b200", "csrc")

HARNESS = r'''
#define __CUDA_ARCH__ 1000            /* hough_core.cuh: the device branch (lanes, shuffles) on top of the host-thread shim */
#include "cuda_threads_emu.h"
#include <vector>
typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64;
namespace lsf { enum { PB_EDGE = @@PB_EDGE@@, PB_BW0 = @@PB_BW0@@, PB_EC0 = @@PB_EC0@@, PB_COUNT = @@PB_COUNT@@ }; }
#include "hough_core.cuh"
@@KERNELS@@

using namespace lsf;
struct PArgs { int h, w, wp, n_tasks; const u32 *planesB; int th, ml, mg; const float *trig; int32_t *accum; u8 *mask; u32 *nz; int32_t *raw; int max_lines; int *count; };
static void body_p(void *p)
{
    PArgs *a = (PArgs *)p;
    k_hough_p(a->h, a->w, a->wp, a->n_tasks, a->planesB, a->th, a->ml, a->mg, a->trig, a->accum, a->mask, a->nz, a->raw, a->max_lines, a->count);
}
struct SArgs { int h, w, wp, n_tasks; const u32 *planesB; const int32_t *raw; int max_lines; const int *count, *offset; int cut; double iw, ih;
               u8 *color; float *lines; double *normals; float *centers; float *pixn; float *nf32; };
static void body_s(void *p)
{
    SArgs *a = (SArgs *)p;
    k_hough_segments(a->h, a->w, a->wp, a->n_tasks, a->planesB, a->raw, a->max_lines, a->count, a->offset, a->cut, a->iw, a->ih, a->color, a->lines,
                     a->normals, a->centers, a->pixn, a->nf32);
}

// planesB: [n][PB_COUNT][h][wp] words as the front end leaves them.  Outputs sized by the caller (cap rows).  Returns the number
// of rows, or -1 if a colour image overflowed max_lines.
extern "C" __attribute__((visibility("default")))
int hke_run(const u32 *planesB, int n, int h, int w, int th, int ml, int mg, int top_cutoff, int img_h, int img_w, int max_lines, int cap,
            int *counts, u8 *color, float *lines, double *normals, float *centers, float *pixn, float *nf32)
{
    const int wp = (w + 31) / 32, tasks = n * 3, nwarps = HOUGH_WARPS_PER_BLOCK;
    const size_t acc_sz = (size_t)hp::NUMANGLE * hp::numrho(w, h), npix = (size_t)h * w;
    std::vector<float> trig(hp::NUMANGLE * 2);
    hp::make_trig(trig.data());
    std::vector<int32_t> accum(nwarps * acc_sz, 0x33333333), raw((size_t)tasks * max_lines * 4, -7);
    std::vector<u8> mask(nwarps * npix, 0xEE);
    std::vector<u32> nz(nwarps * npix, 0xABABABABu);
    std::vector<int> count(tasks + 2, 0), offset(tasks + 1, 0);
    PArgs pa = {h, w, wp, tasks, planesB, th, ml, mg, trig.data(), accum.data(), mask.data(), nz.data(), raw.data(), max_lines, count.data()};
    gridDim = cuemu::Idx{1, 1, 1};
    cuemu_run_block(HOUGH_WARPS_PER_BLOCK * 32, 0, body_p, &pa);
    if (count[tasks + 1]) return -1;
    for (int t = 0; t < tasks; ++t) { counts[t] = count[t]; offset[t + 1] = offset[t] + count[t]; }
    const int S = offset[tasks];
    if (S > cap) return -2;
    SArgs sa = {h, w, wp, tasks, planesB, raw.data(), max_lines, count.data(), offset.data(), top_cutoff, 1.0 / (double)img_w, 1.0 / (double)img_h,
                color, lines, normals, centers, pixn, nf32};
    gridDim = cuemu::Idx{(unsigned)tasks, 4, 1};
    for (int t = 0; t < tasks; ++t)
        for (int by = 0; by < 4; ++by) cuemu_run_block(256, t, body_s, &sa, by);
    return S;
}
'''


def _enum(name):
    import re
    text = open(os.path.join(CSRC, "common.cuh")).read()
    m = re.search(r"\b%s\s*=\s*(\d+)" % name, text)
    assert m, name
    return m.group(1)


def kernels_text():
    text = open(os.path.join(CSRC, "k_hough.cu")).read()
    body = text[text.index("namespace lsf {"):text.index("using namespace lsf;")]
    assert "k_hough_p" in body and "k_hough_segments" in body and "asm" not in body
    return body


def build(out_dir, sanitize=False):
    src = os.path.join(out_dir, "hough_kernels_emu.cpp")
    h = HARNESS.replace("@@KERNELS@@", kernels_text())
    for name in ("PB_EDGE", "PB_BW0", "PB_EC0", "PB_COUNT"):
        h = h.replace("@@%s@@" % name, _enum(name))
    with open(src, "w") as f:
        f.write(h)
    so = os.path.join(out_dir, "libhke%s.so" % ("_tsan" if sanitize else ""))
    cmd = ["g++", "-std=c++14", "-O1" if sanitize else "-O2", "-g", "-shared", "-fPIC", "-pthread", "-I", os.path.join(ROOT, "oracle", "csrc"), "-I", CSRC,
           src, os.path.join(ROOT, "oracle", "csrc", "cuda_threads_emu.cpp"), "-o", so]
    if sanitize:
        cmd.insert(1, "-fsanitize=thread")
    subprocess.check_call(cmd)
    return so
