"""TEST INFRASTRUCTURE: builds a host library that runs the REAL source text of the kernels of lane_slam_b200/csrc/k_jpeg.cu --
k_jpeg_huff (byte un-stuffing, the self-synchronising Huffman rounds with their compaction, prefix sums, coefficient writes, DC
prediction), k_jpeg_idct (flagged rows only, zeroed behind the read) and k_jpeg_color_420 / k_jpeg_color -- as thread blocks made of
host threads (oracle/csrc/cuda_threads_emu.h), with the grids lsf_front_end_batch_jpeg launches.
The kernel text is cut out of the .cu file at build time, so the test always runs what ships; only the two inline-PTX shared-memory
loads are replaced by plain loads.  With sanitize=True the library is built with ThreadSanitizer: the barriers are the only
synchronisation between the threads, so it reports the races a missing __syncthreads leaves."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "lane_slam_b200", "csrc")

HARNESS = r'''
#include "cuda_threads_emu.h"
#include <vector>
typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64;
#define JD_ATOMIC_ROWMASK
#include "jpeg_core.cuh"
static uintptr_t g_smem_anchor;
static inline u32 lds32(u32 a) { return *(const u32 *)((((uintptr_t)&g_smem_anchor) & ~(uintptr_t)0xffffffffu) | (uintptr_t)a); }
@@KERNEL@@
}  // namespace lsf

struct Launch {
    jd::Image g; const u8 *blob; const lsf::JpegItem *items; const jd::Tabs *tabs; u32 *clean; size_t clean_words;
    int16_t *coef; size_t coef_per_img; int16_t *dcdiff; u32 *rowmask; int *status;
};
static void body(void *p)
{
    Launch *l = (Launch *)p;
    lsf::k_jpeg_huff(l->g, l->blob, l->items, l->tabs, l->clean, l->clean_words, l->coef, l->coef_per_img, l->dcdiff, l->rowmask, l->status);
}

struct IdctArgs { jd::Image g; int n; int16_t *coef; size_t coef_per_img; const u32 *rowmask; const u16 *qtabs; u8 *plane; size_t plane_per_img; };
static void body_idct(void *p)
{
    IdctArgs *a = (IdctArgs *)p;
    lsf::k_jpeg_idct(a->g, a->n, a->coef, a->coef_per_img, a->rowmask, a->qtabs, a->plane, a->plane_per_img);
}
struct ColorArgs { jd::Image g; int n; const u8 *plane; size_t plane_per_img; u8 *bgr; size_t frame_bytes; };
static void body_color420(void *p)
{
    ColorArgs *a = (ColorArgs *)p;
    lsf::k_jpeg_color_420(a->g, a->n, a->plane, a->plane_per_img, a->bgr, a->frame_bytes);
}
static void body_color(void *p)
{
    ColorArgs *a = (ColorArgs *)p;
    lsf::k_jpeg_color(a->g, a->n, a->plane, a->plane_per_img, a->bgr, a->frame_bytes);
}

// data: one JPEG file.  bgr: [H][W][3] out.  Returns 0, < 0 for a file outside the decoder's scope, 100 + status if the kernel
// flagged the image.  prefill: low byte = what the stream buffer is filled with first (the kernel's buffer holds an older image);
// bit 8: stop after the Huffman pass (bgr is not written).
extern "C" __attribute__((visibility("default")))
int jhe_decode(const uint8_t *data, size_t len, uint8_t *bgr, int prefill)
{
    jd::Image im; jd::Tabs tabs;
    int rc = jd::parse(data, len, im, &tabs);
    if (rc) return rc;
    if (im.restart) return -2;
    lsf::JpegItem item; item.ent_off = im.ent_off; item.ent_len = im.ent_len; item.tabs = 0; item.pad = 0;
    const size_t nblocks = (size_t)im.mcux * im.mcuy * im.bpm;
    const size_t coef_per_img = ((nblocks + 3) & ~(size_t)3) * 64;
    const size_t clean_words = (im.ent_len + 3) / 4 + 8;
    std::vector<u32> clean(clean_words, 0x01010101u * (u32)(prefill & 0xff));
    std::vector<int16_t> coef(coef_per_img, 0), dcdiff(coef_per_img / 64, 0x7777);
    std::vector<u32> rowmask(coef_per_img / 256, 0);
    int status = 0;
    Launch l = {im, data, &item, &tabs, clean.data(), clean_words, coef.data(), coef_per_img, dcdiff.data(), rowmask.data(), &status};
    cuemu_run_block(lsf::JT, 0, body, &l);
    if (status) return 100 + status;
    if (prefill & 0x100) return 0;       // Huffman pass only (race detection on a large frame without the other kernels' thread spawning)
    // k_jpeg_idct: grid (ceil(blocks / 32), n) x 256 threads
    size_t plane_bytes = 0;
    for (int c = 0; c < im.ncomp; ++c) plane_bytes += (size_t)im.bw[c] * im.bh[c] * 64;
    const size_t plane_per_img = (plane_bytes + 15) & ~(size_t)15;
    std::vector<u8> plane(plane_per_img, 0x99);
    std::vector<u16> qtabs(192);
    memcpy(qtabs.data(), im.q, 192 * sizeof(u16));
    IdctArgs ia = {im, 1, coef.data(), coef_per_img, rowmask.data(), qtabs.data(), plane.data(), plane_per_img};
    const int gx = (int)((nblocks + 31) / 32);
    gridDim = cuemu::Idx{(unsigned)gx, 1, 1};
    for (int b = 0; b < gx; ++b) cuemu_run_block(256, b, body_idct, &ia, 0);
    for (size_t k = 0; k < coef_per_img; ++k) if (coef[k] != 0) return -7;       // the kernel leaves the coefficient buffer zero
    // colour: the 4:2:0 kernel where lsf_front_end_batch_jpeg uses it, else the generic one
    const size_t frame_bytes = (size_t)im.W * im.H * 3;
    ColorArgs ca = {im, 1, plane.data(), plane_per_img, bgr, frame_bytes};
    if (im.ncomp == 3 && im.hs[0] == 2 && im.vs[0] == 2 && (im.W & 7) == 0 && (frame_bytes & 7) == 0) {
        const int npx8 = im.H * (im.W / 8), cx = (npx8 + 255) / 256;
        gridDim = cuemu::Idx{(unsigned)cx, 1, 1};
        for (int b = 0; b < cx; ++b) cuemu_run_block(256, b, body_color420, &ca, 0);
    } else {
        const long long npx4 = (long long)im.H * ((im.W + 3) / 4);
        const int cx = (int)((npx4 + 255) / 256);
        gridDim = cuemu::Idx{(unsigned)cx, 1, 1};
        for (int b = 0; b < cx; ++b) cuemu_run_block(256, b, body_color, &ca, 0);
    }
    return 0;
}
'''


def kernel_text():
    text = open(os.path.join(CSRC, "k_jpeg.cu")).read()
    body = text[text.index("namespace lsf {"):text.index("}  // namespace lsf")]
    body, n = re.subn(r"__device__ __forceinline__ u32 lds32\(u32 a\) \{[^\n]*\}\n", "", body)
    assert n == 1, "the inline-PTX shared-memory load of k_jpeg.cu was not found"
    assert "asm volatile" not in body
    return body


def build(out_dir, jt=64, sanitize=False):
    """sanitize: False, True / "thread" (ThreadSanitizer) or "address" (AddressSanitizer + bounds)."""
    src = os.path.join(out_dir, "jpeg_huff_emu.cpp")
    with open(src, "w") as f:
        f.write(HARNESS.replace("@@KERNEL@@", kernel_text()))
    kind = "thread" if sanitize is True else sanitize
    so = os.path.join(out_dir, "libjhe%s.so" % ("_%s" % kind if kind else ""))
    cmd = ["g++", "-std=c++14", "-O1" if sanitize else "-O2", "-g", "-shared", "-fPIC", "-pthread", "-DLSF_JT=%d" % jt,
           "-I", os.path.join(ROOT, "oracle", "csrc"), "-I", CSRC, src, os.path.join(ROOT, "oracle", "csrc", "cuda_threads_emu.cpp"), "-o", so]
    if kind:
        cmd.insert(1, "-fsanitize=thread" if kind == "thread" else "-fsanitize=address,bounds")
    subprocess.check_call(cmd)
    return so
