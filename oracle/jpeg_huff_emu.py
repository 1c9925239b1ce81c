"""TEST INFRASTRUCTURE: builds a host library that runs the REAL source text of k_jpeg_huff (lane_slam_b200/csrc/k_jpeg.cu: byte
un-stuffing, the self-synchronising Huffman rounds with their compaction, prefix sums, coefficient writes, DC prediction) as one
thread block made of host threads (oracle/csrc/cuda_threads_emu.h), followed by the host IDCT / colour code of the shared header.
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

// data: one JPEG file.  bgr: [H][W][3] out.  Returns 0, < 0 for a file outside the decoder's scope, 100 + status if the kernel
// flagged the image.  prefill: the byte the stream buffer is filled with first (the kernel's buffer holds an older image).
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
    // k_jpeg_idct / k_jpeg_color on the host (the shared header's functions; rows the Huffman pass did not flag are not read)
    std::vector<std::vector<uint8_t>> plane(3);
    for (int c = 0; c < im.ncomp; ++c) plane[c].assign((size_t)im.bw[c] * im.bh[c] * 64, 0);
    for (size_t b = 0; b < nblocks; ++b) {
        const int slot = (int)(b % im.bpm), mcu = (int)(b / im.bpm), c = im.slot_comp[slot];
        const int bx = (mcu % im.mcux) * im.hs[c] + im.slot_bx[slot], by = (mcu / im.mcux) * im.vs[c] + im.slot_by[slot];
        const u32 rows = ((rowmask[b >> 2] >> ((b & 3) * 8)) & 0xfeu) | 1u;
        int16_t blk[64];
        for (int k = 0; k < 64; ++k) blk[k] = ((rows >> (k >> 3)) & 1) ? coef[b * 64 + k] : (int16_t)0x5a5a;
        for (int k = 8; k < 64; ++k) if (!((rows >> (k >> 3)) & 1) && coef[b * 64 + k] != 0) return -6;
        jd::idct_block_rows(blk, im.q[c], rows, &plane[c][((size_t)by * 8) * (im.bw[c] * 8) + bx * 8], im.bw[c] * 8);
    }
    for (int y = 0; y < im.H; ++y)
        for (int x = 0; x < im.W; ++x) {
            uint8_t *o = bgr + ((size_t)y * im.W + x) * 3;
            const int yy = plane[0][(size_t)y * im.bw[0] * 8 + x];
            if (im.ncomp == 1) { o[0] = o[1] = o[2] = (uint8_t)yy; continue; }
            int cc[3] = {yy, 0, 0};
            for (int c = 1; c < 3; ++c) {
                const int hs = im.hmax / im.hs[c], vs = im.vmax / im.vs[c];
                const int dw = (im.W * im.hs[c] + im.hmax - 1) / im.hmax, dh = (im.H * im.vs[c] + im.vmax - 1) / im.vmax;
                cc[c] = jd::chroma_at(plane[c].data(), im.bw[c] * 8, dw, dh, hs, vs, x, y);
            }
            jd::ycc_to_bgr(cc[0], cc[1], cc[2], o);
        }
    return 0;
}
'''


def kernel_text():
    text = open(os.path.join(CSRC, "k_jpeg.cu")).read()
    body = text[text.index("namespace lsf {"):text.index("// plane layout of one image")]
    body, n = re.subn(r"__device__ __forceinline__ u32 lds32\(u32 a\) \{[^\n]*\}\n", "", body)
    assert n == 1, "the inline-PTX shared-memory load of k_jpeg.cu was not found"
    assert "asm volatile" not in body
    return body


def build(out_dir, jt=64, sanitize=False):
    src = os.path.join(out_dir, "jpeg_huff_emu.cpp")
    with open(src, "w") as f:
        f.write(HARNESS.replace("@@KERNEL@@", kernel_text()))
    so = os.path.join(out_dir, "libjhe%s.so" % ("_tsan" if sanitize else ""))
    cmd = ["g++", "-std=c++14", "-O1" if sanitize else "-O2", "-g", "-shared", "-fPIC", "-pthread", "-DLSF_JT=%d" % jt,
           "-I", os.path.join(ROOT, "oracle", "csrc"), "-I", CSRC, src, os.path.join(ROOT, "oracle", "csrc", "cuda_threads_emu.cpp"), "-o", so]
    if sanitize:
        cmd.insert(1, "-fsanitize=thread")
    subprocess.check_call(cmd)
    return so
