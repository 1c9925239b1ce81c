// TEST INFRASTRUCTURE: runs the product's JPEG decode arithmetic (lane_slam_b200/csrc/jpeg_core.cuh, compiled for the host) the
// way the CUDA kernels orchestrate it -- stuffing removal, self-synchronising parallel Huffman decode simulated thread by
// thread, DC prediction, IDCT, fancy upsampling, colour conversion -- so that the algorithm can be checked against cv2.imdecode
// and the sequential oracle without a GPU.  Built and used by tests/test_jpeg_core.py.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../lane_slam_b200/csrc/jpeg_core.cuh"

// jpeg_scan_span of lane_slam_b200/csrc/k_jpeg.cu restated for the host, operation for operation (64-bit bit buffer refilled a
// word ahead, the two-level table with the search as the last resort, the table of a slot from the packed selector, the
// block-end selects).  On a wrong starting state it behaves differently from jd::decode_span (it always skips the value bits):
// the test is that the rounds still converge to the same, exact, image.
static void gpu_scan_span(bool WRITE, const uint32_t *words, const jd::Tabs &tabs, uint32_t tsel, int bpm, jd::Span &st, uint32_t limit,
                          int16_t *coef, uint32_t u_start, uint32_t max_blocks, int16_t *dcdiff, uint32_t *rowmask)
{
    uint32_t pos = st.pos;
    if (!(pos < limit)) return;
    uint32_t z = st.s & 63, slot = st.s >> 6, adv = st.adv;
    uint32_t wi = pos >> 5;
    const uint32_t sh = pos & 31;
    uint64_t w = ((((uint64_t)words[wi]) << 32) | words[wi + 1]) << sh;
    int have = 64 - (int)sh;
    wi += 2;
    uint32_t nx = words[wi];
    const uint32_t *lut = &tabs.lut[0][0], *lut2 = &tabs.lut2[0][0];
    const uint8_t *zz = jd::zigzag();
    uint32_t rm = 0, rm_blk = 0;
    while (pos < limit) {
        const uint32_t top = (uint32_t)(w >> 32);
        const uint32_t tab = (tsel >> (slot * 4 + (z ? 2 : 0))) & 3;
        uint32_t e = lut[(tab << jd::LUT_BITS) + (top >> (32 - jd::LUT_BITS))];
        if ((int32_t)e <= 0) {
            if (e) e = lut2[(e & 0xffffu) + ((top >> (32 - jd::LUT_BITS - 6)) & 63u)];
            if (e == 0) e = jd::search_code(tabs, (int)tab, top);
        }
        const uint32_t tot = (e >> 16) & 31, a = e >> 24;
        const uint32_t zn = z + a < 64u ? z + a : 64u;
        if (WRITE) {
            const uint32_t len = (e >> 8) & 31, size = tot - len;
            if (z == 0 || (size != 0 && z + a <= 64)) {
                const uint32_t bits = size ? (uint32_t)((w << len) >> (64 - size)) : 0;
                const int val = (size && bits < (1u << (size - 1))) ? (int)bits - (int)(1u << size) + 1 : (int)bits;
                const uint32_t blk = (u_start + adv) >> 6;
                if (blk < max_blocks) {
                    if (z == 0) dcdiff[blk] = (int16_t)val;
                    else {
                        const uint32_t nat = zz[z + a - 1];
                        coef[(size_t)blk * 64 + nat] = (int16_t)val;
                        if (blk != rm_blk) { jd::rowmask_flush(rowmask, rm_blk, rm); rm_blk = blk; rm = 0; }
                        rm |= 1u << (nat >> 3);
                    }
                }
            }
        }
        adv += zn - z;
        pos += tot; w <<= tot; have -= (int)tot;
        const bool endblk = zn == 64;
        const uint32_t nslot = slot + 1 == (uint32_t)bpm ? 0 : slot + 1;
        slot = endblk ? nslot : slot;
        z = endblk ? 0 : zn;
        if (have < 32) { w |= (uint64_t)nx << (32 - have); have += 32; nx = words[++wi]; }
    }
    if (WRITE) jd::rowmask_flush(rowmask, rm_blk, rm);
    st.pos = pos; st.s = (slot << 6) | z; st.adv = adv;
}

static int jpc_decode_mode(const uint8_t *data, size_t len, uint8_t *bgr, int sub_bytes, int *rounds_out, int *nsub_out, int mode);

extern "C" __attribute__((visibility("default")))
int jpc_decode(const uint8_t *data, size_t len, uint8_t *bgr, int sub_bytes, int *rounds_out, int *nsub_out)
{
    return jpc_decode_mode(data, len, bgr, sub_bytes, rounds_out, nsub_out, 0);
}

// mode 1: the kernel's own loop (gpu_scan_span), the kernel's subsequence length when sub_bytes == 0 (512 per image), 20 zero
// bytes behind the stream and 0xFF garbage behind those (the kernel's buffer holds a previous image there)
extern "C" __attribute__((visibility("default")))
int jpc_decode_gpu_loop(const uint8_t *data, size_t len, uint8_t *bgr, int sub_bytes, int *rounds_out, int *nsub_out)
{
    return jpc_decode_mode(data, len, bgr, sub_bytes, rounds_out, nsub_out, 1);
}

static int jpc_decode_mode(const uint8_t *data, size_t len, uint8_t *bgr, int sub_bytes, int *rounds_out, int *nsub_out, int mode)
{
    jd::Image im; jd::Tabs tabs;
    int rc = jd::parse(data, len, im, &tabs);
    if (rc) return rc;
    if (im.restart) return -2;
    // stuffing removal -> big-endian words (+ 3 zero words)
    std::vector<uint8_t> clean;
    clean.reserve(im.ent_len);
    const uint8_t *e = data + im.ent_off;
    for (uint32_t i = 0; i < im.ent_len; ++i) {
        if (e[i] == 0x00 && i > 0 && e[i - 1] == 0xFF) continue;
        clean.push_back(e[i]);
    }
    const uint32_t nbytes = (uint32_t)clean.size(), nwords = (nbytes + 3) / 4 + (mode ? 8 : 3);
    std::vector<uint32_t> words(nwords, 0);
    if (mode) {
        for (uint32_t i = nbytes + 20; i < nwords * 4; ++i) words[i >> 2] |= 0xFFu << (24 - 8 * (i & 3));
    }
    for (uint32_t i = 0; i < nbytes; ++i) words[i >> 2] |= (uint32_t)clean[i] << (24 - 8 * (i & 3));
    uint32_t tsel = 0;
    for (int k = 0; k < jd::MAX_BPM; ++k) tsel |= (((uint32_t)im.slot_dc[k] & 3u) | (((uint32_t)im.slot_ac[k] & 3u) << 2)) << (4 * k);
    const uint32_t total_bits = nbytes * 8;
    const uint32_t nblocks = (uint32_t)im.mcux * im.mcuy * im.bpm;
    // subsequences
    int S = sub_bytes > 0 ? sub_bytes : 64;
    S = (S + 3) & ~3;
    if (mode && sub_bytes <= 0) { const uint32_t s512 = ((nbytes + 511) / 512 + 3) & ~3u; S = (int)(s512 > 16 ? s512 : 16); }
    const int nsub = (int)((nbytes + S - 1) / S);
    std::vector<jd::Span> E(nsub), En(nsub);
    std::vector<uint8_t> dirty(nsub + 1, 0), dirty_next(nsub + 1, 0);
    auto limit = [&](int i) { uint32_t l = (uint32_t)(i + 1) * S * 8; return l < total_bits ? l : total_bits; };
    // phase A: guessed start (own first bit, DC of slot 0)
    for (int i = 0; i < nsub; ++i) {
        jd::Span st; st.pos = (uint32_t)i * S * 8; st.s = 0; st.adv = 0;
        if (mode) gpu_scan_span(false, words.data(), tabs, tsel, im.bpm, st, limit(i), nullptr, 0, 0, nullptr, nullptr);
        else jd::decode_span(words.data(), tabs, im.slot_dc, im.slot_ac, im.bpm, st, limit(i), nullptr, 0, 0);
        E[i] = st;
    }
    // phase B: re-decode from the left neighbour's end state until nothing changes
    for (int i = 1; i < nsub; ++i) dirty[i] = 1;
    int rounds = 0;
    while (true) {
        bool changed = false;
        En = E;
        std::fill(dirty_next.begin(), dirty_next.end(), 0);
        for (int i = 1; i < nsub; ++i) {
            if (!dirty[i]) continue;
            jd::Span st; st.pos = E[i - 1].pos; st.s = E[i - 1].s; st.adv = 0;
            if (mode) gpu_scan_span(false, words.data(), tabs, tsel, im.bpm, st, limit(i), nullptr, 0, 0, nullptr, nullptr);
            else jd::decode_span(words.data(), tabs, im.slot_dc, im.slot_ac, im.bpm, st, limit(i), nullptr, 0, 0);
            if (st.pos != E[i].pos || st.s != E[i].s) { dirty_next[i + 1] = 1; changed = true; }
            En[i] = st;
        }
        E = En; dirty = dirty_next;
        ++rounds;
        if (!changed) break;
        if (rounds > nsub + 2) return -5;
    }
    // phase C: output position of every subsequence
    std::vector<uint32_t> ustart(nsub, 0);
    for (int i = 1; i < nsub; ++i) ustart[i] = ustart[i - 1] + E[i - 1].adv;
    // phase D: write
    std::vector<int16_t> coef((size_t)nblocks * 64, 0);
    std::vector<uint32_t> rowmask(nblocks / 4 + 1, 0);
    std::vector<int16_t> dcdiff(nblocks, 0);
    for (int i = 0; i < nsub; ++i) {
        jd::Span st;
        if (i == 0) { st.pos = 0; st.s = 0; } else { st.pos = E[i - 1].pos; st.s = E[i - 1].s; }
        st.adv = 0;
        if (mode) gpu_scan_span(true, words.data(), tabs, tsel, im.bpm, st, limit(i), coef.data(), ustart[i], nblocks, dcdiff.data(), rowmask.data());
        else jd::decode_span(words.data(), tabs, im.slot_dc, im.slot_ac, im.bpm, st, limit(i), coef.data(), ustart[i], nblocks, nullptr, rowmask.data());
    }
    // DC prediction per component, in decode order (the kernel keeps the differences in their own array)
    int pred[3] = {0, 0, 0};
    for (uint32_t b = 0; b < nblocks; ++b) {
        const int c = im.slot_comp[b % im.bpm];
        pred[c] += mode ? dcdiff[b] : coef[(size_t)b * 64];
        coef[(size_t)b * 64] = (int16_t)pred[c];
    }
    // IDCT into component planes
    std::vector<std::vector<uint8_t>> plane(3);
    for (int c = 0; c < im.ncomp; ++c) plane[c].assign((size_t)im.bw[c] * im.bh[c] * 64, 0);
    for (uint32_t b = 0; b < nblocks; ++b) {
        const int slot = b % im.bpm, mcu = b / im.bpm, c = im.slot_comp[slot];
        const int bx = (mcu % im.mcux) * im.hs[c] + im.slot_bx[slot], by = (mcu / im.mcux) * im.vs[c] + im.slot_by[slot];
        // rows the decoder did not flag are not read (the GPU kernel leaves them out the same way): poison them to prove it
        const uint32_t rows = (rowmask[b >> 2] >> ((b & 3) * 8)) & 0xff;
        int16_t blkc[64];
        for (int k = 0; k < 64; ++k) blkc[k] = (k < 8 || ((rows >> (k >> 3)) & 1)) ? coef[(size_t)b * 64 + k] : (int16_t)0x5a5a;
        for (int k = 8; k < 64; ++k) if (!((rows >> (k >> 3)) & 1) && coef[(size_t)b * 64 + k] != 0) return -6;
        jd::idct_block_rows(blkc, im.q[c], rows, &plane[c][((size_t)by * 8) * (im.bw[c] * 8) + bx * 8], im.bw[c] * 8);
    }
    // upsample + colour
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
    if (rounds_out) *rounds_out = rounds;
    if (nsub_out) *nsub_out = nsub;
    return 0;
}

// geometry of a file as jd::parse sees it (callers size the output of jpc_decode* from it); returns parse()'s code
extern "C" __attribute__((visibility("default")))
int jpc_info(const uint8_t *data, size_t len, int *W, int *H)
{
    jd::Image im;
    const int rc = jd::parse(data, len, im, nullptr);
    if (rc == 0) { *W = im.W; *H = im.H; }
    return rc;
}
