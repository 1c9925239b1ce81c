// TEST INFRASTRUCTURE: runs the product's JPEG decode arithmetic (lane_slam_b200/csrc/jpeg_core.cuh, compiled for the host) the
// way the CUDA kernels orchestrate it -- stuffing removal, self-synchronising parallel Huffman decode simulated thread by
// thread, DC prediction, IDCT, fancy upsampling, colour conversion -- so that the algorithm can be checked against cv2.imdecode
// and the sequential oracle without a GPU.  Built and used by tests/test_jpeg_core.py.
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../lane_slam_b200/csrc/jpeg_core.cuh"

extern "C" __attribute__((visibility("default")))
int jpc_decode(const uint8_t *data, size_t len, uint8_t *bgr, int sub_bytes, int *rounds_out, int *nsub_out)
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
    const uint32_t nbytes = (uint32_t)clean.size(), nwords = (nbytes + 3) / 4 + 3;
    std::vector<uint32_t> words(nwords, 0);
    for (uint32_t i = 0; i < nbytes; ++i) words[i >> 2] |= (uint32_t)clean[i] << (24 - 8 * (i & 3));
    const uint32_t total_bits = nbytes * 8;
    const uint32_t nblocks = (uint32_t)im.mcux * im.mcuy * im.bpm;
    // subsequences
    int S = sub_bytes > 0 ? sub_bytes : 64;
    S = (S + 3) & ~3;
    const int nsub = (int)((nbytes + S - 1) / S);
    std::vector<jd::Span> E(nsub), En(nsub);
    std::vector<uint8_t> dirty(nsub + 1, 0), dirty_next(nsub + 1, 0);
    auto limit = [&](int i) { uint32_t l = (uint32_t)(i + 1) * S * 8; return l < total_bits ? l : total_bits; };
    // phase A: guessed start (own first bit, DC of slot 0)
    for (int i = 0; i < nsub; ++i) {
        jd::Span st; st.pos = (uint32_t)i * S * 8; st.s = 0; st.adv = 0;
        jd::decode_span(words.data(), tabs, im.slot_dc, im.slot_ac, im.bpm, st, limit(i), nullptr, 0, 0);
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
            jd::decode_span(words.data(), tabs, im.slot_dc, im.slot_ac, im.bpm, st, limit(i), nullptr, 0, 0);
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
    for (int i = 0; i < nsub; ++i) {
        jd::Span st;
        if (i == 0) { st.pos = 0; st.s = 0; } else { st.pos = E[i - 1].pos; st.s = E[i - 1].s; }
        st.adv = 0;
        jd::decode_span(words.data(), tabs, im.slot_dc, im.slot_ac, im.bpm, st, limit(i), coef.data(), ustart[i], nblocks, nullptr, rowmask.data());
    }
    // DC prediction per component, in decode order
    int pred[3] = {0, 0, 0};
    for (uint32_t b = 0; b < nblocks; ++b) {
        const int c = im.slot_comp[b % im.bpm];
        pred[c] += coef[(size_t)b * 64];
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
