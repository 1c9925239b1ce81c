// jpeg_core.cuh -- the arithmetic of the baseline JPEG decode shared by the CUDA kernels (k_jpeg.cu) and by the host-side
// simulation that tests it without a GPU (oracle/csrc/jpeg_parallel_check.cpp compiles this header with g++).
//
// Replaces cv2.imdecode(..., IMREAD_COLOR) = libjpeg-turbo at its defaults, the step before the line path
// (duckietown_utils/jpg.py:21-31, line_detector_node.py:155): Huffman decode -> dequantise -> jpeg_idct_islow ->
// fancy (triangle) chroma upsampling -> fixed-point YCbCr -> BGR.  All integer; results are bit-identical to cv2
// (oracle/csrc/jpeg_oracle.c is the independent restatement the tests compare with).
//
// Huffman decoding is the only sequential part.  It is parallelised inside one image the self-synchronising way: the
// entropy-coded bytes (stuffing removed) are cut into subsequences; every thread decodes its subsequence from a guessed state,
// then re-decodes from the end state of its left neighbour until nothing changes (Huffman streams re-synchronise within a few
// symbols, so two or three rounds suffice); a prefix sum over the numbers of coefficient slots each subsequence covers
// gives every thread its output position, and a last pass writes the coefficients.  Decoder state between symbols =
// (bit position, slot = block-in-MCU * 64 + zig-zag index).
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define JD_FN __host__ __device__ __forceinline__
#else
#define JD_FN static inline
#endif

namespace jd {

constexpr int MAX_BPM = 6;              // blocks per MCU: 4:2:0 = 4 + 1 + 1
#ifndef LSF_JPEG_LUT_BITS
#define LSF_JPEG_LUT_BITS 10
#endif
constexpr int LUT_BITS = LSF_JPEG_LUT_BITS;
static_assert(LUT_BITS == 10, "the second level assumes 10 + 6 = 16 bits");
constexpr int LUT2_TABLES = 32;         // the standard (Annex K) tables need 13
constexpr uint32_t LUT2_FLAG = 0x80000000u;

// One Huffman table set: [0] DC table 0, [1] DC table 1, [2] AC table 0, [3] AC table 1
struct Tabs {
    uint32_t lut[4][1 << LUT_BITS];     // look-ahead table of lut_entry() words; 0 = no code (or one the second level has no room for);
                                        // LUT2_FLAG | i * 64 = the codes with this prefix are longer: second-level table i
    uint32_t lut2[LUT2_TABLES][64];     // indexed by the six bits after the first LUT_BITS (LUT_BITS + 6 = 16 = the longest code)
    uint32_t n2;                        // second-level tables in use (all four tables of the set draw from one pool)
    int32_t maxcode[4][18];             // largest code of each length 1 .. 16 (-1: none)
    int32_t valoff[4][17];              // index of the first symbol of a length minus the smallest code of that length
    uint8_t vals[4][256];
};

// Geometry and table selection of one image (made by parse() on the host)
struct Image {
    uint32_t ent_off, ent_len;          // entropy-coded segment inside the file: first byte after SOS, bytes up to EOI (stuffed)
    int32_t W, H, ncomp;
    int32_t hs[3], vs[3], tq[3], td[3], ta[3];
    int32_t hmax, vmax, mcux, mcuy, bpm;            // MCUs per row / column, blocks per MCU
    int32_t slot_comp[MAX_BPM], slot_bx[MAX_BPM], slot_by[MAX_BPM];   // component and block offset inside the MCU of every slot
    int32_t slot_dc[MAX_BPM], slot_ac[MAX_BPM];     // indices into Tabs (0/1 DC, 2/3 AC)
    int32_t bw[3], bh[3];               // blocks per row / column of every component plane (padded to whole MCUs)
    uint16_t q[3][64];                  // quantisation table of every component, natural order
    int32_t restart;                    // DRI (0 = none; the parallel decoder supports 0 only)
};

#define JD_ZZ_INIT {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, \
                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63}
#ifdef __CUDACC__
static __device__ __constant__ uint8_t zz_dev[64] = JD_ZZ_INIT;
#endif
static const uint8_t zz_host[64] = JD_ZZ_INIT;
JD_FN const uint8_t *zigzag()
{
#ifdef __CUDA_ARCH__
    return zz_dev;
#else
    return zz_host;
#endif
}

// What one Huffman symbol does, precomputed: symbol | code length << 8 | (code length + value bits) << 16 | zig-zag advance << 24.
// The advance is 1 for a DC value, run + 1 for an AC value, 16 for ZRL and 64 for EOB (it is clamped to the end of the block).
JD_FN uint32_t lut_entry(int table, uint32_t len, uint32_t sym)
{
    uint32_t size, adv;
    if (table < 2) { size = sym > 15 ? 15u : sym; adv = 1; }
    else {
        size = sym & 15u;
        const uint32_t run = sym >> 4;
        adv = size ? run + 1 : (run == 15 ? 16u : 64u);
    }
    return sym | (len << 8) | ((len + size) << 16) | (adv << 24);
}

// the canonical search of jdhuff.c for a code of more than LUT_BITS bits at the top of v (or no code at all: 16 bits, symbol 0)
template <typename TabsT>
JD_FN uint32_t search_code(const TabsT &tabs, int t, uint32_t v)
{
    for (uint32_t l = LUT_BITS + 1; l <= 16; ++l) {
        const int32_t code = (int32_t)(v >> (32 - l));
        if (code <= tabs.maxcode[t][l]) return lut_entry(t, l, tabs.vals[t][(tabs.valoff[t][l] + code) & 255]);
    }
    return lut_entry(t, 16, 0);
}

// ---- host: header parsing and table construction ---------------------------------------------------------------------------
inline int rd16(const uint8_t *p) { return (p[0] << 8) | p[1]; }

inline void build_table(const uint8_t *bits /* [1..16] */, const uint8_t *vals, int nvals, Tabs &t, int slot)
{
    memset(t.lut[slot], 0, sizeof(t.lut[slot]));
    memcpy(t.vals[slot], vals, nvals);
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        t.valoff[slot][l] = k - code;
        for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
            if (l <= LUT_BITS) {
                const int lo = code << (LUT_BITS - l), n = 1 << (LUT_BITS - l);
                for (int j = 0; j < n; ++j) t.lut[slot][lo + j] = lut_entry(slot, (uint32_t)l, vals[k]);
            } else {
                const int rem = l - LUT_BITS, prefix = code >> rem;
                uint32_t &first = t.lut[slot][prefix];
                if (first == 0 && t.n2 < (uint32_t)LUT2_TABLES) first = LUT2_FLAG | (t.n2++ * 64);
                if (first & LUT2_FLAG) {
                    uint32_t *sub = &t.lut2[0][0] + (first & 0xffffu);
                    const int lo = (code & ((1 << rem) - 1)) << (6 - rem), n = 1 << (6 - rem);
                    for (int j = 0; j < n; ++j) sub[lo + j] = lut_entry(slot, (uint32_t)l, vals[k]);
                }
            }
        }
        t.maxcode[slot][l] = bits[l] ? code - 1 : -1;
        code <<= 1;
    }
    t.maxcode[slot][17] = 0x7fffffff;
    t.valoff[slot][0] = 0; t.maxcode[slot][0] = -1;
}

// Returns 0, or < 0: -1 malformed, -2 outside the scope (progressive / arithmetic / 12-bit / CMYK / sampling factors > 2).
// tabs may be NULL (geometry only); *dht_hash (optional) = FNV-1a over the bytes of all DHT segments, so that a caller decoding
// many files can skip rebuilding identical tables.
inline int parse(const uint8_t *data, size_t len, Image &im, Tabs *tabs_p, uint64_t *dht_hash = nullptr)
{
    memset(&im, 0, sizeof(im));
    if (tabs_p) memset(tabs_p, 0, sizeof(*tabs_p));
    uint64_t hsh = 1469598103934665603ull;
    uint16_t qt[4][64];
    memset(qt, 0, sizeof(qt));
    if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return -1;
    size_t i = 2;
    bool sof = false;
    int cid[3] = {0, 0, 0};
    while (i + 4 <= len) {
        if (data[i] != 0xFF) return -1;
        const int m = data[i + 1];
        if (m == 0xFF) { ++i; continue; }
        const int L = rd16(data + i + 2);
        const uint8_t *s = data + i + 4, *e = data + i + 2 + L;
        if (e > data + len) return -1;
        if (m == 0xDB) {
            while (s < e) {
                const int pq = s[0] >> 4, tq = s[0] & 15; ++s;
                if (tq > 3) return -1;
                for (int k = 0; k < 64; ++k) { qt[tq][zigzag()[k]] = (uint16_t)(pq ? rd16(s) : s[0]); s += pq ? 2 : 1; }
            }
        } else if (m == 0xC4) {
            for (const uint8_t *p = s; p < e; ++p) hsh = (hsh ^ *p) * 1099511628211ull;
            while (s < e) {
                const int tc = s[0] >> 4, th = s[0] & 15; ++s;
                if (th > 1 || tc > 1) return -2;
                uint8_t bits[17]; int n = 0;
                bits[0] = 0;
                for (int l = 1; l <= 16; ++l) { bits[l] = s[l - 1]; n += bits[l]; }
                s += 16;
                if (n > 256 || s + n > e) return -1;
                if (tabs_p) build_table(bits, s, n, *tabs_p, tc * 2 + th);
                s += n;
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (s[0] != 8) return -2;
            im.H = rd16(s + 1); im.W = rd16(s + 3); im.ncomp = s[5];
            if ((im.ncomp != 1 && im.ncomp != 3) || im.W <= 0 || im.H <= 0) return -2;
            im.hmax = im.vmax = 1;
            for (int c = 0; c < im.ncomp; ++c) {
                cid[c] = s[6 + 3 * c]; im.hs[c] = s[7 + 3 * c] >> 4; im.vs[c] = s[7 + 3 * c] & 15; im.tq[c] = s[8 + 3 * c];
                if (im.hs[c] < 1 || im.hs[c] > 2 || im.vs[c] < 1 || im.vs[c] > 2 || im.tq[c] > 3) return -2;
                if (im.hs[c] > im.hmax) im.hmax = im.hs[c];
                if (im.vs[c] > im.vmax) im.vmax = im.vs[c];
            }
            if (im.ncomp == 1) { im.hs[0] = im.vs[0] = 1; im.hmax = im.vmax = 1; }
            if (im.ncomp == 3 && (im.hs[1] != 1 || im.vs[1] != 1 || im.hs[2] != 1 || im.vs[2] != 1)) return -2;
            sof = true;
        } else if (m == 0xC2 || (m >= 0xC3 && m <= 0xCF && m != 0xC8 && m != 0xCC)) {
            return -2;
        } else if (m == 0xDD) {
            im.restart = rd16(s);
        } else if (m == 0xDA) {
            if (!sof || s[0] != im.ncomp) return -2;
            for (int k = 0; k < im.ncomp; ++k) {
                int c = -1;
                for (int q = 0; q < im.ncomp; ++q) if (cid[q] == s[1 + 2 * k]) c = q;
                if (c < 0) return -1;
                im.td[c] = s[2 + 2 * k] >> 4; im.ta[c] = s[2 + 2 * k] & 15;
                if (im.td[c] > 1 || im.ta[c] > 1) return -2;
            }
            im.mcux = (im.W + 8 * im.hmax - 1) / (8 * im.hmax); im.mcuy = (im.H + 8 * im.vmax - 1) / (8 * im.vmax);
            im.bpm = 0;
            for (int c = 0; c < im.ncomp; ++c) {
                im.bw[c] = im.mcux * im.hs[c]; im.bh[c] = im.mcuy * im.vs[c];
                memcpy(im.q[c], qt[im.tq[c]], sizeof(qt[0]));
                for (int by = 0; by < im.vs[c]; ++by)
                    for (int bx = 0; bx < im.hs[c]; ++bx) {
                        if (im.bpm >= MAX_BPM) return -2;
                        im.slot_comp[im.bpm] = c; im.slot_bx[im.bpm] = bx; im.slot_by[im.bpm] = by;
                        im.slot_dc[im.bpm] = im.td[c]; im.slot_ac[im.bpm] = 2 + im.ta[c];
                        ++im.bpm;
                    }
            }
            im.ent_off = (uint32_t)(e - data);
            // entropy-coded bytes run up to the EOI marker
            size_t end = len;
            while (end >= 2 && !(data[end - 2] == 0xFF && data[end - 1] == 0xD9)) --end;
            if (end < 2 || end - 2 < im.ent_off) return -1;
            im.ent_len = (uint32_t)(end - 2 - im.ent_off);
            if (dht_hash) *dht_hash = hsh;
            return 0;
        }
        i += 2 + (size_t)L;
    }
    return -1;
}

// ---- Huffman span decoder ----------------------------------------------------------------------------------------------------
// words: the entropy-coded bits with the stuffed zero bytes removed, as BIG-ENDIAN 32-bit words (bit 0 of the stream is bit 31
// of words[0]), followed by at least three zero words.
struct Span {
    uint32_t pos;       // bit position of the next symbol
    uint32_t s;         // slot * 64 + zig-zag index of the next coefficient (0 .. bpm*64 - 1)
    uint32_t adv;       // coefficient slots covered so far (64 per completed block)
};

// two spans can hold coefficients of the same block (the one that straddles their boundary): OR, atomically on the device
JD_FN void rowmask_flush(uint32_t *rowmask, uint32_t blk, uint32_t rm)
{
    rm &= 0xfeu;                                         // row 0 holds the DC value: always present
    if (!rm) return;
#if defined(__CUDA_ARCH__) || defined(JD_ATOMIC_ROWMASK)      /* the second: the kernel's source run on host threads (tests) */
    atomicOr(rowmask + (blk >> 2), rm << ((blk & 3) * 8));
#else
    rowmask[blk >> 2] |= rm << ((blk & 3) * 8);
#endif
}

// The loop keeps a 64-bit window of the stream in registers (the next word is fetched two words ahead, so no load sits on
// the symbol-to-symbol dependency chain) and decides with selects instead of branches (lanes of a warp decode different
// streams; only the rare long codes and the stores diverge).
template <typename TabsT>
JD_FN void decode_span(const uint32_t *words, const TabsT &tabs, const int32_t *slot_dc, const int32_t *slot_ac, int bpm, Span &st,
                       uint32_t pos_limit, int16_t *coef /* or NULL */, uint32_t u_start, uint32_t max_blocks,
                       int16_t *dcdiff = nullptr /* optional: DC differences go here (one per block) instead of coef[blk * 64] */,
                       uint32_t *rowmask = nullptr /* optional: one byte per block, bit r = row r (1..7) of the block holds a coefficient */,
                       const uint8_t *zz = nullptr /* optional: the caller's copy of the zig-zag table */)
{
    uint32_t pos = st.pos, s = st.s, adv = st.adv;
    if (!(pos < pos_limit)) return;
    const uint8_t *ZZ = zz ? zz : zigzag();
    uint32_t rm = 0, rm_blk = 0;
    uint32_t wi = pos >> 5;
    uint32_t hi = words[wi], lo = words[wi + 1], nx = words[wi + 2];
    uint32_t slot = s >> 6;
    int tdc = slot_dc[slot], tac = slot_ac[slot];
    while (pos < pos_limit) {
        const uint32_t sh = pos & 31;
        const uint32_t v = (uint32_t)(((((uint64_t)hi << 32) | lo) << sh) >> 32);      // the next 32 bits of the stream
        const uint32_t z = s & 63;
        const bool isdc = z == 0;
        const int t = isdc ? tdc : tac;
        uint32_t e = tabs.lut[t][v >> (32 - LUT_BITS)];
        if (e & LUT2_FLAG) e = (&tabs.lut2[0][0])[(e & 0xffffu) + ((v >> (32 - LUT_BITS - 6)) & 63u)];
        if (e == 0) e = search_code(tabs, t, v);        // no second-level table left for this prefix, or no such code
        const uint32_t len = (e >> 8) & 31, sym = e & 0xff;
        const uint32_t size = isdc ? (sym > 15 ? 15u : sym) : (sym & 15u);
        const uint32_t run = isdc ? 0u : (sym >> 4);
        const bool nosize = !isdc && size == 0;         // EOB or ZRL
        const bool eob = nosize && run != 15;
        uint32_t nz = z + run;                          // zig-zag index the value is written at (ZRL: z + 15, then + 1 below)
        const bool over = !eob && nz > 63;              // only reachable from a wrong starting state
        const bool endblk = eob || over;
        const bool write = !nosize && !over;
        nz = nz > 63 ? 63 : nz;
        if (write && coef) {
            const uint32_t bits = size ? ((v << len) >> (32 - size)) : 0;
            const int32_t val = (size && bits < (1u << (size - 1))) ? (int32_t)bits - (int32_t)(1u << size) + 1 : (int32_t)bits;
            const uint32_t blk = (u_start + adv) >> 6;  // adv counts the slots from the span's start
            if (blk < max_blocks) {
                if (isdc && dcdiff) dcdiff[blk] = (int16_t)val;
                else {
                    const uint32_t nat = ZZ[nz];
                    coef[(size_t)blk * 64 + nat] = (int16_t)val;
                    if (rowmask) {
                        if (blk != rm_blk) { rowmask_flush(rowmask, rm_blk, rm); rm_blk = blk; rm = 0; }
                        rm |= 1u << (nat >> 3);
                    }
                }
            }
        }
        pos += len + (write ? size : 0);
        const uint32_t znext = endblk ? 64 : nz + 1;
        adv += znext - z;
        if (znext >= 64) {
            slot = (slot + 1 == (uint32_t)bpm) ? 0 : slot + 1;
            tdc = slot_dc[slot]; tac = slot_ac[slot];
            s = slot << 6;
        } else s = (slot << 6) | znext;
        const uint32_t nwi = pos >> 5;
        if (nwi != wi) {                                // at most one word further (a symbol is at most 27 bits)
            wi = nwi; hi = lo; lo = nx;
            nx = words[wi + 2];
        }
    }
    if (rowmask) rowmask_flush(rowmask, rm_blk, rm);
    st.pos = pos; st.s = s; st.adv = adv;
}

// ---- IDCT (jidctint.c jpeg_idct_islow, CONST_BITS 13, PASS1_BITS 2) ------------------------------------------------------------
JD_FN int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
JD_FN uint8_t range_limit(int x)
{
    x &= 1023;
    if (x >= 512) x -= 1024;
    x += 128;
    return (uint8_t)(x < 0 ? 0 : x > 255 ? 255 : x);
}
JD_FN void idct_1d(int d0, int d1, int d2, int d3, int d4, int d5, int d6, int d7, int *o /* o[0..7] unscaled */)
{
    int z2 = d2, z3 = d6;
    int z1 = (z2 + z3) * 4433;
    int tmp2 = z1 + z3 * (-15137), tmp3 = z1 + z2 * 6270;
    int tmp0 = (d0 + d4) * 8192, tmp1 = (d0 - d4) * 8192;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = d7; tmp1 = d5; tmp2 = d3; tmp3 = d1;
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * 9633;
    tmp0 *= 2446; tmp1 *= 16819; tmp2 *= 25172; tmp3 *= 12299;
    z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    o[0] = tmp10 + tmp3; o[7] = tmp10 - tmp3; o[1] = tmp11 + tmp2; o[6] = tmp11 - tmp2;
    o[2] = tmp12 + tmp1; o[5] = tmp12 - tmp1; o[3] = tmp13 + tmp0; o[4] = tmp13 - tmp0;
}
// coef: 64 coefficients of one block (natural order), q: quantisation table; out: 8 rows of 8 samples, row stride `stride`
JD_FN void idct_block(const int16_t *coef, const uint16_t *q, uint8_t *out, int stride)
{
    int ws[64];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int o[8];
        idct_1d(coef[c] * q[c], coef[8 + c] * q[8 + c], coef[16 + c] * q[16 + c], coef[24 + c] * q[24 + c], coef[32 + c] * q[32 + c],
                coef[40 + c] * q[40 + c], coef[48 + c] * q[48 + c], coef[56 + c] * q[56 + c], o);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[8 * r + c] = descale(o[r], 11);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int o[8];
        idct_1d(ws[8 * r], ws[8 * r + 1], ws[8 * r + 2], ws[8 * r + 3], ws[8 * r + 4], ws[8 * r + 5], ws[8 * r + 6], ws[8 * r + 7], o);
#pragma unroll
        for (int c = 0; c < 8; ++c) out[r * stride + c] = range_limit(descale(o[c], 18));
    }
}

// the same with the rows that hold no coefficient left out (rows: bit r = row r may be non-zero; row 0 always is): exactly what
// idct_block computes when those rows are zero -- a column whose only entry is row 0 transforms to d0 << 2 in every row
JD_FN void idct_block_rows(const int16_t *coef, const uint16_t *q, uint32_t rows, uint8_t *out, int stride)
{
    int ws[64];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int d[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) d[r] = (r == 0 || ((rows >> r) & 1)) ? coef[8 * r + c] * q[8 * r + c] : 0;
        if ((rows & 0xfeu) == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) ws[8 * r + c] = d[0] << 2;
        } else {
            int o[8];
            idct_1d(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], o);
#pragma unroll
            for (int r = 0; r < 8; ++r) ws[8 * r + c] = descale(o[r], 11);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int o[8];
        idct_1d(ws[8 * r], ws[8 * r + 1], ws[8 * r + 2], ws[8 * r + 3], ws[8 * r + 4], ws[8 * r + 5], ws[8 * r + 6], ws[8 * r + 7], o);
#pragma unroll
        for (int c = 0; c < 8; ++c) out[r * stride + c] = range_limit(descale(o[c], 18));
    }
}

// ---- fancy upsampling + colour (jdsample.c h2v2 / h2v1 / h1v2 fancy upsample, jdcolor.c) -----------------------------------------
// chroma sample of plane `p` (row stride pw, real size dw x dh) at full-resolution pixel (x, y)
JD_FN int chroma_at(const uint8_t *p, int pw, int dw, int dh, int hs, int vs, int x, int y)
{
    if (hs == 1 && vs == 1) return p[(size_t)y * pw + x];
    if (hs == 2 && vs == 2) {
        const int sy = y >> 1, sx = x >> 1;
        int ny = (y & 1) ? sy + 1 : sy - 1;
        ny = ny < 0 ? 0 : (ny > dh - 1 ? dh - 1 : ny);
        const uint8_t *in0 = p + (size_t)sy * pw, *in1 = p + (size_t)ny * pw;
        const int cur = in0[sx] * 3 + in1[sx];
        if (x & 1) {
            if (sx + 1 > dw - 1) return (cur * 4 + 7) >> 4;
            return (cur * 3 + (in0[sx + 1] * 3 + in1[sx + 1]) + 7) >> 4;
        }
        if (sx == 0) return (cur * 4 + 8) >> 4;
        return (cur * 3 + (in0[sx - 1] * 3 + in1[sx - 1]) + 8) >> 4;
    }
    if (hs == 2) {   // h2v1
        const uint8_t *in = p + (size_t)y * pw;
        const int sx = x >> 1;
        if (x & 1) return sx + 1 > dw - 1 ? in[sx] : (in[sx] * 3 + in[sx + 1] + 2) >> 2;
        return sx == 0 ? in[sx] : (in[sx] * 3 + in[sx - 1] + 1) >> 2;
    }
    // h1v2
    const int sy = y >> 1;
    int ny = (y & 1) ? sy + 1 : sy - 1;
    ny = ny < 0 ? 0 : (ny > dh - 1 ? dh - 1 : ny);
    return (p[(size_t)sy * pw + x] * 3 + p[(size_t)ny * pw + x] + ((y & 1) ? 2 : 1)) >> 2;
}

JD_FN uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
JD_FN void ycc_to_bgr(int yy, int cb, int cr, uint8_t *bgr)
{
    const int xb = cb - 128, xr = cr - 128;
    bgr[2] = clamp8(yy + ((91881 * xr + 32768) >> 16));                          // FIX(1.40200)
    bgr[1] = clamp8(yy + ((-22554 * xb + 32768 + -46802 * xr) >> 16));           // FIX(0.34414), FIX(0.71414)
    bgr[0] = clamp8(yy + ((116130 * xb + 32768) >> 16));                         // FIX(1.77200)
}

}  // namespace jd
