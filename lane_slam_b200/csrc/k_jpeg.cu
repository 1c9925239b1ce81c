// k_jpeg.cu -- baseline JPEG decode on the GPU, bit-identical to cv2.imdecode (libjpeg-turbo defaults): the step before the
// line path (duckietown_utils/jpg.py:21-31, line_detector_node.py:155).  A frame crosses PCIe as ~75 KB of JPEG instead of
// 921.6 KB of BGR.  Arithmetic in jpeg_core.cuh (shared with the host-side simulation the CPU tests run).
//   k_jpeg_huff   one CTA per image: stuffing removal, self-synchronising parallel Huffman decode (subsequence per thread),
//                 prefix sum of the coefficient slots, coefficient write, DC prediction
//   k_jpeg_idct   one thread per 8x8 block: dequantise + jpeg_idct_islow -> component planes
//   k_jpeg_color  fancy upsampling + YCbCr -> BGR, 4 pixels per thread, into the frame buffer the front end reads
// and lsf_front_end_batch_jpeg (the C ABI entry): header parsing on the host (a few hundred bytes per file), one H2D copy of
// the concatenated files, the three kernels, then the ordinary batch pipeline on the decoded frames (device resident).
#include <vector>

#include "ctx.cuh"
#include "jpeg_core.cuh"

namespace lsf {

#ifndef LSF_JT
#define LSF_JT 512
#endif
constexpr int JT = LSF_JT;               // threads (= subsequences) per image

struct JpegItem {
    u32 ent_off, ent_len;                // entropy-coded bytes inside the blob
    u32 tabs, pad;                       // index of the Huffman table set
};

struct JpegState {
    u8 *blob; size_t blob_cap;
    u32 *clean; size_t clean_words;      // per image: stuffing-free big-endian words
    int16_t *coef; size_t coef_per_img;  // per image: nblocks * 64
    int16_t *dcdiff;                     // per image: nblocks DC differences (decode order)
    u32 *rowmask;                        // per image: one byte per block, bit r = row r of the block holds a coefficient
    bool coef_dirty;                     // the coefficient buffer may hold values (a batch ended between the Huffman pass and the IDCT)
    u8 *plane; size_t plane_per_img;     // per image: Y, Cb, Cr planes back to back (padded to whole MCUs)
    JpegItem *items, *h_items;
    u16 *qtabs, *h_qtabs;                // [n][3][64]
    jd::Tabs *tabsets, *h_tabsets; int tabs_cap;
    int *status, *h_status;
    int n_cap, W, H;
};

__device__ __forceinline__ int block_excl_scan(int v, int *sm, int &total)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += x;
    }
    __syncthreads();
    if (lane == 31) sm[warp] = incl;
    __syncthreads();
    int off = 0, tot = 0;
    for (int k = 0; k < JT / 32; ++k) { const int w = sm[k]; if (k < warp) off += w; tot += w; }
    total = tot;
    return off + incl - v;
}

// a code the two table levels do not hold (a table set with more long prefixes than the second level has room for, or no code
// at all: zero padding, a wrong starting state): the canonical search of jdhuff.c
__device__ __noinline__ u32 jpeg_long_code(const jd::Tabs &tabs, int t, u32 top) { return jd::search_code(tabs, t, top); }

// One pass over the bits [st.pos, limit) of an image's stream, from the state st: what jd::decode_span does on every state a
// valid stream can be in (the host-side simulation the CPU tests run uses that one), arranged for the GPU -- a left-aligned
// 64-bit bit buffer in registers refilled a word ahead, one table word per symbol (jd::lut_entry) that already holds the bits
// to skip and the zig-zag advance, so the symbol-to-symbol chain is a shift, a shared-memory load and a handful of adds.
// WRITE = false: only the state moves (phases A and B).  WRITE = true: the values go to coef / dcdiff / rowmask (phase D).
__device__ __forceinline__ u32 lds32(u32 a) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

// tsel: which look-ahead table a slot uses, four bits per slot -- bits 0..1 the DC table, bits 2..3 the AC table (so the table
// of the next symbol is a shift and a mask away, not a load); lut_s: shared-memory byte address of tabs.lut
template <bool WRITE>
__device__ __forceinline__ void jpeg_scan_span(const u32 *__restrict__ words, const jd::Tabs &tabs, u32 tsel, u32 lut_s, u32 lut2_s, int bpm, jd::Span &st, u32 limit,
                                               int16_t *__restrict__ coef, u32 u_start, u32 max_blocks, int16_t *__restrict__ dcdiff,
                                               u32 *__restrict__ rowmask, const u8 *zz)
{
    u32 pos = st.pos;
    if (!(pos < limit)) return;
    u32 z = st.s & 63, slot = st.s >> 6, adv = st.adv;
    u32 wi = pos >> 5;
    const u32 sh = pos & 31;
    unsigned long long w = ((((unsigned long long)words[wi]) << 32) | words[wi + 1]) << sh;
    int have = 64 - (int)sh;                            // valid bits in w (kept >= 32)
    wi += 2;
    u32 nx = words[wi];
    u32 rm = 0, rm_blk = 0;
    while (pos < limit) {
        const u32 top = (u32)(w >> 32);
        const u32 tab = (tsel >> (slot * 4 + (z ? 2 : 0))) & 3;
        u32 e = lds32(lut_s + (tab << (jd::LUT_BITS + 2)) + ((top >> (30 - jd::LUT_BITS)) & ~3u));
        if ((int)e <= 0) {                              // longer than LUT_BITS: 2 - 3 % of the symbols, but every other warp iteration
            if (e) e = lds32(lut2_s + (((e & 0xffffu) + ((top >> (32 - jd::LUT_BITS - 6)) & 63u)) << 2));
            if (e == 0) e = jpeg_long_code(tabs, (int)tab, top);
        }
        const u32 tot = (e >> 16) & 31, a = e >> 24;
        const u32 zn = min(z + a, 64u);
        if (WRITE) {
            const u32 len = (e >> 8) & 31, size = tot - len;
            if (z == 0 || (size != 0 && z + a <= 64)) {             // a DC value, or an AC value inside the block
                const u32 bits = size ? (u32)((w << len) >> (64 - size)) : 0;
                const int val = (size && bits < (1u << (size - 1))) ? (int)bits - (int)(1u << size) + 1 : (int)bits;
                const u32 blk = (u_start + adv) >> 6;
                if (blk < max_blocks) {
                    if (z == 0) dcdiff[blk] = (int16_t)val;
                    else {
                        const u32 nat = zz[z + a - 1];
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
        const u32 nslot = slot + 1 == (u32)bpm ? 0 : slot + 1;
        slot = endblk ? nslot : slot;
        z = endblk ? 0 : zn;
        if (have < 32) { w |= (unsigned long long)nx << (32 - have); have += 32; nx = words[++wi]; }
    }
    if (WRITE) jd::rowmask_flush(rowmask, rm_blk, rm);
    st.pos = pos; st.s = (slot << 6) | z; st.adv = adv;
}

__global__ void __launch_bounds__(JT) k_jpeg_huff(jd::Image g, const u8 *__restrict__ blob, const JpegItem *__restrict__ items,
                                                 const jd::Tabs *__restrict__ tabsets, u32 *__restrict__ clean_all, size_t clean_words,
                                                 int16_t *__restrict__ coef_all, size_t coef_per_img, int16_t *__restrict__ dcdiff_all,
                                                 u32 *__restrict__ rowmask_all, int *__restrict__ status)
{
    __shared__ jd::Tabs tabs;
    __shared__ jd::Span E[2][JT];
    __shared__ int s_scan[JT / 32];
    __shared__ u8 s_chg[JT];
    __shared__ u8 s_zz[64];                      // per-lane indices: shared memory, not the constant bank
    __shared__ u16 s_list[JT];                   // phase B: the subsequences to decode again, compacted
    const int t = threadIdx.x, img = blockIdx.x;
    const JpegItem it = items[img];
#ifdef LSF_JPEG_PROF
    long long pc[8]; int pci = 0;
#define JPROF() do { __syncthreads(); pc[pci++] = clock64(); } while (0)
#else
#define JPROF() do { } while (0)
#endif
    JPROF();
    {
        const u32 *src = reinterpret_cast<const u32 *>(tabsets + it.tabs);
        u32 *dst = reinterpret_cast<u32 *>(&tabs);
        for (int i = t; i < (int)(sizeof(jd::Tabs) / 4); i += JT) dst[i] = src[i];
        if (t < 64) s_zz[t] = jd::zigzag()[t];
    }
    const u32 lut_s = (u32)__cvta_generic_to_shared(&tabs.lut[0][0]), lut2_s = (u32)__cvta_generic_to_shared(&tabs.lut2[0][0]);
    u32 tsel = 0;
    for (int k = 0; k < jd::MAX_BPM; ++k) tsel |= (((u32)g.slot_dc[k] & 3u) | (((u32)g.slot_ac[k] & 3u) << 2)) << (4 * k);
    // ---- phase 0: remove the stuffed zero bytes (a 0x00 that follows a 0xFF), write big-endian words ----
    const u8 *raw = blob + it.ent_off;
    const u32 len = it.ent_len;
    u32 *words = clean_all + (size_t)img * clean_words;
    u8 *cbytes = reinterpret_cast<u8 *>(words);
    // a warp takes a run of 128-byte steps, a lane 4 bytes of each; a byte goes when it is 0x00 and follows a 0xFF
    constexpr int NW = JT / 32;
    const int lane = t & 31, warp = t >> 5;
    const u32 CW = ((len + NW - 1) / NW + 127) & ~127u;
    const u32 w0 = min(len, (u32)warp * CW), w1 = min(len, w0 + CW);
    auto load4 = [&](u32 base, u32 &b4, u32 &drop) {          // b4: the lane's bytes (first byte in bits 0..7); drop: bit i = byte i goes
        const u32 j = base + lane * 4;
        u32 v = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) v |= (j + i < w1 ? (u32)raw[j + i] : 0xAAu) << (8 * i);
        u32 prev = __shfl_up_sync(0xffffffffu, v >> 24, 1);
        if (lane == 0) prev = base > 0 ? raw[base - 1] : 0;
        const u32 pv = (v << 8) | prev;                        // byte i of pv = the byte before byte i of v
        drop = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) drop |= ((((v >> (8 * i)) & 0xff) == 0 && ((pv >> (8 * i)) & 0xff) == 0xff && j + i < w1) ? 1u : 0u) << i;
        b4 = v;
    };
    int removed = 0;                                           // warp-uniform
    for (u32 base = w0; base < w1; base += 128) {
        u32 b4, drop;
        load4(base, b4, drop);
        removed += __reduce_add_sync(0xffffffffu, __popc(drop));
    }
    if (lane == 0) s_scan[warp] = removed;
    __syncthreads();
    int before = 0, total_removed = 0;
    for (int k = 0; k < NW; ++k) { const int w = s_scan[k]; if (k < warp) before += w; total_removed += w; }
    {
        u32 k0 = w0 - (u32)before;                             // where the warp's next kept byte goes
        for (u32 base = w0; base < w1; base += 128) {
            u32 b4, drop;
            load4(base, b4, drop);
            const u32 j = base + lane * 4;
            const int kept = (int)min(4u, w1 > j ? w1 - j : 0u) - __popc(drop);
            int incl = kept;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int x = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += x; }
            u32 k = k0 + (u32)(incl - kept);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (j + i < w1 && !((drop >> i) & 1)) { cbytes[k ^ 3u] = (u8)(b4 >> (8 * i)); ++k; }
            k0 += (u32)__shfl_sync(0xffffffffu, incl, 31);
        }
    }
    const u32 nb = len - (u32)total_removed;           // stuffing-free bytes
    if (t < 20) { const u32 k = nb + t; if ((size_t)k < clean_words * 4) cbytes[k ^ 3u] = 0; }   // zero tail: the bit buffer reads up to four words ahead
    __syncthreads();
    JPROF();
    // ---- phase A: every thread decodes its subsequence from a guessed state (own first bit, DC of slot 0) ----
    const u32 S = max(16u, ((nb + JT - 1) / JT + 3) & ~3u);
    const int nsub = (int)((nb + S - 1) / S);
    const u32 total_bits = nb * 8;
    const u32 limit = min(total_bits, (u32)(t + 1) * S * 8);
    {
        jd::Span mine; mine.pos = (u32)t * S * 8; mine.s = 0; mine.adv = 0;
        if (t < nsub) jpeg_scan_span<false>(words, tabs, tsel, lut_s, lut2_s, g.bpm, mine, limit, nullptr, 0u, 0u, nullptr, nullptr, s_zz);
        E[0][t] = mine;
    }
    JPROF();
    // ---- phase B: decode again from the left neighbour's end state wherever that state changed, until nothing changes.  The
    // subsequences to redo are compacted every round (thread i takes the i-th of them): after two rounds most are settled and
    // the rest are scattered, one or two per warp.
    int cur = 0, rounds = 0;
    bool dirty = t >= 1 && t < nsub;
    while (true) {
        const u32 bal = __ballot_sync(0xffffffffu, dirty);
        if (lane == 0) s_scan[warp] = __popc(bal);
        E[cur ^ 1][t] = E[cur][t];
        __syncthreads();
        s_chg[t] = 0;                                  // only now: before the barrier a slower warp may still be reading the previous
                                                       // round's flag of its left neighbour (the `dirty` line at the end of the loop)
        int off = 0, nd = 0;
        for (int k = 0; k < NW; ++k) { const int w = s_scan[k]; if (k < warp) off += w; nd += w; }
        if (dirty) s_list[off + __popc(bal & ((1u << lane) - 1))] = (u16)t;
        __syncthreads();
        if (nd == 0) break;
#ifdef LSF_JPEG_PROF
        const long long r0 = clock64();
#endif
        if (t < nd) {
            const int q = s_list[t];
            jd::Span st; st.pos = E[cur][q - 1].pos; st.s = E[cur][q - 1].s; st.adv = 0;
            jpeg_scan_span<false>(words, tabs, tsel, lut_s, lut2_s, g.bpm, st, min(total_bits, (u32)(q + 1) * S * 8), nullptr, 0u, 0u, nullptr, nullptr, s_zz);
            s_chg[q] = (st.pos != E[cur][q].pos || st.s != E[cur][q].s) ? 1 : 0;
            E[cur ^ 1][q] = st;
        }
        __syncthreads();
#ifdef LSF_JPEG_PROF
        if (t == 0 && img == 0) printf("  round %d: %d dirty, %lld cycles\n", rounds, nd, clock64() - r0);
#endif
        dirty = t >= 1 && t < nsub && s_chg[t - 1];
        cur ^= 1;
        if (++rounds > JT + 2) { if (t == 0) status[img] = 1; break; }
    }
    __syncthreads();
    JPROF();
    // ---- phase C: first coefficient slot of every subsequence ----
    int tot_adv;
    const u32 ustart = (u32)block_excl_scan(t < nsub ? (int)E[cur][t].adv : 0, s_scan, tot_adv);
    // ---- phase D: decode once more from the true state, writing the coefficients (DC entries hold differences) ----
    const u32 nblocks = (u32)g.mcux * g.mcuy * g.bpm;
    int16_t *coef = coef_all + (size_t)img * coef_per_img;
    if (t < nsub) {
        jd::Span st;
        if (t == 0) { st.pos = 0; st.s = 0; } else { st.pos = E[cur][t - 1].pos; st.s = E[cur][t - 1].s; }
        st.adv = 0;
        jpeg_scan_span<true>(words, tabs, tsel, lut_s, lut2_s, g.bpm, st, limit, coef, ustart, nblocks, dcdiff_all + (size_t)img * (coef_per_img / 64),
                             rowmask_all + (size_t)img * (coef_per_img / 256), s_zz);
    }
    __syncthreads();
    JPROF();
    // ---- DC prediction: running sum of the differences per component, in decode order ----
    {
        const int nm = g.mcux * g.mcuy, per = (nm + JT - 1) / JT;
        const int m0 = min(nm, t * per), m1 = min(nm, m0 + per);
        const int16_t *dd = dcdiff_all + (size_t)img * (coef_per_img / 64);       // DC differences, one per block, decode order
        int sum[3] = {0, 0, 0};
        for (int m = m0; m < m1; ++m)
            for (int s = 0; s < g.bpm; ++s) sum[g.slot_comp[s]] += dd[m * g.bpm + s];
        int base[3], dummy;
        for (int c = 0; c < 3; ++c) base[c] = block_excl_scan(sum[c], s_scan, dummy);
        for (int m = m0; m < m1; ++m)
            for (int s = 0; s < g.bpm; ++s) {
                const int c = g.slot_comp[s];
                base[c] += dd[m * g.bpm + s];
                coef[((size_t)m * g.bpm + s) * 64] = (int16_t)base[c];
            }
    }
#ifdef LSF_JPEG_PROF
    JPROF();
    if (t == 0 && (img == 0 || img == 500))
        printf("jpeg prof img %d: unstuff %lld  A %lld  B %lld (%d rounds)  CD %lld  dc %lld  [cycles] len %u S %u\n", img, pc[1] - pc[0],
               pc[2] - pc[1], pc[3] - pc[2], rounds, pc[4] - pc[3], pc[5] - pc[4], len, S);
#endif
}

// plane layout of one image: component c at plane_off[c], row stride bw[c] * 8
__device__ __forceinline__ size_t plane_off(const jd::Image &g, int c)
{
    size_t o = 0;
    for (int k = 0; k < c; ++k) o += (size_t)g.bw[k] * g.bh[k] * 64;
    return o;
}

// 8 threads per block: thread j runs the column pass on column j, the 8 x 8 intermediate goes through shared memory (rows padded
// to 9 words), then the row pass on row j and one 8-byte store.  256 threads = 32 blocks per CTA.
// Only the rows the Huffman pass flagged are read (rowmask; row 0 always), and they are zeroed behind the read: the coefficient
// buffer is all zero again when the kernel ends, so no batch pays for clearing 900 KB per frame.
__global__ void __launch_bounds__(256) k_jpeg_idct(jd::Image g, int n, int16_t *__restrict__ coef_all, size_t coef_per_img,
                                                  const u32 *__restrict__ rowmask_all, const u16 *__restrict__ qtabs, u8 *__restrict__ plane_all,
                                                  size_t plane_per_img)
{
    __shared__ int ws[32][8][9];
    const int nblocks = g.mcux * g.mcuy * g.bpm;
    const int j = threadIdx.x & 7, lb = threadIdx.x >> 3;
    const int img = blockIdx.y;
    int tt = blockIdx.x * 32 + lb;
    const bool ok = tt < nblocks;
    int c = 0, by = 0, bx = 0;
    int16_t *cf = coef_all;
    bool zrow = false;
    if (ok) {
        // enumerate blocks plane by plane (neighbouring groups write neighbouring blocks of a plane row)
        while (c < g.ncomp - 1 && tt >= g.bw[c] * g.bh[c]) { tt -= g.bw[c] * g.bh[c]; ++c; }
        by = tt / g.bw[c]; bx = tt - by * g.bw[c];
        int slot0 = 0;
        for (int k = 0; k < c; ++k) slot0 += g.hs[k] * g.vs[k];
        const int mcu = (by / g.vs[c]) * g.mcux + (bx / g.hs[c]);
        const int slot = slot0 + (by % g.vs[c]) * g.hs[c] + (bx % g.hs[c]);
        cf = coef_all + (size_t)img * coef_per_img + ((size_t)mcu * g.bpm + slot) * 64;
        const u16 *q = qtabs + ((size_t)img * 3 + c) * 64;
        const u32 blk = (u32)mcu * g.bpm + slot;
        const u32 rows = ((__ldg(rowmask_all + (size_t)img * (coef_per_img / 256) + (blk >> 2)) >> ((blk & 3) * 8)) & 0xfeu) | 1u;
        int d[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) d[r] = ((rows >> r) & 1) ? (int)cf[8 * r + j] : 0;      // predicated loads, all in flight together
#pragma unroll
        for (int r = 0; r < 8; ++r) d[r] *= (int)__ldg(q + 8 * r + j);
        zrow = (rows >> j) & 1;
        if (rows == 1) {
#pragma unroll
            for (int r = 0; r < 8; ++r) ws[lb][r][j] = d[0] << 2;      // what the column pass gives when only row 0 is there
        } else {
            int o[8];
            jd::idct_1d(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], o);
#pragma unroll
            for (int r = 0; r < 8; ++r) ws[lb][r][j] = jd::descale(o[r], 11);
        }
    }
    __syncwarp();                                      // orders the loads above before the stores below, too
    if (zrow) *reinterpret_cast<uint4 *>(cf + 8 * j) = make_uint4(0, 0, 0, 0);       // one 16-byte row per thread
    if (ok) {
        const int *w = ws[lb][j];
        int o[8];
        jd::idct_1d(w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], o);
        u32 lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            lo |= (u32)jd::range_limit(jd::descale(o[k], 18)) << (8 * k);
            hi |= (u32)jd::range_limit(jd::descale(o[4 + k], 18)) << (8 * k);
        }
        const int stride = g.bw[c] * 8;
        u8 *dst = plane_all + (size_t)img * plane_per_img + plane_off(g, c) + ((size_t)by * 8 + j) * stride + bx * 8;
        *reinterpret_cast<uint2 *>(dst) = make_uint2(lo, hi);
    }
}

__global__ void __launch_bounds__(256) k_jpeg_color(jd::Image g, int n, const u8 *__restrict__ plane_all, size_t plane_per_img,
                                                   u8 *__restrict__ bgr, size_t frame_bytes)
{
    const int W = g.W, H = g.H, W4 = (W + 3) / 4;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * H * W4) return;
    const int img = (int)(gid / ((long long)H * W4));
    const int rem = (int)(gid - (long long)img * H * W4);
    const int y = rem / W4, x0 = (rem - y * W4) * 4;
    const u8 *pl = plane_all + (size_t)img * plane_per_img;
    const u8 *py = pl, *pcb = pl + plane_off(g, 1), *pcr = pl + plane_off(g, 2);
    __align__(16) u8 out[12];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = min(x0 + i, W - 1);
        const int yy = py[(size_t)y * g.bw[0] * 8 + x];
        if (g.ncomp == 1) { out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = (u8)yy; continue; }
        const int hs1 = g.hmax / g.hs[1], vs1 = g.vmax / g.vs[1];
        const int dw = (W * g.hs[1] + g.hmax - 1) / g.hmax, dh = (H * g.vs[1] + g.vmax - 1) / g.vmax;
        const int cb = jd::chroma_at(pcb, g.bw[1] * 8, dw, dh, hs1, vs1, x, y);
        const int cr = jd::chroma_at(pcr, g.bw[2] * 8, dw, dh, hs1, vs1, x, y);
        jd::ycc_to_bgr(yy, cb, cr, out + 3 * i);
    }
    u8 *dst = bgr + (size_t)img * frame_bytes + ((size_t)y * W + x0) * 3;
    if (x0 + 3 < W && (W & 3) == 0) {
        u32 *d32 = reinterpret_cast<u32 *>(dst);
        d32[0] = reinterpret_cast<const u32 *>(out)[0]; d32[1] = reinterpret_cast<const u32 *>(out)[1]; d32[2] = reinterpret_cast<const u32 *>(out)[2];
    } else {
        for (int i = 0; i < 4 && x0 + i < W; ++i) { dst[3 * i] = out[3 * i]; dst[3 * i + 1] = out[3 * i + 1]; dst[3 * i + 2] = out[3 * i + 2]; }
    }
}

// 4:2:0, 8 pixels per thread (W % 8 == 0): the six vertical sums 3 * near + far of each chroma plane are formed once and shared
// by the eight pixels; aligned word loads for luma and chroma, three 8-byte stores.
__global__ void __launch_bounds__(256) k_jpeg_color_420(jd::Image g, int n, const u8 *__restrict__ plane_all, size_t plane_per_img,
                                                       u8 *__restrict__ bgr, size_t frame_bytes)
{
    const int W = g.W, H = g.H, W8 = W / 8;
    const int img = blockIdx.y;
    const int rem = blockIdx.x * blockDim.x + threadIdx.x;
    if (rem >= H * W8) return;
    const int y = rem / W8, x0 = (rem - y * W8) * 8;
    const u8 *pl = plane_all + (size_t)img * plane_per_img;
    const int ys = g.bw[0] * 8, cs = g.bw[1] * 8;
    const int dw = (W + 1) / 2, dh = (H + 1) / 2;
    const int sy = y >> 1, sx = x0 >> 1;
    int ny = (y & 1) ? sy + 1 : sy - 1;
    ny = ny < 0 ? 0 : (ny > dh - 1 ? dh - 1 : ny);
    const uint2 yv = *reinterpret_cast<const uint2 *>(pl + (size_t)y * ys + x0);
    int col[2][6];       // vertical sums of chroma columns sx-1 .. sx+4 (edge columns are only used inside the image)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const u8 *p = pl + plane_off(g, 1 + c);
        const u8 *r0 = p + (size_t)sy * cs + sx, *r1 = p + (size_t)ny * cs + sx;
        const u32 a = *reinterpret_cast<const u32 *>(r0), b = *reinterpret_cast<const u32 *>(r1);
#pragma unroll
        for (int k = 0; k < 4; ++k) col[c][1 + k] = (int)((a >> (8 * k)) & 0xff) * 3 + (int)((b >> (8 * k)) & 0xff);
        col[c][0] = sx > 0 ? r0[-1] * 3 + r1[-1] : 0;
        col[c][5] = sx + 4 < dw ? r0[4] * 3 + r1[4] : 0;
    }
    __align__(8) u8 out[24];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int yy = (int)(((i < 4 ? yv.x : yv.y) >> (8 * (i & 3))) & 0xff);
        const int k = 1 + (i >> 1);
        int cc[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int cur = col[c][k];
            if (i & 1) cc[c] = (sx + (i >> 1) + 1 > dw - 1) ? (cur * 4 + 7) >> 4 : (cur * 3 + col[c][k + 1] + 7) >> 4;
            else cc[c] = (sx + (i >> 1) == 0) ? (cur * 4 + 8) >> 4 : (cur * 3 + col[c][k - 1] + 8) >> 4;
        }
        jd::ycc_to_bgr(yy, cc[0], cc[1], out + 3 * i);
    }
    uint2 *dst = reinterpret_cast<uint2 *>(bgr + (size_t)img * frame_bytes + ((size_t)y * W + x0) * 3);
    dst[0] = reinterpret_cast<const uint2 *>(out)[0]; dst[1] = reinterpret_cast<const uint2 *>(out)[1]; dst[2] = reinterpret_cast<const uint2 *>(out)[2];
}

}  // namespace lsf

using namespace lsf;

void jpeg_destroy(lsf_ctx *ctx)
{
    JpegState *j = (JpegState *)ctx->jpeg;
    if (!j) return;
    for (void *p : {(void *)j->blob, (void *)j->clean, (void *)j->coef, (void *)j->dcdiff, (void *)j->rowmask, (void *)j->plane, (void *)j->items, (void *)j->qtabs, (void *)j->tabsets,
                    (void *)j->status})
        if (p) cudaFree(p);
    for (void *p : {(void *)j->h_items, (void *)j->h_qtabs, (void *)j->h_tabsets, (void *)j->h_status}) if (p) cudaFreeHost(p);
    delete j;
    ctx->jpeg = nullptr;
}

static int jpeg_reserve(lsf_ctx *ctx, const jd::Image &g, int n, size_t blob_bytes)
{
    JpegState *j = (JpegState *)ctx->jpeg;
    if (!j) {
        j = new JpegState();
        memset(j, 0, sizeof(*j));
        ctx->jpeg = j;
    }
    const size_t nblocks = (size_t)g.mcux * g.mcuy * g.bpm;
    size_t plane = 0;
    for (int c = 0; c < g.ncomp; ++c) plane += (size_t)g.bw[c] * g.bh[c] * 64;
    if (n > j->n_cap || g.W != j->W || g.H != j->H || nblocks * 64 > j->coef_per_img || plane > j->plane_per_img) {
        for (void *p : {(void *)j->coef, (void *)j->dcdiff, (void *)j->rowmask, (void *)j->plane, (void *)j->items, (void *)j->qtabs, (void *)j->status})
            if (p) cudaFree(p);
        for (void *p : {(void *)j->h_items, (void *)j->h_qtabs, (void *)j->h_status}) if (p) cudaFreeHost(p);
        j->coef = nullptr; j->dcdiff = nullptr; j->rowmask = nullptr; j->plane = nullptr; j->items = nullptr; j->qtabs = nullptr; j->status = nullptr;
        j->h_items = nullptr; j->h_qtabs = nullptr; j->h_status = nullptr;
        const int cap = std::max(n, ctx->max_batch);
        j->coef_per_img = ((nblocks + 3) & ~(size_t)3) * 64;      // whole row-mask words per image
        j->plane_per_img = (plane + 15) & ~(size_t)15;
        CK(cudaMalloc((void **)&j->coef, (size_t)cap * j->coef_per_img * sizeof(int16_t)));
        CK(cudaMalloc((void **)&j->dcdiff, (size_t)cap * (j->coef_per_img / 64) * sizeof(int16_t)));
        CK(cudaMalloc((void **)&j->rowmask, (size_t)cap * (j->coef_per_img / 256) * sizeof(u32)));
        j->coef_dirty = true;
        CK(cudaMalloc((void **)&j->plane, (size_t)cap * j->plane_per_img));
        CK(cudaMalloc((void **)&j->items, (size_t)cap * sizeof(JpegItem)));
        CK(cudaMalloc((void **)&j->qtabs, (size_t)cap * 3 * 64 * sizeof(u16)));
        CK(cudaMalloc((void **)&j->status, (size_t)cap * sizeof(int)));
        CK(cudaMallocHost((void **)&j->h_items, (size_t)cap * sizeof(JpegItem)));
        CK(cudaMallocHost((void **)&j->h_qtabs, (size_t)cap * 3 * 64 * sizeof(u16)));
        CK(cudaMallocHost((void **)&j->h_status, (size_t)cap * sizeof(int)));
        j->n_cap = cap; j->W = g.W; j->H = g.H;
    }
    if (blob_bytes + 64 > j->blob_cap) {
        if (j->blob) cudaFree(j->blob);
        j->blob = nullptr; j->blob_cap = 0;
        const size_t cap = std::max(blob_bytes + 64, (size_t)ctx->max_batch * g.W * g.H / 4 + 4096);
        CK(cudaMalloc((void **)&j->blob, cap));
        j->blob_cap = cap;
    }
    return LSF_OK;
}

extern "C" int lsf_front_end_batch_jpeg(lsf_ctx *ctx, const uint8_t *blob, const int64_t *offsets, int n, int stages, int k, lsf_segments *out)
{
    if (!ctx) return LSF_E_ARG;
    if (!blob || !offsets || !out || n <= 0) return fail(ctx, LSF_E_ARG, "lsf_front_end_batch_jpeg: null argument or n <= 0");
    if (n > ctx->max_batch) return fail(ctx, LSF_E_CAPACITY, "lsf_front_end_batch_jpeg: n exceeds max_batch " + std::to_string(ctx->max_batch));
    ENTER(ctx);
    // ---- headers (host): geometry of the first file is the batch geometry; Huffman table sets are deduplicated ----
    jd::Image g, gi;
    std::vector<jd::Tabs> sets;
    std::vector<uint64_t> set_hash;
    std::vector<JpegItem> items(n);
    std::vector<jd::Image> imgs;      // only the quantisation tables differ per image
    imgs.reserve(n);
    size_t max_ent = 0;
    for (int i = 0; i < n; ++i) {
        const uint8_t *d = blob + offsets[i];
        const size_t len = (size_t)(offsets[i + 1] - offsets[i]);
        uint64_t h = 0;
        int rc = jd::parse(d, len, gi, nullptr, &h);
        if (rc == 0 && gi.restart) rc = -2;
        if (rc) return fail(ctx, LSF_E_ARG, "lsf_front_end_batch_jpeg: frame " + std::to_string(i) + (rc == -2 ?
                            " is not a baseline Huffman JPEG this decoder covers (8-bit, 1 or 3 components, sampling <= 2, no restart "
                            "intervals): decode it on the host and pass BGR to lsf_front_end_batch" : " is malformed"));
        if (i == 0) g = gi;
        if (gi.W != g.W || gi.H != g.H || gi.ncomp != g.ncomp || gi.bpm != g.bpm || memcmp(gi.hs, g.hs, sizeof(g.hs)) || memcmp(gi.vs, g.vs, sizeof(g.vs)) ||
            memcmp(gi.slot_dc, g.slot_dc, sizeof(g.slot_dc)) || memcmp(gi.slot_ac, g.slot_ac, sizeof(g.slot_ac)))
            return fail(ctx, LSF_E_ARG, "lsf_front_end_batch_jpeg: frame " + std::to_string(i) + " differs from frame 0 in size / sampling / table selection");
        int si = -1;
        for (size_t q = 0; q < set_hash.size(); ++q) if (set_hash[q] == h) si = (int)q;
        if (si < 0) {
            jd::Tabs t;
            jd::parse(d, len, gi, &t, nullptr);
            sets.push_back(t); set_hash.push_back(h);
            si = (int)sets.size() - 1;
        }
        items[i].ent_off = (u32)(offsets[i] - offsets[0]) + gi.ent_off; items[i].ent_len = gi.ent_len; items[i].tabs = (u32)si; items[i].pad = 0;
        max_ent = std::max(max_ent, (size_t)gi.ent_len);
        imgs.push_back(gi);
    }
    if (g.H > ctx->max_src_h || g.W > ctx->max_src_w) return fail(ctx, LSF_E_CAPACITY, "lsf_front_end_batch_jpeg: frame larger than max_src_h x max_src_w");
    const size_t blob_bytes = (size_t)(offsets[n] - offsets[0]);
    if (blob_bytes >= ((size_t)1 << 32)) return fail(ctx, LSF_E_CAPACITY, "lsf_front_end_batch_jpeg: more than 4 GB of JPEG data in one batch");
    int rc = jpeg_reserve(ctx, g, n, blob_bytes);
    if (rc) return rc;
    JpegState *j = (JpegState *)ctx->jpeg;
    const size_t cw = (max_ent + 3) / 4 + 8;
    if (cw * (size_t)n > j->clean_words) {
        if (j->clean) cudaFree(j->clean);
        j->clean = nullptr; j->clean_words = 0;
        const size_t want = std::max(cw * (size_t)n, (size_t)ctx->max_batch * ((size_t)g.W * g.H / 16 + 16));
        CK(cudaMalloc((void **)&j->clean, want * 4));
        j->clean_words = want;
    }
    if ((int)sets.size() > j->tabs_cap) {
        if (j->tabsets) cudaFree(j->tabsets);
        if (j->h_tabsets) cudaFreeHost(j->h_tabsets);
        j->tabsets = nullptr; j->h_tabsets = nullptr;
        const int cap = std::max<int>((int)sets.size(), 8);
        CK(cudaMalloc((void **)&j->tabsets, (size_t)cap * sizeof(jd::Tabs)));
        CK(cudaMallocHost((void **)&j->h_tabsets, (size_t)cap * sizeof(jd::Tabs)));
        j->tabs_cap = cap;
    }
    memcpy(j->h_items, items.data(), (size_t)n * sizeof(JpegItem));
    memcpy(j->h_tabsets, sets.data(), sets.size() * sizeof(jd::Tabs));
    for (int i = 0; i < n; ++i) memcpy(j->h_qtabs + (size_t)i * 192, imgs[i].q, 192 * sizeof(u16));
    ctx->n_events = 0;
    mark(ctx, "start");
    cudaStream_t st = ctx->st;
    // the frame buffer the decoded frames go to: the ctx's first staging buffer (nothing may be staged in it)
    if (ctx->staged[0].valid || ctx->staged[1].valid) { CK(cudaStreamSynchronize(ctx->copy_st)); ctx->staged[0].valid = ctx->staged[1].valid = false; }
    u8 *frames = ctx->stage_buf[0];
    const size_t frame_bytes = (size_t)g.W * g.H * 3;
    CK(cudaMemcpyAsync(j->blob, blob + offsets[0], blob_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(j->items, j->h_items, (size_t)n * sizeof(JpegItem), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(j->tabsets, j->h_tabsets, sets.size() * sizeof(jd::Tabs), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(j->qtabs, j->h_qtabs, (size_t)n * 192 * sizeof(u16), cudaMemcpyHostToDevice, st));
    if (j->coef_dirty) CK(cudaMemsetAsync(j->coef, 0, (size_t)j->n_cap * j->coef_per_img * sizeof(int16_t), st));
    j->coef_dirty = true;                // until k_jpeg_idct is in the stream: it leaves the buffer zero again
    CK(cudaMemsetAsync(j->rowmask, 0, (size_t)n * (j->coef_per_img / 256) * sizeof(u32), st));
    CK(cudaMemsetAsync(j->status, 0, (size_t)n * sizeof(int), st));
    mark(ctx, "jpeg_h2d");
    const size_t per_img_words = j->clean_words / (size_t)std::max(n, 1);
    k_jpeg_huff<<<n, JT, 0, st>>>(g, j->blob, j->items, j->tabsets, j->clean, per_img_words, j->coef, j->coef_per_img, j->dcdiff, j->rowmask, j->status);
    if (getenv("LSF_JPEG_SPLIT_TIMING")) mark(ctx, "jpeg_huffman");
    const long long nblk = (long long)n * g.mcux * g.mcuy * g.bpm;
    k_jpeg_idct<<<dim3((unsigned)((nblk / n + 31) / 32), n), 256, 0, st>>>(g, n, j->coef, j->coef_per_img, j->rowmask, j->qtabs, j->plane, j->plane_per_img);
    CK(cudaGetLastError());
    j->coef_dirty = false;
    if (getenv("LSF_JPEG_SPLIT_TIMING")) mark(ctx, "jpeg_idct");
    if (g.ncomp == 3 && g.hs[0] == 2 && g.vs[0] == 2 && (g.W & 7) == 0 && (frame_bytes & 7) == 0) {
        const int npx8 = g.H * (g.W / 8);
        k_jpeg_color_420<<<dim3((unsigned)((npx8 + 255) / 256), n), 256, 0, st>>>(g, n, j->plane, j->plane_per_img, frames, frame_bytes);
    } else {
        const long long npx4 = (long long)n * g.H * ((g.W + 3) / 4);
        k_jpeg_color<<<(unsigned)((npx4 + 255) / 256), 256, 0, st>>>(g, n, j->plane, j->plane_per_img, frames, frame_bytes);
    }
    g_launches += 3;
    mark(ctx, "jpeg_decode");
    ctx->events_keep = true;
    CK(cudaMemcpyAsync(j->h_status, j->status, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaGetLastError());
    ctx->jpeg_last_bytes = (long long)blob_bytes;
    rc = lsf_front_end_batch(ctx, frames, n, g.H, g.W, (size_t)g.W * 3, LSF_MEM_DEVICE, stages, k, out);
    for (int i = 0; i < n; ++i)
        if (j->h_status[i]) return fail(ctx, LSF_E_INTERNAL, "lsf_front_end_batch_jpeg: the Huffman decode of frame " + std::to_string(i) + " did not converge");
    return rc;
}
